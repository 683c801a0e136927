/* n2m_b200_raster.h -- C ABI of the stage-1 mesh path of libn2m_b200.so (BASELINE config 5).
 *
 * The reference's stage 1 (NeRFRenderer.render_stage1, nerf/renderer.py:806-935) rasterizes the refined mesh with the third-party
 * nvdiffrast library and runs the colour MLPs on the covered pixels.  These entry points replace the two nvdiffrast operators on
 * that path, with nvdiffrast's documented output convention; the reference-side binding is `nerf2mesh_b200.raster`
 * (`rasterize`, `interpolate` with the argument order of `nvdiffrast.torch`), see INTEGRATION.md.
 *
 *   n2m_rasterize             dr.rasterize(glctx, pos, tri, (H, W))   (renderer.py:860; also :338, :968)
 *       pos  [V,4] f32 clip-space vertices, tri [F,3] i32, rast [H,W,4] f32 = (u, v, z/w, triangle_id + 1), zeros where empty;
 *       pixel (x, y) samples NDC ((x+0.5)/W*2-1, (y+0.5)/H*2-1); (u, v) perspective-correct barycentrics of vertices 0 / 1;
 *       vis [H*W] u64 and queue [F+1] u32 are caller-allocated scratch.  Fragments outside -1 <= z/w <= 1 are clipped; triangles that
 *       cross the camera plane (some w <= 0) are rasterised in homogeneous coordinates (the part in front of the near plane).
 *   n2m_interpolate_forward   dr.interpolate(attr, rast, tri)         (renderer.py:862-863): out [H*W,A] = u a0 + v a1 + (1-u-v) a2
 *   n2m_interpolate_backward  its gradient w.r.t. attr [V,A] (accumulated into the caller's zero-initialised buffer)
 *   n2m_compact_covered       xyzs[mask], dirs[mask] of renderer.py:865-880 without the boolean-mask host sync: covered pixel
 *                             indices + their positions / view directions, count in *counter (device)
 *   n2m_antialias_*           dr.antialias(color, rast, pos, tri, pos_gradient_boost=b)   (renderer.py:886-887), csrc/antialias.cu:
 *       topology: edge -> opposing-vertex hash of the mesh (the library's "topology hash"), `keys` [slots] u64 and `opp` [2*slots] i32
 *       caller-allocated, slots = n2m_antialias_topology_slots(F) (a power of two >= 3 F); rebuilt only when `tri` changes;
 *       forward: out [H*W,C] = color + silhouette blends, C = 1..4, out must not alias color;
 *       backward: grad_color [H*W,C] (may be NULL) and grad_pos [V,4] (may be NULL; ACCUMULATED into the caller's zero-initialised
 *       buffer: gradients w.r.t. clip-space x, y, w, multiplied by pos_gradient_boost), from grad_out [H*W,C] and the forward inputs.
 */
#ifndef N2M_B200_RASTER_H
#define N2M_B200_RASTER_H

#include "n2m_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

int n2m_rasterize(const float* pos, uint32_t V, const int32_t* tri, uint32_t F, uint32_t H, uint32_t W, void* vis, uint32_t* queue,
                  float* rast, n2m_stream_t stream);
int n2m_interpolate_forward(const float* attr, uint32_t V, uint32_t A, const float* rast, const int32_t* tri, uint32_t num_pixels,
                            float* out, n2m_stream_t stream);
int n2m_interpolate_backward(const float* grad_out, const float* rast, const int32_t* tri, uint32_t num_pixels, uint32_t V, uint32_t A,
                             float* grad_attr, n2m_stream_t stream);
int n2m_compact_covered(const float* rast, const float* xyz, const float* dirs, uint32_t num_pixels, uint32_t cap, int32_t* counter,
                        int32_t* pix, float* pts, float* pdirs, n2m_stream_t stream);

uint32_t n2m_antialias_topology_slots(uint32_t F);
int n2m_antialias_topology(const int32_t* tri, uint32_t F, void* keys, int32_t* opp, uint32_t slots, n2m_stream_t stream);
int n2m_antialias_forward(const float* color, const float* rast, const float* pos, const int32_t* tri, const void* keys, const int32_t* opp,
                          uint32_t slots, uint32_t H, uint32_t W, uint32_t C, float* out, n2m_stream_t stream);
int n2m_antialias_backward(const float* color, const float* rast, const float* pos, const int32_t* tri, const void* keys, const int32_t* opp,
                           uint32_t slots, uint32_t H, uint32_t W, uint32_t C, const float* grad_out, float pos_gradient_boost,
                           float* grad_color, float* grad_pos, n2m_stream_t stream);

/* ---- stage-1 texture-MLP step (csrc/stage1.cu; host side nerf2mesh_b200/stage1.py) ----
 * n2m_s1_points: covered pixels of `rast` [h,w,4] -> compacted surface points for the stage-0 gather / MLP / backward kernels:
 *   pts [cap,3] = dr.interpolate(vertices, rast, tri) at the covered pixels, pdirs [cap,3] = rays_d [h/ssaa * w/ssaa, 3] up-sampled
 *   nearest-neighbour (renderer.py:828-829), recs [cap] float4 = (0, 0, 0, k) (a march record whose sample is its own origin),
 *   inv [h*w] = slot of each pixel or -1; counters (>= 16 int32): [0] covered pixels, [1] min([0], cap), [2] overflow flag.
 * n2m_s1_loss: per low-res pixel: image = mean(alpha * rgb) + (1 - mean(alpha)) * bg, MSE (+ lambda_mask * mask term for rgba targets)
 *   averaged over the h0*w0 pixels into loss_out[0]; dout [cap] float4 = (0, dL/drgb) * loss_scale for the covered pixels. */
int n2m_s1_points(const float* rast, const float* verts, const int32_t* tri, const float* rays_d, uint32_t h, uint32_t w, uint32_t ssaa,
                  uint32_t cap, int32_t* counters, int32_t* inv, float* pts, float* pdirs, void* recs, n2m_stream_t stream);
int n2m_s1_loss(const void* out, const int32_t* inv, const float* gt, uint32_t gt_channels, const float* bg, uint32_t h0, uint32_t w0,
                uint32_t ssaa, float lambda_mask, const float* loss_scale, void* dout, float* image, float* weights_sum, float* loss_out,
                n2m_stream_t stream);


/* antialiased variant of the step (renderer.py:881-907 with dr.antialias): n2m_s1_rgba scatters the per-point colours `out` [cap] float4
 * (sigma, r, g, b) into rgba [h*w] float4 = (r, g, b, mask) (zero where uncovered); after n2m_antialias_forward (C = 4),
 * n2m_s1_loss_aa evaluates clamp, alphas * rgbs, the ssaa average, the background mix and the loss of n2m_s1_loss on the antialiased
 * image aa [h*w] float4 and writes d_aa = d loss / d aa * loss_scale; after n2m_antialias_backward, n2m_s1_dout gathers the colour
 * part of the image gradient into dout [cap] float4 = (0, dr, dg, db) for the covered pixels. */
int n2m_s1_rgba(const void* out, const int32_t* inv, uint32_t num_pixels, void* rgba, n2m_stream_t stream);
int n2m_s1_loss_aa(const void* aa, const float* gt, uint32_t gt_channels, const float* bg, uint32_t h0, uint32_t w0, uint32_t ssaa,
                   float lambda_mask, const float* loss_scale, void* d_aa, float* image, float* weights_sum, float* loss_out,
                   n2m_stream_t stream);
int n2m_s1_dout(const void* grad_rgba, const int32_t* inv, uint32_t num_pixels, void* dout, n2m_stream_t stream);

/* vertex-offset optimizer of stage 1 (`vertices_offsets`: nn.Parameter at renderer.py:160, Adam group with lr_vert at :180; regularisers
 * utils.py:750-779).  n2m_s1_vert_check: non-finite scan of the loss-scaled clip-space gradient grad_vclip [V,4] into found_inf
 * (opt_state[3]) -- call it BEFORE the optimizer head.  n2m_s1_vert_step (between the optimizer head and its post kernel):
 *   grad = (grad_vclip . mvp[:, :3]) / loss_scale + lambda_lap * d/dv mean_i |(L v)_i| + lambda_offsets * d/doff mean_i sum_c off_ic^2
 * with the uniform Laplacian L = D - A over the unique edges of the mesh (the edge hash of n2m_antialias_topology), Adam(0.9, 0.999, eps)
 * with its own step count vert_state[0] on offsets [V,3], vertices = base + offsets; skipped when found_inf is set.  scratch [6 V] f32;
 * grad_out [V,3] (nullable) receives the total gradient; loss_out (nullable) += lambda_lap * mean |L v| (the offsets term is left to
 * the caller: it needs no kernel).  mvp [4,4] row-major on the device; lr_vert < 0: the learning rate is read from vert_state[1]. */
int n2m_s1_vert_check(const float* grad_vclip, uint32_t V, float* opt_state, n2m_stream_t stream);
int n2m_s1_vert_step(const float* grad_vclip, const float* mvp, const void* topo_keys, uint32_t topo_slots, const float* base, float* offsets,
                     float* m, float* v, float* vertices, float* scratch, float* grad_out, uint32_t V, float lambda_lap, float lambda_offsets,
                     float lr_vert, float eps, const float* opt_state, float* vert_state, float* loss_out, n2m_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* N2M_B200_RASTER_H */
