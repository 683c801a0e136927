// stage1.cu -- the texture-MLP step of the reference's stage 1 (NeRFRenderer.render_stage1, nerf/renderer.py:806-935, and the
// stage-1 branch of Trainer.train_step, nerf/utils.py:703-716) around the rasterizer of raster.cu and the stage-0 kernels:
//
//   n2m_rasterize                     dr.rasterize at the super-sampled resolution (h, w) = ssaa * (h0, w0)        renderer.py:824-860
//   n2m_s1_points                     covered pixels -> surface points: attribute interpolation of the vertex positions (dr.interpolate,
//                                     renderer.py:862), nearest-neighbour up-sampling of the view directions (:828-829), compaction
//                                     (xyzs[mask_flatten], :875-877) -- one kernel, no boolean-mask host sync; writes the march-record
//                                     form (t = 0, origin = point) the stage-0 gather / backward kernels consume
//   n2m_s0_encode_points, n2m_s0_mlp_fwd   self.rgb(x, d) (network.py:170-189) on tensor cores (sigma is computed and ignored)
//   n2m_s1_loss                       alphas * rgbs, ssaa average (scale_img_hwc bilinear at factor 2 == 2x2 mean, :899-901), background
//                                     mix (:907), MSE (+ mask) loss (utils.py:707-712) and its gradient w.r.t. every covered pixel's rgb
//   n2m_s0_bwd_fused_part, n2m_s0_adam_*   backward of the colour MLPs + colour hash table, optimizer
// With dr.antialias (renderer.py:886-887; csrc/antialias.cu) the middle of the step becomes
//   n2m_s1_rgba                       scatter of the per-point colours into a full-resolution (r, g, b, mask) image (`rgbs[mask_flatten] =
//                                     mask_rgbs`, `alphas = mask`, :881-884) -- ONE 4-channel image, so that one antialias launch serves both
//                                     of the reference's calls (the operator is linear and channel-wise)
//   n2m_antialias_forward             alphas, rgbs = dr.antialias(...)                                                  :886-887
//   n2m_s1_loss_aa                    clamp, alphas * rgbs, ssaa average, background mix, loss; gradient w.r.t. the antialiased image
//   n2m_antialias_backward            -> gradient w.r.t. the (r, g, b, mask) image and w.r.t. the clip-space vertices (vertices_offsets)
//   n2m_s1_dout                       gather of the colour gradient back to the compacted points
#include "n2m_common.cuh"
#include "../../include/n2m_b200_raster.h"

namespace n2m {
namespace {

__global__ void __launch_bounds__(256)
k_s1_points(const float4* __restrict__ rast, const float* __restrict__ verts, const int32_t* __restrict__ tri,
            const float* __restrict__ rays_d, uint32_t h, uint32_t w, uint32_t ssaa, uint32_t cap, int32_t* __restrict__ counters,
            int32_t* __restrict__ inv, float* __restrict__ pts, float* __restrict__ pdirs, float4* __restrict__ recs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = h * w;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) r = rast[i];
    const bool cov = i < n && r.w > 0.f;
    const uint32_t mask = __ballot_sync(0xffffffffu, cov);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (mask != 0) {
        const int leader = __ffs(mask) - 1;
        if ((int)lane == leader) base = (uint32_t)atomicAdd(counters + 0, (int)__popc(mask));       // [0]: covered pixels (uncapped)
        base = __shfl_sync(0xffffffffu, base, leader);
    }
    if (i >= n) return;
    int32_t slot = -1;
    if (cov) {
        const uint32_t k = base + __popc(mask & ((1u << lane) - 1u));
        if (k < cap) {
            slot = (int32_t)k;
            const uint32_t f = (uint32_t)r.w - 1u;
            const int i0 = tri[3 * f], i1 = tri[3 * f + 1], i2 = tri[3 * f + 2];
            const float u = r.x, v = r.y, ww = 1.f - r.x - r.y;
            const uint32_t y = i / w, x = i % w;
            const uint32_t q = (y / ssaa) * (w / ssaa) + x / ssaa;            // nearest-neighbour source pixel of the low-res direction
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                pts[3 * (size_t)k + a] = u * __ldg(verts + 3 * (size_t)i0 + a) + v * __ldg(verts + 3 * (size_t)i1 + a) + ww * __ldg(verts + 3 * (size_t)i2 + a);
                pdirs[3 * (size_t)k + a] = __ldg(rays_d + 3 * (size_t)q + a);
            }
            recs[k] = make_float4(0.f, 0.f, 0.f, __int_as_float((int)k));      // t = 0: the "sample" is its origin
        }
    }
    inv[i] = slot;
}

// counters[1] = min(counters[0], cap)  (the sample count the stage-0 kernels read), [2] = overflow flag
__global__ void k_s1_finish_count(int32_t* __restrict__ counters, uint32_t cap) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int32_t c = counters[0];
    counters[1] = c < (int32_t)cap ? c : (int32_t)cap;
    counters[2] = c > (int32_t)cap ? 1 : 0;
    counters[3] = 0;
}

// one thread per low-resolution pixel
__global__ void __launch_bounds__(256)
k_s1_loss(const float4* __restrict__ out, const int32_t* __restrict__ inv, const float* __restrict__ gt, uint32_t gt_channels,
          const float* __restrict__ bg, uint32_t h0, uint32_t w0, uint32_t ssaa, float lambda_mask, const float* __restrict__ loss_scale,
          float4* __restrict__ dout, float* __restrict__ image, float* __restrict__ weights_sum, float* __restrict__ loss_out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t Q = h0 * w0;
    float my_loss = 0.f;
    if (q < Q) {
        const uint32_t y0 = q / w0, x0 = q % w0, w = w0 * ssaa;
        const float inv_s2 = 1.0f / (float)(ssaa * ssaa);
        float r = 0.f, g = 0.f, b = 0.f, a = 0.f;
        for (uint32_t dy = 0; dy < ssaa; ++dy)
            for (uint32_t dx = 0; dx < ssaa; ++dx) {
                const int32_t k = inv[(size_t)(y0 * ssaa + dy) * w + x0 * ssaa + dx];
                if (k >= 0) {
                    const float4 o = out[k];
                    // alphas * rgbs with both clamped to [0, 1] (renderer.py:886-889; the colour is already in [0, 1])
                    r += fminf(fmaxf(o.y, 0.f), 1.f); g += fminf(fmaxf(o.z, 0.f), 1.f); b += fminf(fmaxf(o.w, 0.f), 1.f); a += 1.f;
                }
            }
        r *= inv_s2; g *= inv_s2; b *= inv_s2; a *= inv_s2;
        const float T = 1.f - a;
        const float b0 = bg[3 * q], b1 = bg[3 * q + 1], b2 = bg[3 * q + 2];
        const float pr = r + T * b0, pg = g + T * b1, pb = b + T * b2;
        float t0, t1, t2, m = 0.f;
        if (gt_channels == 4) {          // utils.py:662-667
            m = gt[4 * q + 3];
            t0 = gt[4 * q] * m + b0 * (1 - m); t1 = gt[4 * q + 1] * m + b1 * (1 - m); t2 = gt[4 * q + 2] * m + b2 * (1 - m);
        } else {
            t0 = gt[3 * q]; t1 = gt[3 * q + 1]; t2 = gt[3 * q + 2];
        }
        const float e0 = pr - t0, e1 = pg - t1, e2 = pb - t2;
        my_loss = (e0 * e0 + e1 * e1 + e2 * e2) * (1.0f / 3.0f);
        if (gt_channels == 4 && lambda_mask > 0) { const float em = a - m; my_loss += lambda_mask * em * em; }
        image[3 * q] = pr; image[3 * q + 1] = pg; image[3 * q + 2] = pb;
        weights_sum[q] = a;
        const float sc = loss_scale[0] / (float)Q * (2.0f / 3.0f) * inv_s2;
        const float4 d = make_float4(0.f, sc * e0, sc * e1, sc * e2);
        for (uint32_t dy = 0; dy < ssaa; ++dy)
            for (uint32_t dx = 0; dx < ssaa; ++dx) {
                const int32_t k = inv[(size_t)(y0 * ssaa + dy) * w + x0 * ssaa + dx];
                if (k >= 0) dout[k] = d;
            }
        my_loss /= (float)Q;
    }
    // block reduction -> one atomic per block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = my_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(loss_out, s);
    }
}

// ---- antialiased variant ----
__global__ void __launch_bounds__(256)
k_s1_rgba(const float4* __restrict__ out, const int32_t* __restrict__ inv, uint32_t n, float4* __restrict__ rgba) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t k = inv[i];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k >= 0) { const float4 o = out[k]; v = make_float4(o.y, o.z, o.w, 1.f); }
    rgba[i] = v;
}

__global__ void __launch_bounds__(256)
k_s1_dout(const float4* __restrict__ grad_rgba, const int32_t* __restrict__ inv, uint32_t n, float4* __restrict__ dout) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t k = inv[i];
    if (k >= 0) { const float4 g = grad_rgba[i]; dout[k] = make_float4(0.f, g.x, g.y, g.z); }
}

// one thread per low-resolution pixel; aa [h*w] float4 = antialiased (r, g, b, alpha); d_aa = d loss / d aa * loss_scale
__global__ void __launch_bounds__(256)
k_s1_loss_aa(const float4* __restrict__ aa, const float* __restrict__ gt, uint32_t gt_channels, const float* __restrict__ bg, uint32_t h0,
             uint32_t w0, uint32_t ssaa, float lambda_mask, const float* __restrict__ loss_scale, float4* __restrict__ d_aa,
             float* __restrict__ image, float* __restrict__ weights_sum, float* __restrict__ loss_out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t Q = h0 * w0;
    float my_loss = 0.f;
    if (q < Q) {
        const uint32_t y0 = q / w0, x0 = q % w0, w = w0 * ssaa;
        const float inv_s2 = 1.0f / (float)(ssaa * ssaa);
        float r = 0.f, g = 0.f, b = 0.f, a = 0.f;
        for (uint32_t dy = 0; dy < ssaa; ++dy)
            for (uint32_t dx = 0; dx < ssaa; ++dx) {
                const float4 v = aa[(size_t)(y0 * ssaa + dy) * w + x0 * ssaa + dx];
                const float al = clampf(v.w, 0.f, 1.f);                                 // .clamp(0, 1) of both (renderer.py:886-887)
                r += al * clampf(v.x, 0.f, 1.f); g += al * clampf(v.y, 0.f, 1.f); b += al * clampf(v.z, 0.f, 1.f); a += al;
            }
        r *= inv_s2; g *= inv_s2; b *= inv_s2; a *= inv_s2;
        const float T = 1.f - a;
        const float b0 = bg[3 * q], b1 = bg[3 * q + 1], b2 = bg[3 * q + 2];
        const float pr = r + T * b0, pg = g + T * b1, pb = b + T * b2;
        float t0, t1, t2, m = 0.f;
        if (gt_channels == 4) {
            m = gt[4 * q + 3];
            t0 = gt[4 * q] * m + b0 * (1 - m); t1 = gt[4 * q + 1] * m + b1 * (1 - m); t2 = gt[4 * q + 2] * m + b2 * (1 - m);
        } else {
            t0 = gt[3 * q]; t1 = gt[3 * q + 1]; t2 = gt[3 * q + 2];
        }
        const float e0 = pr - t0, e1 = pg - t1, e2 = pb - t2;
        my_loss = (e0 * e0 + e1 * e1 + e2 * e2) * (1.0f / 3.0f);
        float dmask = 0.f;
        if (gt_channels == 4 && lambda_mask > 0) {
            const float em = a - m;
            my_loss += lambda_mask * em * em;
            dmask = loss_scale[0] / (float)Q * 2.0f * lambda_mask * em * inv_s2;
        }
        image[3 * q] = pr; image[3 * q + 1] = pg; image[3 * q + 2] = pb;
        weights_sum[q] = a;
        const float sc = loss_scale[0] / (float)Q * (2.0f / 3.0f) * inv_s2;
        for (uint32_t dy = 0; dy < ssaa; ++dy)
            for (uint32_t dx = 0; dx < ssaa; ++dx) {
                const size_t i = (size_t)(y0 * ssaa + dy) * w + x0 * ssaa + dx;
                const float4 v = aa[i];
                const float al = clampf(v.w, 0.f, 1.f);
                const float cr = clampf(v.x, 0.f, 1.f), cg = clampf(v.y, 0.f, 1.f), cb = clampf(v.z, 0.f, 1.f);
                float4 d;
                // clamp passes the gradient inside [0, 1] (bounds included, as torch.clamp does)
                d.x = (v.x >= 0.f && v.x <= 1.f) ? sc * e0 * al : 0.f;
                d.y = (v.y >= 0.f && v.y <= 1.f) ? sc * e1 * al : 0.f;
                d.z = (v.z >= 0.f && v.z <= 1.f) ? sc * e2 * al : 0.f;
                d.w = (v.w >= 0.f && v.w <= 1.f) ? sc * (e0 * (cr - b0) + e1 * (cg - b1) + e2 * (cb - b2)) + dmask : 0.f;
                d_aa[i] = d;
            }
        my_loss /= (float)Q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = my_loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        atomicAdd(loss_out, s);
    }
}

// ---- vertex-offset optimizer (NeRFRenderer.vertices_offsets: renderer.py:160,180 -- Adam group with lr_vert; regularisers
// utils.py:750-779: lambda_lap * laplacian_smooth_loss (uniform Laplacian, utils.py:176-221) + lambda_offsets * mean(sum(offsets^2))) ----
// y += L x for the uniform Laplacian L = D - A, one thread per slot of the edge hash of csrc/antialias.cu (one slot = one undirected edge)
__global__ void __launch_bounds__(256)
k_s1_laplacian(const unsigned long long* __restrict__ keys, uint32_t slots, const float* __restrict__ x, float* __restrict__ y) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const unsigned long long k = keys[i];
    if (k == ~0ull) return;
    const uint32_t a = (uint32_t)(k >> 32), b = (uint32_t)(k & 0xffffffffull);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float d = x[3 * (size_t)a + c] - x[3 * (size_t)b + c];
        atomicAdd(y + 3 * (size_t)a + c, d);
        atomicAdd(y + 3 * (size_t)b + c, -d);
    }
}

// u [V,3] = L v  ->  u_i / |u_i| in place (0 where |u_i| = 0: the subgradient torch.norm uses), loss_out[0] += lambda_lap * mean |u_i|
__global__ void __launch_bounds__(256)
k_s1_lap_normalize(float* __restrict__ u, uint32_t V, float lambda_lap, float* __restrict__ loss_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float n = 0.f;
    if (i < V) {
        const float a = u[3 * (size_t)i], b = u[3 * (size_t)i + 1], c = u[3 * (size_t)i + 2];
        n = sqrtf(a * a + b * b + c * c);
        const float r = n > 0.f ? __fdiv_rn(1.f, n) : 0.f;
        u[3 * (size_t)i] = a * r; u[3 * (size_t)i + 1] = b * r; u[3 * (size_t)i + 2] = c * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (loss_out && (threadIdx.x & 31) == 0 && n != 0.f) atomicAdd(loss_out, lambda_lap * n / (float)V);
}

// non-finite scan of the loss-scaled clip-space gradient -> found_inf (GradScaler.unscale_ over the vertices_offsets group)
__global__ void __launch_bounds__(256)
k_s1_vert_check(const float* __restrict__ g, uint32_t n, float* __restrict__ st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < n && !isfinite(g[i]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) st[3] = 1.f;
}

// grad = (grad_vclip . mvp[:, :3]) / loss_scale + (lambda_lap / V) * (L w) + (2 lambda_offsets / V) * offsets; Adam(eps) on the offsets with
// its own step count vst[0]; vertices = base + offsets.  Skipped as a whole when found_inf is set (st[3]).
__global__ void __launch_bounds__(256)
k_s1_vert_adam(const float4* __restrict__ grad_vclip, const float* __restrict__ mvp, const float* __restrict__ lap_grad, const float* __restrict__ base,
               float* __restrict__ offsets, float* __restrict__ m, float* __restrict__ v, float* __restrict__ vertices, float* __restrict__ grad_out,
               uint32_t V, float lambda_lap, float lambda_offsets, float lr, float eps, const float* __restrict__ st, const float* __restrict__ vst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    if (lr < 0.f) lr = vst[1];                                          // learning rate kept on the device (graph-replayed steps)
    const bool skip = st[3] != 0.f;
    const float inv_scale = st[7];
    const float t = vst[0] + (skip ? 0.f : 1.f);                       // this step's count (k_s1_vert_tick advances it afterwards)
    const float bc1 = 1.f - powf(0.9f, fmaxf(t, 1.f)), bc2s = sqrtf(1.f - powf(0.999f, fmaxf(t, 1.f)));
    const float4 gc = grad_vclip[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t j = 3 * (size_t)i + c;
        const float gi = (gc.x * mvp[c] + gc.y * mvp[4 + c] + gc.z * mvp[8 + c] + gc.w * mvp[12 + c]) * inv_scale;
        const float off = offsets[j];
        const float g = gi + (lambda_lap > 0.f ? lambda_lap / (float)V * lap_grad[j] : 0.f) + 2.f * lambda_offsets / (float)V * off;
        if (grad_out) grad_out[j] = g;
        if (skip) continue;
        const float mi = 0.9f * m[j] + 0.1f * g;
        const float vi = 0.999f * v[j] + 0.001f * g * g;
        m[j] = mi; v[j] = vi;
        const float denom = __fdiv_rn(__fsqrt_rn(vi), bc2s) + eps;
        const float o2 = off - __fdiv_rn(lr, bc1) * __fdiv_rn(mi, denom);
        offsets[j] = o2;
        vertices[j] = base[j] + o2;
    }
}

__global__ void k_s1_vert_tick(const float* __restrict__ st, float* __restrict__ vst) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && st[3] == 0.f) vst[0] += 1.f;
}

}  // namespace
}  // namespace n2m

using namespace n2m;

extern "C" {

int n2m_s1_points(const float* rast, const float* verts, const int32_t* tri, const float* rays_d, uint32_t h, uint32_t w, uint32_t ssaa,
                  uint32_t cap, int32_t* counters, int32_t* inv, float* pts, float* pdirs, void* recs, n2m_stream_t stream) {
    N2M_REQUIRE(rast && verts && tri && rays_d && counters && inv && pts && pdirs && recs, "s1_points", "null pointer");
    N2M_REQUIRE(ssaa >= 1 && h % ssaa == 0 && w % ssaa == 0 && h > 0 && w > 0, "s1_points", "resolution must be a multiple of ssaa");
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(counters, 0, 4 * sizeof(int32_t), st);
    k_s1_points<<<div_up(h * w, 256u), 256, 0, st>>>(reinterpret_cast<const float4*>(rast), verts, tri, rays_d, h, w, ssaa, cap, counters, inv,
                                                   pts, pdirs, static_cast<float4*>(recs));
    if (int e = check_launch("s1_points")) return e;
    k_s1_finish_count<<<1, 32, 0, st>>>(counters, cap);
    return check_launch("s1_points(count)");
}

int n2m_s1_loss(const void* out, const int32_t* inv, const float* gt, uint32_t gt_channels, const float* bg, uint32_t h0, uint32_t w0,
                uint32_t ssaa, float lambda_mask, const float* loss_scale, void* dout, float* image, float* weights_sum, float* loss_out,
                n2m_stream_t stream) {
    N2M_REQUIRE(out && inv && gt && bg && loss_scale && dout && image && weights_sum && loss_out, "s1_loss", "null pointer");
    N2M_REQUIRE(gt_channels == 3 || gt_channels == 4, "s1_loss", "gt must have 3 or 4 channels");
    k_s1_loss<<<div_up(h0 * w0, 256u), 256, 0, as_stream(stream)>>>(static_cast<const float4*>(out), inv, gt, gt_channels, bg, h0, w0, ssaa,
                                                                   lambda_mask, loss_scale, static_cast<float4*>(dout), image, weights_sum, loss_out);
    return check_launch("s1_loss");
}

int n2m_s1_rgba(const void* out, const int32_t* inv, uint32_t num_pixels, void* rgba, n2m_stream_t stream) {
    N2M_REQUIRE(out && inv && rgba, "s1_rgba", "null pointer");
    if (num_pixels == 0) return 0;
    k_s1_rgba<<<div_up(num_pixels, 256u), 256, 0, as_stream(stream)>>>(static_cast<const float4*>(out), inv, num_pixels, static_cast<float4*>(rgba));
    return check_launch("s1_rgba");
}

int n2m_s1_dout(const void* grad_rgba, const int32_t* inv, uint32_t num_pixels, void* dout, n2m_stream_t stream) {
    N2M_REQUIRE(grad_rgba && inv && dout, "s1_dout", "null pointer");
    if (num_pixels == 0) return 0;
    k_s1_dout<<<div_up(num_pixels, 256u), 256, 0, as_stream(stream)>>>(static_cast<const float4*>(grad_rgba), inv, num_pixels, static_cast<float4*>(dout));
    return check_launch("s1_dout");
}

int n2m_s1_loss_aa(const void* aa, const float* gt, uint32_t gt_channels, const float* bg, uint32_t h0, uint32_t w0, uint32_t ssaa,
                   float lambda_mask, const float* loss_scale, void* d_aa, float* image, float* weights_sum, float* loss_out,
                   n2m_stream_t stream) {
    N2M_REQUIRE(aa && gt && bg && loss_scale && d_aa && image && weights_sum && loss_out, "s1_loss_aa", "null pointer");
    N2M_REQUIRE(gt_channels == 3 || gt_channels == 4, "s1_loss_aa", "gt must have 3 or 4 channels");
    N2M_REQUIRE(ssaa >= 1, "s1_loss_aa", "ssaa must be >= 1");
    k_s1_loss_aa<<<div_up(h0 * w0, 256u), 256, 0, as_stream(stream)>>>(static_cast<const float4*>(aa), gt, gt_channels, bg, h0, w0, ssaa, lambda_mask,
                                                                      loss_scale, static_cast<float4*>(d_aa), image, weights_sum, loss_out);
    return check_launch("s1_loss_aa");
}

int n2m_s1_vert_check(const float* grad_vclip, uint32_t V, float* opt_state, n2m_stream_t stream) {
    N2M_REQUIRE(grad_vclip && opt_state, "s1_vert_check", "null pointer");
    if (V == 0) return 0;
    k_s1_vert_check<<<div_up(4 * V, 256u), 256, 0, as_stream(stream)>>>(grad_vclip, 4 * V, opt_state);
    return check_launch("s1_vert_check");
}

int n2m_s1_vert_step(const float* grad_vclip, const float* mvp, const void* topo_keys, uint32_t topo_slots, const float* base, float* offsets,
                     float* m, float* v, float* vertices, float* scratch, float* grad_out, uint32_t V, float lambda_lap, float lambda_offsets,
                     float lr_vert, float eps, const float* opt_state, float* vert_state, float* loss_out, n2m_stream_t stream) {
    N2M_REQUIRE(grad_vclip && mvp && base && offsets && m && v && vertices && scratch && opt_state && vert_state, "s1_vert_step", "null pointer");
    N2M_REQUIRE(lambda_lap <= 0.f || (topo_keys && topo_slots > 0), "s1_vert_step", "the Laplacian term needs the mesh's edge hash");
    if (V == 0) return 0;
    cudaStream_t st = as_stream(stream);
    float* u = scratch;                       // [V,3]: L v, then its row-normalised form
    float* lg = scratch + 3 * (size_t)V;      // [V,3]: L (u / |u|)
    if (lambda_lap > 0.f) {
        const unsigned long long* keys = static_cast<const unsigned long long*>(topo_keys);
        cudaError_t e = cudaMemsetAsync(scratch, 0, 6 * (size_t)V * sizeof(float), st);
        if (e != cudaSuccess) return fail("s1_vert_step(memset)", cudaGetErrorString(e));
        k_s1_laplacian<<<div_up(topo_slots, 256u), 256, 0, st>>>(keys, topo_slots, vertices, u);
        if (int err = check_launch("s1_vert_step(L v)")) return err;
        k_s1_lap_normalize<<<div_up(V, 256u), 256, 0, st>>>(u, V, lambda_lap, loss_out);
        if (int err = check_launch("s1_vert_step(normalize)")) return err;
        k_s1_laplacian<<<div_up(topo_slots, 256u), 256, 0, st>>>(keys, topo_slots, u, lg);
        if (int err = check_launch("s1_vert_step(L w)")) return err;
    }
    k_s1_vert_adam<<<div_up(V, 256u), 256, 0, st>>>(reinterpret_cast<const float4*>(grad_vclip), mvp, lg, base, offsets, m, v, vertices, grad_out, V,
                                                   lambda_lap, lambda_offsets, lr_vert, eps, opt_state, vert_state);
    if (int err = check_launch("s1_vert_step(adam)")) return err;
    k_s1_vert_tick<<<1, 32, 0, st>>>(opt_state, vert_state);
    return check_launch("s1_vert_step(tick)");
}

}  // extern "C"
