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

}  // extern "C"
