// grid_aux.cu -- NeRFRenderer.mark_untrained_grid (nerf/renderer.py:985-1071, called once before training, nerf/utils.py:925):
// a density-grid cell that no training camera sees, or that lies outside the training AABB, is marked -1 so that the density-grid
// update skips it for good (renderer.py:1121: `valid_mask = (density_grid >= 0) & ...`).  The reference loops over 64^3 blocks of
// cells x cascades x batches of 64 cameras with ~20 torch ops and a [64, 262144, 3] intermediate each; here it is one kernel:
// one thread per (cascade, Morton cell), cameras streamed from shared memory, early exit at the first camera that sees the cell.
#include "n2m_common.cuh"
#include "../../include/n2m_b200_fused.h"

namespace n2m {
namespace {

constexpr uint32_t kCamChunk = 128;            // cameras staged per shared-memory round: 128 x (12 + 4 + 1) floats

__global__ void __launch_bounds__(256)
k_mark_untrained(const float* __restrict__ poses, uint32_t B, const float* __restrict__ intrinsics, uint32_t intr_count,
                 const float* __restrict__ cam_near, float min_near, const float* __restrict__ aabb, float bound, uint32_t cascades,
                 uint32_t H, float* __restrict__ density_grid, int32_t* __restrict__ count_out) {
    __shared__ float s_pose[kCamChunk][12];      // rows of [R | t]
    __shared__ float s_k[kCamChunk][3];          // cx/fx, cy/fy, near
    const uint32_t cells = H * H * H;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = gid < cascades * cells;
    const uint32_t cas = live ? gid / cells : 0, m = live ? gid % cells : 0;
    // cell position (renderer.py:1021-1029): world = 2 * coords / (H - 1) - 1, scaled to the cascade
    const float cb = fminf((float)(1u << cas), bound);
    const float hgs = cb / (float)H;
    const float span = cb - hgs;
    const uint32_t c[3] = {compact3(m), compact3(m >> 1), compact3(m >> 2)};
    float x[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        x[a] = __fmul_rn(__fadd_rn(__fdiv_rn(__fmul_rn(2.0f, (float)c[a]), (float)(H - 1)), -1.0f), span);
    bool in_aabb = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) in_aabb = in_aabb && x[a] >= aabb[a] - hgs && x[a] <= aabb[3 + a] + hgs;
    bool seen = false;
    for (uint32_t b0 = 0; b0 < B; b0 += kCamChunk) {
        const uint32_t nb = min(kCamChunk, B - b0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb * 12; i += blockDim.x) {
            const uint32_t cam = i / 12, e = i % 12;
            s_pose[cam][e] = poses[(size_t)(b0 + cam) * 16 + (e / 4) * 4 + (e % 4)];
        }
        for (uint32_t cam = threadIdx.x; cam < nb; cam += blockDim.x) {
            const float* K = intrinsics + (size_t)(intr_count > 1 ? (b0 + cam) : 0) * 4;
            s_k[cam][0] = __fdiv_rn(K[2], K[0]); s_k[cam][1] = __fdiv_rn(K[3], K[1]);
            s_k[cam][2] = cam_near ? cam_near[b0 + cam] : min_near;
        }
        __syncthreads();
        if (live && in_aabb && !seen) {
            for (uint32_t cam = 0; cam < nb; ++cam) {
                const float* P = s_pose[cam];
                const float dx = x[0] - P[3], dy = x[1] - P[7], dz = x[2] - P[11];
                // (xyz - t) @ R  (row vector times the c2w rotation == world -> camera), then z negated (renderer.py:1042-1044)
                const float cx_ = __fmaf_rn(dz, P[8], __fmaf_rn(dy, P[4], dx * P[0]));
                const float cy_ = __fmaf_rn(dz, P[9], __fmaf_rn(dy, P[5], dx * P[1]));
                const float cz_ = -__fmaf_rn(dz, P[10], __fmaf_rn(dy, P[6], dx * P[2]));
                if (cz_ > s_k[cam][2] && fabsf(cx_) < s_k[cam][0] * cz_ + hgs * 2 && fabsf(cy_) < s_k[cam][1] * cz_ + hgs * 2) { seen = true; break; }
            }
        }
    }
    if (!live) return;
    if (!(seen && in_aabb)) {
        density_grid[gid] = -1.0f;
        if (count_out) atomicAdd(count_out, 1);
    }
}

}  // namespace
}  // namespace n2m

using namespace n2m;

extern "C" int n2m_mark_untrained_grid(const float* poses, uint32_t num_poses, const float* intrinsics, uint32_t intr_count,
                                       const float* cam_near, float min_near, const float* aabb, float bound, uint32_t cascades,
                                       uint32_t H, float* density_grid, int32_t* count_out, n2m_stream_t stream) {
    N2M_REQUIRE(poses && intrinsics && aabb && density_grid, "mark_untrained_grid", "null pointer");
    N2M_REQUIRE(num_poses > 0 && (intr_count == 1 || intr_count == num_poses) && H > 1 && H <= 1024 && cascades > 0,
                "mark_untrained_grid", "bad arguments");
    cudaStream_t st = as_stream(stream);
    if (count_out) cudaMemsetAsync(count_out, 0, sizeof(int32_t), st);
    const uint32_t n = cascades * H * H * H;
    k_mark_untrained<<<div_up(n, 256u), 256, 0, st>>>(poses, num_poses, intrinsics, intr_count, cam_near, min_near, aabb, bound, cascades, H,
                                                    density_grid, count_out);
    return check_launch("mark_untrained_grid");
}
