// mlp_common.cuh -- constants and device helpers shared by the tensor-core MLP kernels (mlp_tc.cu) and the fused kernels (fused.cu):
// packed-weight layout, flat parameter layout, epilogue helpers (TMEM accumulator row -> fp16 row of a chunk-major smem tile).
#pragma once
#include "n2m_common.cuh"
#include "tc05.cuh"
#include "s0_geom.cuh"

namespace n2m {
namespace {

constexpr uint32_t kChunk = kChunkBytes;     // 2048: one 8-column chunk of a 128-row tile (kTile / kTileBytes: s0_geom.cuh)

// ---- packed weights: fp16 chunk-major tiles [rows = out (padded), cols = in (padded)] --------------
constexpr uint32_t W_C1 = 0;                        // color_net.0   64 x 64  (in: enc tile cols)
constexpr uint32_t W_C2 = W_C1 + 64 * 64 * 2;       // color_net.1   64 x 64
constexpr uint32_t W_C3 = W_C2 + 64 * 64 * 2;       // color_net.2   16 x 64  (6 real outputs)
constexpr uint32_t W_S1 = W_C3 + 16 * 64 * 2;       // sigma_net.0   32 x 64
constexpr uint32_t W_S2 = W_S1 + 32 * 64 * 2;       // sigma_net.1   16 x 32  (1 real output)
constexpr uint32_t W_P1 = W_S2 + 16 * 32 * 2;       // specular_net.0 32 x 16 (6 real inputs)
constexpr uint32_t W_P2 = W_P1 + 32 * 16 * 2;       // specular_net.1 16 x 32 (3 real outputs)
constexpr uint32_t W_BYTES = W_P2 + 16 * 32 * 2;    // 25600

// flat fp32 parameter vector (reference nn.Linear layouts [out, in])
constexpr uint32_t P_S0 = 0, P_S1 = 608, P_C0 = 640, P_C1 = 2880, P_C2 = 6976, P_P0 = 7360, P_P1 = 7552, P_COUNT = 7648;

// enc tile column -> input index of the first-layer weights (-1: not an input of that net)
__host__ __device__ __forceinline__ int map_c1(uint32_t k) { return k < 3 ? (int)k : (k >= 19 && k < 51) ? (int)(k - 16) : -1; }
__host__ __device__ __forceinline__ int map_s1(uint32_t k) { return k < 19 ? (int)k : -1; }

// ---- small device helpers ---------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(tc::smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float round_h(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float sigmoid_h(float pre_acc) {          // torch.sigmoid on an fp16 tensor
    const float x = round_h(pre_acc);
    return round_h(1.0f / (1.0f + __expf(-x)));
}

// everyone: make generic smem writes visible to the tensor core, order TMEM reads, then barrier
__device__ __forceinline__ void sync_before_mma() {
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}

// accumulator row (NCOL fp32 columns of this thread's TMEM lane) -> optional ReLU / mask -> fp16 row of a
// 128-row chunk-major tile.  mask_tile != nullptr: zero where the fp16 activation stored there is <= 0.
// Latency-tuned: all TMEM loads of the row are issued before one wait, ReLU and the mask are applied on packed
// half2 values (HMNMX2 / HSETP-free multiply by __hgt2), half the instructions of the fp32 formulation and
// bit-identical results (rounding to fp16 commutes with max(.,0) and with zeroing).
template <int NCOL, bool RELU>
__device__ __forceinline__ void epi_store_row(uint32_t taddr, uint8_t* tile, uint32_t r, const uint8_t* mask_tile) {
    static_assert(NCOL == 32 || NCOL == 64, "row width");
    uint32_t raw[NCOL];
    {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[0]);
        tc::tmem_ld32(taddr, lo);
        if (NCOL == 64) {
            uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&raw[NCOL == 64 ? 32 : 0]);
            tc::tmem_ld32(taddr + 32, hi);
        }
    }
    uint4 msk[NCOL / 8];
    if (mask_tile) {
#pragma unroll
        for (int ch = 0; ch < NCOL / 8; ++ch) msk[ch] = *reinterpret_cast<const uint4*>(mask_tile + ch * kChunk + r * 16);
    }
    tc::tmem_ld_wait();
    const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
    for (int ch = 0; ch < NCOL / 8; ++ch) {
        __half2 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h[i] = __floats2half2_rn(__uint_as_float(raw[8 * ch + 2 * i]), __uint_as_float(raw[8 * ch + 2 * i + 1]));
            if (RELU) h[i] = __hmax2(h[i], zero2);
        }
        if (mask_tile) {
            const uint32_t mm[4] = {msk[ch].x, msk[ch].y, msk[ch].z, msk[ch].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __hmul2(h[i], __hgt2(*reinterpret_cast<const __half2*>(&mm[i]), zero2));
        }
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&h[0]); o.y = *reinterpret_cast<uint32_t*>(&h[1]);
        o.z = *reinterpret_cast<uint32_t*>(&h[2]); o.w = *reinterpret_cast<uint32_t*>(&h[3]);
        *reinterpret_cast<uint4*>(tile + ch * kChunk + r * 16) = o;
    }
}

__device__ __forceinline__ void store_chunk(uint8_t* tile, uint32_t chunk, uint32_t r, const float (&v)[8]) {
    uint4 o;
    o.x = pack2(v[0], v[1]); o.y = pack2(v[2], v[3]); o.z = pack2(v[4], v[5]); o.w = pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + chunk * kChunk + r * 16) = o;
}

__device__ __forceinline__ tc::Operand opK(const uint8_t* tile, uint32_t rows) { return tc::Operand{tc::smem_u32(tile), rows, false}; }
__device__ __forceinline__ tc::Operand opMN(const uint8_t* tile, uint32_t rows) { return tc::Operand{tc::smem_u32(tile), rows, true}; }

constexpr uint32_t B_W = 0;
constexpr uint32_t B_ACT = B_W + W_BYTES;            // activations, 34 chunks: A | H2 | H1 | S1 | P1 | As2
constexpr uint32_t A_A = 0, A_H2 = 16384, A_H1 = 32768, A_S1 = 49152, A_P1 = 57344, A_AS2 = 65536, ACT_BYTES = 69632;
constexpr uint32_t B_GRAD = B_ACT + ACT_BYTES;        // gradients: dH | dS1 | dP1 | dO | dOs | dO2
constexpr uint32_t G_DH = 0, G_DS1 = 16384, G_DP1 = 24576, G_DO = 32768, G_DOS = 36864, G_DO2 = 40960, GRAD_BYTES = 45056;
constexpr uint32_t B_BYTES = B_GRAD + GRAD_BYTES;    // 140288
// TMEM columns
constexpr uint32_t T_K0 = 0, T_K1 = 64, T_C1 = 128, T_C2 = 192, T_S1 = 256, T_P1 = 288, T_C3 = 320, T_S2 = 336, T_P2 = 352;

}  // namespace
}  // namespace n2m
