// mlp_tc.cu -- the three tiny MLPs of NeRFNetwork (nerf/network.py:66-189) forward and backward on
// Blackwell tensor cores: hand-written tcgen05.mma (kind::f16, fp16 operands, fp32 accumulation in
// TMEM), operands staged in shared memory as UMMA core-matrix tiles, first-layer activations brought
// in with one bulk async copy (TMA unit, cp.async.bulk) per 128-sample tile.
//
// Mapping: one CTA = 128 threads = one 128-sample tile = the M dimension of every forward / dgrad
// GEMM; tcgen05.ld in the 32x32b shape hands thread t exactly the accumulator row of sample t, so
// all epilogues (ReLU, sigmoid, exp, clamp, loss-side chain rule) are thread-per-sample with no
// shuffles.  Weight gradients are GEMMs whose reduction dimension is the SAMPLE index; the same
// shared-memory activation / gradient tiles are re-read MN-major for them (tc05.cuh) and the
// accumulators stay resident in TMEM across all tiles a persistent CTA processes, then are flushed
// once with atomics.  Numerics follow torch.autocast(fp16): operands and layer outputs rounded to
// fp16, fp32 accumulation; gradients are carried loss-scaled in fp16 like GradScaler does.
#include "n2m_common.cuh"
#include "tc05.cuh"
#include "mlp_common.cuh"
#include "../../include/n2m_b200_fused.h"

namespace n2m {
namespace {

__global__ void __launch_bounds__(256)
k_pack_weights(const float* __restrict__ P, uint8_t* __restrict__ wpack) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;      // one fp16 element
    if (i >= W_BYTES / 2) return;
    const uint32_t byte = i * 2;
    uint32_t base, rows; int which;
    if (byte < W_C2) { base = W_C1; rows = 64; which = 0; }
    else if (byte < W_C3) { base = W_C2; rows = 64; which = 1; }
    else if (byte < W_S1) { base = W_C3; rows = 16; which = 2; }
    else if (byte < W_S2) { base = W_S1; rows = 32; which = 3; }
    else if (byte < W_P1) { base = W_S2; rows = 16; which = 4; }
    else if (byte < W_P2) { base = W_P1; rows = 32; which = 5; }
    else { base = W_P2; rows = 16; which = 6; }
    // invert tile_off: byte offset -> (r, c)
    const uint32_t off = byte - base;
    const uint32_t chunk_bytes = rows * 16;
    const uint32_t ch = off / chunk_bytes, rem = off % chunk_bytes;
    const uint32_t r = rem / 16, c = ch * 8 + (rem % 16) / 2;
    float v = 0.f;
    switch (which) {
        case 0: { const int k = map_c1(c); if (k >= 0) v = P[P_C0 + r * 35 + k]; break; }
        case 1: v = P[P_C1 + r * 64 + c]; break;
        case 2: if (r < 6) v = P[P_C2 + r * 64 + c]; break;
        case 3: { const int k = map_s1(c); if (k >= 0) v = P[P_S0 + r * 19 + k]; break; }
        case 4: if (r < 1) v = P[P_S1 + c]; break;
        case 5: if (c < 6) v = P[P_P0 + r * 6 + c]; break;
        case 6: if (r < 3) v = P[P_P1 + r * 32 + c]; break;
    }
    *reinterpret_cast<__half*>(wpack + byte) = __float2half_rn(v);
}

// ================================================================================================
// forward
// ================================================================================================
constexpr uint32_t F_W = 0, F_A = F_W + W_BYTES, F_H = F_A + kTileBytes, F_S1 = F_H + kTileBytes,
                   F_P1 = F_S1 + 8192, F_AS2 = F_P1 + 8192, F_BYTES = F_AS2 + 4096;
// COMPACT layout (experimental, n2m_s0_set_mlp_fwd_compact): the specular hidden tile P1 aliases the sigma hidden tile S1, which is
// dead once sigma_net.1 has completed in round 2 (P1 is written in round 4, S1 again in round 1 of the next tile, after the
// wait on specular_net.1) -- 79 KB -> 71 KB of shared memory per CTA, i.e. three CTAs per SM instead of two (TMEM 3 x 128 columns).
constexpr uint32_t FC_AS2 = F_S1 + 8192, FC_BYTES = FC_AS2 + 4096;

template <bool COMPACT>
__global__ void __launch_bounds__(128)
k_mlp_fwd(n2m_s0_params p, const uint8_t* __restrict__ enc_tiles, const int32_t* __restrict__ counters,
          const uint8_t* __restrict__ wpack, float4* __restrict__ out, float* __restrict__ spec_sq_sum,
          uint32_t part, uint32_t nparts) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_mma, bar_tma;
    __shared__ uint32_t tmem_s;
    __shared__ float red[4];
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    // samples [lo, hi) of this part: tiles [t0, t1); the boundary tiles are also computed by the neighbouring parts,
    // every part writes only its own rows
    const PartRange pr = part_range(counters, part, nparts);
    const uint32_t t0 = pr.lo / kTile, t1 = (pr.hi + kTile - 1) / kTile;
    if (pr.hi <= pr.lo || t0 + blockIdx.x >= t1) return;

    if (tid == 0) { tc::mbar_init(&bar_mma, 1); tc::mbar_init(&bar_tma, 1); tc::mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 128);
    for (uint32_t i = tid; i < W_BYTES / 16; i += 128)
        reinterpret_cast<uint4*>(smem + F_W)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
    {   // second K chunk of the specular input tile is always zero
        *reinterpret_cast<uint4*>(smem + (COMPACT ? FC_AS2 : F_AS2) + kChunk + tid * 16) = make_uint4(0, 0, 0, 0);
    }
    sync_before_mma();
    const uint32_t tmem = tmem_s, D0 = tmem, D1 = tmem + 64;
    const uint32_t lane_t = (warp * 32u) << 16;
    uint32_t ph_mma = 0, ph_tma = 0;
    float spec_sq = 0.f;
    uint8_t* sW = smem + F_W; uint8_t* sA = smem + F_A; uint8_t* sH = smem + F_H;
    uint8_t* sS1 = smem + F_S1; uint8_t* sP1 = smem + (COMPACT ? F_S1 : F_P1); uint8_t* sAs2 = smem + (COMPACT ? FC_AS2 : F_AS2);
    // operand descriptors, built once (only the issuing thread uses them)
    const tc::OpDesc dA = tc::make_opdesc(opK(sA, 128)), dH = tc::make_opdesc(opK(sH, 128)), dS1 = tc::make_opdesc(opK(sS1, 128)),
                     dP1 = tc::make_opdesc(opK(sP1, 128)), dAs2 = tc::make_opdesc(opK(sAs2, 128));
    const tc::OpDesc wC1 = tc::make_opdesc(opK(sW + W_C1, 64)), wC2 = tc::make_opdesc(opK(sW + W_C2, 64)),
                     wC3 = tc::make_opdesc(opK(sW + W_C3, 16)), wS1 = tc::make_opdesc(opK(sW + W_S1, 32)),
                     wS2 = tc::make_opdesc(opK(sW + W_S2, 16)), wP1 = tc::make_opdesc(opK(sW + W_P1, 32)),
                     wP2 = tc::make_opdesc(opK(sW + W_P2, 16));

    for (uint32_t tile = t0 + blockIdx.x; tile < t1; tile += gridDim.x) {
        if (tid == 0) bulk_g2s(sA, enc_tiles + (size_t)tile * kTileBytes, kTileBytes, &bar_tma);
        tc::mbar_wait(&bar_tma, ph_tma); ph_tma ^= 1;

        // round 1: first layers of color_net and sigma_net
        if (tid == 0) {
            tc::gemm_issue_fast<64, 4, false, false>(D0, dA, wC1, false);
            tc::gemm_issue_fast<32, 4, false, false>(D1, dA, wS1, false);
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        epi_store_row<64, true>(D0 + lane_t, sH, tid, nullptr);
        epi_store_row<32, true>(D1 + lane_t, sS1, tid, nullptr);
        sync_before_mma();

        // round 2: color_net.1, sigma_net.1
        if (tid == 0) {
            tc::gemm_issue_fast<64, 4, false, false>(D0, dH, wC2, false);
            tc::gemm_issue_fast<16, 2, false, false>(D1, dS1, wS2, false);
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        float sigma;
        {
            float v[8];
            tc::tmem_ld8(D1 + lane_t, v);
            sigma = __expf(round_h(v[0]));                 // trunc_exp forward (activation.py:5-11)
        }
        epi_store_row<64, true>(D0 + lane_t, sH, tid, nullptr);       // H2 overwrites H1 (its reader has completed)
        sync_before_mma();

        // round 3: color_net.2
        if (tid == 0) {
            tc::gemm_issue_fast<16, 4, false, false>(D0, dH, wC3, false);
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        float feat[6];
        {
            float v[8];
            tc::tmem_ld8(D0 + lane_t, v);
#pragma unroll
            for (int i = 0; i < 6; ++i) feat[i] = sigmoid_h(v[i]);
        }
        float cr = feat[0], cg = feat[1], cb = feat[2];
        float sp[3] = {0.f, 0.f, 0.f};
        if (p.shading_full) {
            // specular input [dir(3), feat[3:6]]; dir sits in enc cols 51..53 = chunk 6, elements 3..5
            const uint4 dq = *reinterpret_cast<const uint4*>(sA + 6 * kChunk + tid * 16);
            const __half2 d01 = *reinterpret_cast<const __half2*>(&dq.y);     // elements 2,3
            const __half2 d23 = *reinterpret_cast<const __half2*>(&dq.z);     // elements 4,5
            const float in[8] = {__high2float(d01), __low2float(d23), __high2float(d23), feat[3], feat[4], feat[5], 0.f, 0.f};
            store_chunk(sAs2, 0, tid, in);
            sync_before_mma();
            if (tid == 0) {
                tc::gemm_issue_fast<32, 1, false, false>(D1, dAs2, wP1, false);
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
            epi_store_row<32, true>(D1 + lane_t, sP1, tid, nullptr);
            sync_before_mma();
            if (tid == 0) {
                tc::gemm_issue_fast<16, 2, false, false>(D0, dP1, wP2, false);
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
            float v[8];
            tc::tmem_ld8(D0 + lane_t, v);
#pragma unroll
            for (int i = 0; i < 3; ++i) sp[i] = sigmoid_h(v[i]);
            // color = (specular + diffuse).clamp(0, 1) on fp16 tensors (network.py:187)
            cr = fminf(fmaxf(round_h(sp[0] + cr), 0.f), 1.f);
            cg = fminf(fmaxf(round_h(sp[1] + cg), 0.f), 1.f);
            cb = fminf(fmaxf(round_h(sp[2] + cb), 0.f), 1.f);
        }
        const uint32_t j = tile * kTile + tid;
        if (j >= pr.lo && j < pr.hi) {
            out[j] = make_float4(sigma, cr, cg, cb);
            spec_sq += sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2];
        }
        sync_before_mma();          // all reads of this tile's smem / TMEM are done before the next bulk copy
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) spec_sq += __shfl_xor_sync(0xffffffffu, spec_sq, o);
    if ((tid & 31) == 0) red[warp] = spec_sq;
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0 && spec_sq_sum) atomicAdd(spec_sq_sum, red[0] + red[1] + red[2] + red[3]);
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ================================================================================================
// backward (forward recompute + dgrad + wgrad)
// ================================================================================================
__global__ void __launch_bounds__(128)
k_mlp_bwd(n2m_s0_params p, const uint8_t* __restrict__ enc_tiles, const float4* __restrict__ dout,
          const int32_t* __restrict__ counters, const uint8_t* __restrict__ wpack, uint8_t* __restrict__ denc_tiles,
          float* __restrict__ g_mlp, const float* __restrict__ loss_scale, uint32_t part, uint32_t nparts) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_mma, bar_tma;
    __shared__ uint32_t tmem_s;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    // rows outside [lo, hi) of a boundary tile belong to a neighbouring part: zero upstream gradient (so they add
    // nothing to the weight gradients; their activations are whatever finite values the tile holds) and no store
    const PartRange pr = part_range(counters, part, nparts);
    const uint32_t M = pr.M;
    const uint32_t t0 = pr.lo / kTile, t1 = (pr.hi + kTile - 1) / kTile;
    if (pr.hi <= pr.lo || t0 + blockIdx.x >= t1) return;

    if (tid == 0) { tc::mbar_init(&bar_mma, 1); tc::mbar_init(&bar_tma, 1); tc::mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_s, 512);
    for (uint32_t i = tid; i < W_BYTES / 16; i += 128)
        reinterpret_cast<uint4*>(smem + B_W)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
    uint8_t* sW = smem + B_W; uint8_t* act = smem + B_ACT; uint8_t* grd = smem + B_GRAD;
    uint8_t* sA = act + A_A; uint8_t* sH2 = act + A_H2; uint8_t* sH1 = act + A_H1; uint8_t* sS1 = act + A_S1;
    uint8_t* sP1 = act + A_P1; uint8_t* sAs2 = act + A_AS2;
    uint8_t* sdH = grd + G_DH; uint8_t* sdS1 = grd + G_DS1; uint8_t* sdP1 = grd + G_DP1; uint8_t* sdO = grd + G_DO;
    uint8_t* sdOs = grd + G_DOS; uint8_t* sdO2 = grd + G_DO2;
    {   // constant-zero parts of the narrow tiles (their second K chunk, and unused columns of the first)
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sAs2 + kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sdO + kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sdOs + kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sdO2 + kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sAs2 + tid * 16) = z;
        *reinterpret_cast<uint4*>(sP1 + tid * 16) = z; *reinterpret_cast<uint4*>(sP1 + kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sP1 + 2 * kChunk + tid * 16) = z; *reinterpret_cast<uint4*>(sP1 + 3 * kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sdP1 + tid * 16) = z; *reinterpret_cast<uint4*>(sdP1 + kChunk + tid * 16) = z;
        *reinterpret_cast<uint4*>(sdP1 + 2 * kChunk + tid * 16) = z; *reinterpret_cast<uint4*>(sdP1 + 3 * kChunk + tid * 16) = z;
    }
    sync_before_mma();
    const uint32_t tmem = tmem_s;
    const uint32_t lane_t = (warp * 32u) << 16;
    const uint32_t K0 = tmem + T_K0, K1 = tmem + T_K1;
    uint32_t ph_mma = 0, ph_tma = 0;
    const bool full = p.shading_full != 0;
    const float ls = loss_scale[0];
    const float spec_reg = (M > 0) ? 2.0f * p.lambda_specular / (float)M * ls : 0.f;   // d/dspec of lambda * mean_j sum_c spec^2
    bool first = true;                                  // wgrad accumulators are overwritten by the first tile

    const tc::Operand G1 = opMN(sA, 128);               // [A | H2]   as M = 128 input features
    const tc::Operand G2 = opMN(sH2, 128);              // [H2 | H1]
    const tc::Operand G3 = opMN(sS1, 128);              // [S1 | P1 | As2 | (don't care)]

    for (uint32_t tile = t0 + blockIdx.x; tile < t1; tile += gridDim.x) {
        if (tid == 0) bulk_g2s(sA, enc_tiles + (size_t)tile * kTileBytes, kTileBytes, &bar_tma);
        const uint32_t j = tile * kTile + tid;
        float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool own = j >= pr.lo && j < pr.hi;
        if (own) dv = dout[j];
        tc::mbar_wait(&bar_tma, ph_tma); ph_tma ^= 1;

        // ---------------- forward recompute ----------------
        if (tid == 0) {
            tc::gemm_issue(K0, opK(sA, 128), opK(sW + W_C1, 64), 128, 64, 64, false);
            tc::gemm_issue(K1, opK(sA, 128), opK(sW + W_S1, 32), 128, 32, 64, false);
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        epi_store_row<64, true>(K0 + lane_t, sH1, tid, nullptr);
        epi_store_row<32, true>(K1 + lane_t, sS1, tid, nullptr);
        sync_before_mma();
        if (tid == 0) {
            tc::gemm_issue(K0, opK(sH1, 128), opK(sW + W_C2, 64), 128, 64, 64, false);
            tc::gemm_issue(K1, opK(sS1, 128), opK(sW + W_S2, 16), 128, 16, 32, false);
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        float h_sig;
        { float v[8]; tc::tmem_ld8(K1 + lane_t, v); h_sig = round_h(v[0]); }
        epi_store_row<64, true>(K0 + lane_t, sH2, tid, nullptr);
        sync_before_mma();
        if (tid == 0) {
            tc::gemm_issue(K0, opK(sH2, 128), opK(sW + W_C3, 16), 128, 16, 64, false);
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        float feat[6];
        { float v[8]; tc::tmem_ld8(K0 + lane_t, v);
#pragma unroll
          for (int i = 0; i < 6; ++i) feat[i] = sigmoid_h(v[i]); }
        float sp[3] = {0.f, 0.f, 0.f};
        if (full) {
            const uint4 dq = *reinterpret_cast<const uint4*>(sA + 6 * kChunk + tid * 16);
            const __half2 d01 = *reinterpret_cast<const __half2*>(&dq.y);
            const __half2 d23 = *reinterpret_cast<const __half2*>(&dq.z);
            const float in[8] = {__high2float(d01), __low2float(d23), __high2float(d23), feat[3], feat[4], feat[5], 0.f, 0.f};
            store_chunk(sAs2, 0, tid, in);
            sync_before_mma();
            if (tid == 0) {
                tc::gemm_issue(K1, opK(sAs2, 128), opK(sW + W_P1, 32), 128, 32, 16, false);
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
            epi_store_row<32, true>(K1 + lane_t, sP1, tid, nullptr);
            sync_before_mma();
            if (tid == 0) {
                tc::gemm_issue(K0, opK(sP1, 128), opK(sW + W_P2, 16), 128, 16, 32, false);
                tc::mma_commit(&bar_mma);
            }
            tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
            float v[8];
            tc::tmem_ld8(K0 + lane_t, v);
#pragma unroll
            for (int i = 0; i < 3; ++i) sp[i] = sigmoid_h(v[i]);
        }

        // ---------------- output-side chain rule (thread-per-sample) ----------------
        float dfeat[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        {
            const float dcol[3] = {dv.y, dv.z, dv.w};
            float dO2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float g = dcol[c];
                if (full) {
                    const float cs = round_h(sp[c] + feat[c]);
                    if (!(cs >= 0.f && cs <= 1.f)) g = 0.f;            // clamp(0,1) backward
                    const float dsp = own ? g + spec_reg * sp[c] : 0.f;
                    dO2[c] = dsp * sp[c] * (1.0f - sp[c]);            // sigmoid backward
                }
                dfeat[c] = g;
            }
            float dOs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            dOs[0] = dv.x * __expf(fminf(fmaxf(h_sig, -15.f), 15.f));   // trunc_exp backward (activation.py:13-17)
            store_chunk(sdOs, 0, tid, dOs);
            if (full) store_chunk(sdO2, 0, tid, dO2);
        }
        sync_before_mma();

        // ---------------- B1: specular_net.1 / sigma_net.1 dgrad + their wgrads ----------------
        if (tid == 0) {
            tc::gemm_issue(K0, opK(sdOs, 128), opMN(sW + W_S2, 16), 128, 32, 16, false);            // dS1 (pre-mask)
            tc::gemm_issue(tmem + T_S2, G3, opMN(sdOs, 128), 128, 16, 128, !first);                  // rows 0..31: S1^T dOs
            if (full) {
                tc::gemm_issue(K1, opK(sdO2, 128), opMN(sW + W_P2, 16), 128, 32, 16, false);        // dP1 (pre-mask)
                tc::gemm_issue(tmem + T_P2, G3, opMN(sdO2, 128), 128, 16, 128, !first);              // rows 32..63: P1^T dO2
            }
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        epi_store_row<32, false>(K0 + lane_t, sdS1, tid, sS1);
        if (full) epi_store_row<32, false>(K1 + lane_t, sdP1, tid, sP1);
        sync_before_mma();

        // ---------------- B2: first-layer dgrads of sigma / specular nets + wgrads ----------------
        if (tid == 0) {
            tc::gemm_issue(K0, opK(sdS1, 128), opMN(sW + W_S1, 32), 128, 64, 32, false);            // d enc (sigma part) -> K0
            tc::gemm_issue(tmem + T_S1, G1, opMN(sdS1, 128), 128, 32, 128, !first);                  // rows 0..63: A^T dS1
            if (full) {
                tc::gemm_issue(K1, opK(sdP1, 128), opMN(sW + W_P1, 32), 128, 16, 32, false);        // d As2
                tc::gemm_issue(tmem + T_P1, G3, opMN(sdP1, 128), 128, 32, 128, !first);              // rows 64..79: As2^T dP1
            }
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        {
            if (full) {
                float v[8];
                tc::tmem_ld8(K1 + lane_t, v);
                dfeat[3] = v[3]; dfeat[4] = v[4]; dfeat[5] = v[5];
            }
            float dO[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 6; ++i) dO[i] = dfeat[i] * feat[i] * (1.0f - feat[i]);
            store_chunk(sdO, 0, tid, dO);
        }
        sync_before_mma();

        // ---------------- B3: color_net.2 dgrad + wgrad ----------------
        if (tid == 0) {
            tc::gemm_issue(K1, opK(sdO, 128), opMN(sW + W_C3, 16), 128, 64, 16, false);             // dH2 (pre-mask)
            tc::gemm_issue(tmem + T_C3, G2, opMN(sdO, 128), 128, 16, 128, !first);                   // rows 0..63: H2^T dO
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        epi_store_row<64, false>(K1 + lane_t, sdH, tid, sH2);
        sync_before_mma();

        // ---------------- B4: color_net.1 dgrad + wgrad ----------------
        if (tid == 0) {
            tc::gemm_issue(K1, opK(sdH, 128), opMN(sW + W_C2, 64), 128, 64, 64, false);             // dH1 (pre-mask)
            tc::gemm_issue(tmem + T_C2, G2, opMN(sdH, 128), 128, 64, 128, !first);                   // rows 64..127: H1^T dH2
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        epi_store_row<64, false>(K1 + lane_t, sdH, tid, sH1);
        sync_before_mma();

        // ---------------- B5: color_net.0 dgrad (accumulated onto the sigma part) + wgrad ----------------
        if (tid == 0) {
            tc::gemm_issue(K0, opK(sdH, 128), opMN(sW + W_C1, 64), 128, 64, 64, true);              // d enc += dH1 W_c1
            tc::gemm_issue(tmem + T_C1, G1, opMN(sdH, 128), 128, 64, 128, !first);                   // rows 0..63: A^T dH1
            tc::mma_commit(&bar_mma);
        }
        tc::mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; tc::fence_after_sync();
        {   // this sample's feature gradients -> its row of the denc tile image (global, coalesced per chunk)
            uint8_t* img = denc_tiles + (size_t)tile * kTileBytes + tid * 16;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float v[16];
                tc::tmem_ld16(K0 + lane_t + c0, v);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    uint4 o;
                    o.x = pack2(v[8 * q + 0], v[8 * q + 1]); o.y = pack2(v[8 * q + 2], v[8 * q + 3]);
                    o.z = pack2(v[8 * q + 4], v[8 * q + 5]); o.w = pack2(v[8 * q + 6], v[8 * q + 7]);
                    if (own || nparts == 1) *reinterpret_cast<uint4*>(img + (c0 / 8 + q) * kChunk) = o;
                }
            }
        }
        first = false;
        sync_before_mma();
    }

    // ---------------- flush the weight-gradient accumulators (one row of each per thread) ----------------
    {
        const uint32_t i = tid;            // accumulator row = input-feature index within its 128-wide group
        float v[16];
        // color_net.0: rows 0..63 = enc column i, cols = out o (64)
        {
            const int k = i < 64 ? map_c1(i) : -1;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                tc::tmem_ld16(tmem + T_C1 + lane_t + c0, v);
                if (k >= 0) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_C0 + (c0 + o) * 35 + k, v[o]);
                }
            }
        }
        // color_net.1: rows 64..127 = H1 feature i-64
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
            tc::tmem_ld16(tmem + T_C2 + lane_t + c0, v);
            if (i >= 64) {
#pragma unroll
                for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_C1 + (c0 + o) * 64 + (i - 64), v[o]);
            }
        }
        // color_net.2: rows 0..63 = H2 feature i, cols o < 6
        tc::tmem_ld16(tmem + T_C3 + lane_t, v);
        if (i < 64) {
#pragma unroll
            for (int o = 0; o < 6; ++o) atomicAdd(g_mlp + P_C2 + o * 64 + i, v[o]);
        }
        // sigma_net.0: rows 0..63 = enc column i, cols o < 32
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
            tc::tmem_ld16(tmem + T_S1 + lane_t + c0, v);
            const int k = i < 64 ? map_s1(i) : -1;
            if (k >= 0) {
#pragma unroll
                for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_S0 + (c0 + o) * 19 + k, v[o]);
            }
        }
        // sigma_net.1: rows 0..31 = S1 feature i, col 0
        tc::tmem_ld16(tmem + T_S2 + lane_t, v);
        if (i < 32) atomicAdd(g_mlp + P_S1 + i, v[0]);
        if (full) {
            // specular_net.1: rows 32..63 = P1 feature i-32, cols o < 3
            tc::tmem_ld16(tmem + T_P2 + lane_t, v);
            if (i >= 32 && i < 64) {
#pragma unroll
                for (int o = 0; o < 3; ++o) atomicAdd(g_mlp + P_P1 + o * 32 + (i - 32), v[o]);
            }
            // specular_net.0: rows 64..69 = As2 feature i-64 (< 6), cols o < 32
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                tc::tmem_ld16(tmem + T_P1 + lane_t + c0, v);
                if (i >= 64 && i < 70) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_P0 + (c0 + o) * 6 + (i - 64), v[o]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// ================================================================================================
// backward, pipelined: TWO 128-sample tiles in flight per CTA + a dedicated MMA-issuer warp
// ================================================================================================
// The single-tile kernel above is a chain of ten dependent [MMA -> commit -> wait -> TMEM load -> epilogue ->
// smem store -> fence -> barrier] rounds per tile, and TMEM (368 of 512 columns) allows only one such CTA per
// SM, so the tensor pipe idles during every epilogue and vice versa (profiles/r1_ncu_summary.md: 6 % issue
// utilisation).  Here warps 0-3 and 4-7 each own one tile (same thread-per-sample epilogues), warp 8 issues all
// tcgen05.mma for both: while one group runs an epilogue the other group's GEMMs execute.  The weight-gradient
// accumulators are shared by both groups (one issuing thread => one program order), working columns are
// per group: 2 x 128 + 240 = 496 TMEM columns.  Shared memory per group shrinks to 76 KB by reusing one 16 KB
// region for S1|P1 -> dS1|dP1 (written in place over the activations they mask) -> dH.
// MEASURED (profiles/r1_ncu_summary.md, r1f): 123 us against 104 us for the single-tile kernel -- the groups take turns
// on the tensor pipe, whose pace for these shapes is set by the shared-memory operand read (59-68 cycles per
// tcgen05.mma whatever N), so overlapping the (already short) epilogues buys nothing and the extra hand-overs cost.
// Kept behind n2m_s0_set_mlp_bwd_pipelined(1) with its phase profiler (n2m_s0_set_prof); the default is k_mlp_bwd.
constexpr uint32_t P_A = 0, P_H2 = 16384, P_H1 = 32768, P_Z = 49152, P_AS2 = 65536, P_DY = 69632, P_GRP = 77824;
constexpr uint32_t P_BYTES = W_BYTES + 2 * P_GRP + 4096;          // + tail so the last group's G3 operand stays in bounds
constexpr uint32_t Q_C1 = 256, Q_C2 = 320, Q_S1 = 384, Q_P1 = 416, Q_C3 = 448, Q_S2 = 464, Q_P2 = 480;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(tc::smem_u32(bar)) : "memory");
}
// group thread: operands written / TMEM reads finished -> hand the round to the issuer.  The 128 threads meet at
// a hardware named barrier and ONE of them arrives on the mbarrier (128 arrivals on one shared-memory word
// serialise: ~1k cycles per round when every thread arrived itself).
__device__ __forceinline__ void group_ready(uint64_t* bar, uint32_t g, uint32_t tg) {
    tc::fence_async_smem();
    tc::fence_before_sync();
    asm volatile("bar.sync %0, 128;" :: "r"(1 + g) : "memory");
    if (tg == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void group_wait(uint64_t* bar, uint32_t& ph) {
    tc::mbar_wait(bar, ph); ph ^= 1;
    tc::fence_after_sync();
}

// optional phase profiler (n2m_s0_set_prof): block 0 stamps clock64() at every hand-over of its first tile --
// slots 0..31 tile group 0 (thread 0), 32..63 the issuer's view of group 0, 64..95 tile group 1.  Null => off.
__device__ unsigned long long* g_prof = nullptr;
constexpr uint32_t kProfIter = 3;                 // which of block 0's tile iterations is stamped (steady state, not the cold first one)
#define PROF_STAMP() do { if (pf) { pf[ps++] = (unsigned long long)clock64(); } } while (0)

// ISSUERS = 1: warp 8 issues every tcgen05.mma (the measured variant).  ISSUERS = 2 (experimental, compiled only): warps 8 and 9
// issue for tile group 0 and 1 respectively -- two issuing threads reach 39-48 cycles per MMA aggregate against 59-68 for one
// (profiles/tcbench.py).  Both accumulate into the SAME weight-gradient columns, which are therefore zeroed up front with
// tcgen05.st and always accumulated into (no 'first MMA overwrites' flag that two threads would have to agree on).
template <int ISSUERS>
__global__ void __launch_bounds__(ISSUERS == 2 ? 320 : 288)
k_mlp_bwd2(n2m_s0_params p, const uint8_t* __restrict__ enc_tiles, const float4* __restrict__ dout,
           const int32_t* __restrict__ counters, const uint8_t* __restrict__ wpack, uint8_t* __restrict__ denc_tiles,
           float* __restrict__ g_mlp, const float* __restrict__ loss_scale, uint32_t part, uint32_t nparts) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_ready[2], bar_done[2], bar_tma[2];
    __shared__ uint32_t tmem_s;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const PartRange pr = part_range(counters, part, nparts);
    const uint32_t M = pr.M;
    const uint32_t t0 = pr.lo / kTile, ntiles = (pr.hi + kTile - 1) / kTile;       // tiles [t0, ntiles)
    if (pr.hi <= pr.lo || t0 + blockIdx.x * 2 >= ntiles) return;
    const bool full = p.shading_full != 0;

    if (tid == 0) {
        for (int g = 0; g < 2; ++g) { tc::mbar_init(&bar_ready[g], 1); tc::mbar_init(&bar_done[g], 1); tc::mbar_init(&bar_tma[g], 1); }
        tc::mbar_init_fence();
    }
    if (warp == 8) tc::tmem_alloc(&tmem_s, 512);
    for (uint32_t i = tid; i < W_BYTES / 16; i += (ISSUERS == 2 ? 320u : 288u))
        reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(wpack) + i);
    if (tid < 256) {    // constant-zero chunks of the narrow tiles, finite contents for the never-written ones
        uint8_t* grp = smem + W_BYTES + (tid >> 7) * P_GRP;
        const uint32_t r16 = (tid & 127) * 16;
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(grp + P_AS2 + r16) = z; *reinterpret_cast<uint4*>(grp + P_AS2 + kChunk + r16) = z;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(grp + P_DY + ch * kChunk + r16) = z;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(grp + P_Z + ch * kChunk + r16) = z;
    }
    if (tid < 256) *reinterpret_cast<uint4*>(smem + W_BYTES + 2 * P_GRP + tid * 16) = make_uint4(0, 0, 0, 0);
    sync_before_mma();
    const uint32_t tmem = tmem_s;
    uint8_t* sW = smem;
    if (ISSUERS == 2) {
        if (warp < 4) {
#pragma unroll
            for (uint32_t c = Q_C1; c < Q_P2 + 16; c += 16) tc::tmem_st16_zero(tmem + ((warp * 32u) << 16) + c);
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }

    if (ISSUERS == 2 ? warp >= 8 : warp == 8) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            uint32_t ph_ready[2] = {0, 0};
            uint32_t n_g[2];
            for (int g = 0; g < 2; ++g) {
                const uint32_t first_tile = t0 + blockIdx.x * 2 + g;
                n_g[g] = first_tile < ntiles ? (ntiles - first_tile + 2 * gridDim.x - 1) / (2 * gridDim.x) : 0;
            }
            bool f_c1 = ISSUERS == 2, f_c2 = ISSUERS == 2, f_s1 = ISSUERS == 2, f_p1 = ISSUERS == 2, f_c3 = ISSUERS == 2,
                 f_s2 = ISSUERS == 2, f_p2 = ISSUERS == 2;   // accumulator initialised? (two issuers: zeroed above)
            // all operand descriptors, built once per CTA (weights: K-major for forward, MN-major for dgrad)
            const tc::OpDesc wC1k = tc::make_opdesc(opK(sW + W_C1, 64)), wC2k = tc::make_opdesc(opK(sW + W_C2, 64)),
                             wC3k = tc::make_opdesc(opK(sW + W_C3, 16)), wS1k = tc::make_opdesc(opK(sW + W_S1, 32)),
                             wS2k = tc::make_opdesc(opK(sW + W_S2, 16)), wP1k = tc::make_opdesc(opK(sW + W_P1, 32)),
                             wP2k = tc::make_opdesc(opK(sW + W_P2, 16));
            const tc::OpDesc wC1m = tc::make_opdesc(opMN(sW + W_C1, 64)), wC2m = tc::make_opdesc(opMN(sW + W_C2, 64)),
                             wC3m = tc::make_opdesc(opMN(sW + W_C3, 16)), wS1m = tc::make_opdesc(opMN(sW + W_S1, 32)),
                             wS2m = tc::make_opdesc(opMN(sW + W_S2, 16)), wP1m = tc::make_opdesc(opMN(sW + W_P1, 32)),
                             wP2m = tc::make_opdesc(opMN(sW + W_P2, 16));
            struct GroupDesc {
                tc::OpDesc A, H2, H1, S1, P1, As2, dOs, dO2, Z;            // K-major activation / gradient tiles
                tc::OpDesc G1, G2, G3;                                     // MN-major 128-feature groups (wgrad A operands)
                tc::OpDesc S1m, P1m, Zm, dOsm, dO2m;                       // MN-major gradient tiles (wgrad B operands)
            } gd[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint8_t* grp = smem + W_BYTES + g * P_GRP;
                uint8_t* sZ = grp + P_Z; uint8_t* sDY = grp + P_DY;
                gd[g].A = tc::make_opdesc(opK(grp + P_A, 128));   gd[g].H2 = tc::make_opdesc(opK(grp + P_H2, 128));
                gd[g].H1 = tc::make_opdesc(opK(grp + P_H1, 128)); gd[g].S1 = tc::make_opdesc(opK(sZ, 128));
                gd[g].P1 = tc::make_opdesc(opK(sZ + 4 * kChunk, 128)); gd[g].As2 = tc::make_opdesc(opK(grp + P_AS2, 128));
                gd[g].dOs = tc::make_opdesc(opK(sDY, 128)); gd[g].dO2 = tc::make_opdesc(opK(sDY + 2 * kChunk, 128));
                gd[g].Z = tc::make_opdesc(opK(sZ, 128));
                gd[g].G1 = tc::make_opdesc(opMN(grp + P_A, 128)); gd[g].G2 = tc::make_opdesc(opMN(grp + P_H2, 128));
                gd[g].G3 = tc::make_opdesc(opMN(sZ, 128));
                gd[g].S1m = tc::make_opdesc(opMN(sZ, 128)); gd[g].P1m = tc::make_opdesc(opMN(sZ + 4 * kChunk, 128));
                gd[g].Zm = tc::make_opdesc(opMN(sZ, 128));
                gd[g].dOsm = tc::make_opdesc(opMN(sDY, 128)); gd[g].dO2m = tc::make_opdesc(opMN(sDY + 2 * kChunk, 128));
            }
            // The issuer serves whichever group has handed over its operands: the two tiles drift apart by about half a
            // round, so one group's GEMMs execute under the other group's epilogue (a fixed g0,g1,g0,... order would keep
            // both groups in lockstep: both in their epilogues, then both waiting on the tensor pipe).
            unsigned long long* pf0 = (blockIdx.x == 0 && warp == 8) ? g_prof : nullptr;
            if (pf0) pf0 += 32;
            uint32_t ps = 0;
            uint32_t rd[2] = {0, 0}, itg[2] = {0, 0};
            bool act[2] = {n_g[0] > 0 && (ISSUERS == 1 || warp == 8), n_g[1] > 0 && (ISSUERS == 1 || warp == 9)};
            uint32_t spins = 0;
            while (act[0] || act[1]) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        if (!act[g]) continue;
                        if (!tc::mbar_test_wait(&bar_ready[g], ph_ready[g])) {
                            if (++spins > (1u << 28)) __trap();
                            continue;
                        }
                        spins = 0;
                        ph_ready[g] ^= 1;
                        tc::fence_after_sync();
                        const int round = (int)rd[g];
                        const GroupDesc& q = gd[g];
                        const uint32_t K0 = tmem + g * 128, K1 = K0 + 64;
                        unsigned long long* pf = (itg[g] == kProfIter && g == 0) ? pf0 : nullptr;
                        PROF_STAMP();
                        switch (round) {
                            case 0:   // R1: first layers
                                tc::gemm_issue_fast<64, 4, false, false>(K0, q.A, wC1k, false);
                                tc::gemm_issue_fast<32, 4, false, false>(K1, q.A, wS1k, false);
                                break;
                            case 1:   // R2
                                tc::gemm_issue_fast<64, 4, false, false>(K0, q.H1, wC2k, false);
                                tc::gemm_issue_fast<16, 2, false, false>(K1, q.S1, wS2k, false);
                                break;
                            case 2:   // R3
                                tc::gemm_issue_fast<16, 4, false, false>(K0, q.H2, wC3k, false);
                                break;
                            case 3:   // R4
                                tc::gemm_issue_fast<32, 1, false, false>(K1, q.As2, wP1k, false);
                                break;
                            case 4:   // R5
                                tc::gemm_issue_fast<16, 2, false, false>(K0, q.P1, wP2k, false);
                                break;
                            case 5:   // B1: last-layer dgrads of sigma / specular nets + their wgrads
                                tc::gemm_issue_fast<32, 1, false, true>(K0, q.dOs, wS2m, false);
                                tc::gemm_issue_fast<16, 8, true, true>(tmem + Q_S2, q.G3, q.dOsm, f_s2); f_s2 = true;
                                if (full) {
                                    tc::gemm_issue_fast<32, 1, false, true>(K1, q.dO2, wP2m, false);
                                    tc::gemm_issue_fast<16, 8, true, true>(tmem + Q_P2, q.G3, q.dO2m, f_p2); f_p2 = true;
                                }
                                break;
                            case 6:   // B2 (S1|P1 now hold dS1|dP1)
                                tc::gemm_issue_fast<64, 2, false, true>(K0, q.S1, wS1m, false);
                                tc::gemm_issue_fast<32, 8, true, true>(tmem + Q_S1, q.G1, q.S1m, f_s1); f_s1 = true;
                                if (full) {
                                    tc::gemm_issue_fast<16, 2, false, true>(K1, q.P1, wP1m, false);
                                    tc::gemm_issue_fast<32, 8, true, true>(tmem + Q_P1, q.G3, q.P1m, f_p1); f_p1 = true;
                                }
                                break;
                            case 7:   // B3 (dO aliases dOs)
                                tc::gemm_issue_fast<64, 1, false, true>(K1, q.dOs, wC3m, false);
                                tc::gemm_issue_fast<16, 8, true, true>(tmem + Q_C3, q.G2, q.dOsm, f_c3); f_c3 = true;
                                break;
                            case 8:   // B4 (Z holds dH2)
                                tc::gemm_issue_fast<64, 4, false, true>(K1, q.Z, wC2m, false);
                                tc::gemm_issue_fast<64, 8, true, true>(tmem + Q_C2, q.G2, q.Zm, f_c2); f_c2 = true;
                                break;
                            case 9:   // B5 (Z holds dH1)
                                tc::gemm_issue_fast<64, 4, false, true>(K0, q.Z, wC1m, true);
                                tc::gemm_issue_fast<64, 8, true, true>(tmem + Q_C1, q.G1, q.Zm, f_c1); f_c1 = true;
                                break;
                        }
                        tc::mma_commit(&bar_done[g]);
                        PROF_STAMP();
                        ++rd[g];
                        if (!full && rd[g] == 3) rd[g] = 5;             // no specular rounds in 'diffuse' shading
                        if (rd[g] == 10) { rd[g] = 0; if (++itg[g] == n_g[g]) act[g] = false; }
                    }
            }
        }
    } else {
        // ------------------------------ tile groups ------------------------------
        const uint32_t g = warp >> 2, tg = tid & 127;
        uint8_t* grp = smem + W_BYTES + g * P_GRP;
        uint8_t* sA = grp + P_A; uint8_t* sH2 = grp + P_H2; uint8_t* sH1 = grp + P_H1; uint8_t* sZ = grp + P_Z;
        uint8_t* sAs2 = grp + P_AS2; uint8_t* sDY = grp + P_DY;
        uint8_t* sS1 = sZ; uint8_t* sP1 = sZ + 4 * kChunk;
        uint8_t* sdOs = sDY; uint8_t* sdO2 = sDY + 2 * kChunk; uint8_t* sdO = sDY;
        const uint32_t lane_t = ((warp & 3) * 32u) << 16;
        const uint32_t K0 = tmem + g * 128 + lane_t, K1 = K0 + 64;
        uint32_t ph_done = 0, ph_tma = 0;
        const float ls = loss_scale[0];
        const float spec_reg = (M > 0) ? 2.0f * p.lambda_specular / (float)M * ls : 0.f;

        unsigned long long* pfb = (blockIdx.x == 0 && tg == 0 && g_prof) ? g_prof + g * 64 : nullptr;
        uint32_t ps = 0;
        for (uint32_t tile = t0 + blockIdx.x * 2 + g; tile < ntiles; tile += 2 * gridDim.x) {
            unsigned long long* pf = (tile == t0 + blockIdx.x * 2 + g + kProfIter * 2 * gridDim.x) ? pfb : nullptr;
            PROF_STAMP();
            if (tg == 0) bulk_g2s(sA, enc_tiles + (size_t)tile * kTileBytes, kTileBytes, &bar_tma[g]);
            const uint32_t j = tile * kTile + tg;
            float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool own = j >= pr.lo && j < pr.hi;
            if (own) dv = dout[j];
            tc::mbar_wait(&bar_tma[g], ph_tma); ph_tma ^= 1;
            PROF_STAMP();
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);                                   // R1 may start
            // ---- forward recompute ----
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            epi_store_row<64, true>(K0, sH1, tg, nullptr);
            epi_store_row<32, true>(K1, sS1, tg, nullptr);
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            float h_sig;
            { float v[8]; tc::tmem_ld8(K1, v); h_sig = round_h(v[0]); }
            epi_store_row<64, true>(K0, sH2, tg, nullptr);
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            float feat[6];
            { float v[8]; tc::tmem_ld8(K0, v);
#pragma unroll
              for (int i = 0; i < 6; ++i) feat[i] = sigmoid_h(v[i]); }
            float sp[3] = {0.f, 0.f, 0.f};
            if (full) {
                const uint4 dq = *reinterpret_cast<const uint4*>(sA + 6 * kChunk + tg * 16);
                const __half2 d01 = *reinterpret_cast<const __half2*>(&dq.y);
                const __half2 d23 = *reinterpret_cast<const __half2*>(&dq.z);
                const float in[8] = {__high2float(d01), __low2float(d23), __high2float(d23), feat[3], feat[4], feat[5], 0.f, 0.f};
                store_chunk(sAs2, 0, tg, in);
                PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
                group_wait(&bar_done[g], ph_done); PROF_STAMP();
                epi_store_row<32, true>(K1, sP1, tg, nullptr);
                PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
                group_wait(&bar_done[g], ph_done); PROF_STAMP();
                float v[8];
                tc::tmem_ld8(K0, v);
#pragma unroll
                for (int i = 0; i < 3; ++i) sp[i] = sigmoid_h(v[i]);
            }
            // ---- output-side chain rule ----
            float dfeat[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            {
                const float dcol[3] = {dv.y, dv.z, dv.w};
                float dO2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float gg = dcol[c];
                    if (full) {
                        const float cs = round_h(sp[c] + feat[c]);
                        if (!(cs >= 0.f && cs <= 1.f)) gg = 0.f;
                        const float dsp = own ? gg + spec_reg * sp[c] : 0.f;
                        dO2[c] = dsp * sp[c] * (1.0f - sp[c]);
                    }
                    dfeat[c] = gg;
                }
                float dOs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                dOs[0] = dv.x * __expf(fminf(fmaxf(h_sig, -15.f), 15.f));
                store_chunk(sdOs, 0, tg, dOs);
                if (full) store_chunk(sdO2, 0, tg, dO2);
            }
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            // ---- B1 ----
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            epi_store_row<32, false>(K0, sS1, tg, sS1);                   // dS1 written over S1 (own row: read mask, then write)
            if (full) epi_store_row<32, false>(K1, sP1, tg, sP1);
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            // ---- B2 ----
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            {
                if (full) {
                    float v[8];
                    tc::tmem_ld8(K1, v);
                    dfeat[3] = v[3]; dfeat[4] = v[4]; dfeat[5] = v[5];
                }
                float dO[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 6; ++i) dO[i] = dfeat[i] * feat[i] * (1.0f - feat[i]);
                store_chunk(sdO, 0, tg, dO);
            }
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            // ---- B3 ----
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            epi_store_row<64, false>(K1, sZ, tg, sH2);
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            // ---- B4 ----
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            epi_store_row<64, false>(K1, sZ, tg, sH1);
            PROF_STAMP(); group_ready(&bar_ready[g], g, tg);
            // ---- B5 ----
            group_wait(&bar_done[g], ph_done); PROF_STAMP();
            {
                uint8_t* img = denc_tiles + (size_t)tile * kTileBytes + tg * 16;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float v[16];
                    tc::tmem_ld16(K0 + c0, v);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint4 o;
                        o.x = pack2(v[8 * q + 0], v[8 * q + 1]); o.y = pack2(v[8 * q + 2], v[8 * q + 3]);
                        o.z = pack2(v[8 * q + 4], v[8 * q + 5]); o.w = pack2(v[8 * q + 6], v[8 * q + 7]);
                        if (own || nparts == 1) *reinterpret_cast<uint4*>(img + (c0 / 8 + q) * kChunk) = o;
                    }
                }
            }
            PROF_STAMP();
            // the group's smem / TMEM working columns are free again: the next tile's bulk copy and R1 may proceed
            // (every MMA of this tile has completed, all TMEM loads above have been waited for)
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    // ---------------- flush the shared weight-gradient accumulators (warps 0-3: one row each) ----------------
    if (warp < 4) {
        const uint32_t i = tid;
        const uint32_t lane_t = (warp * 32u) << 16;
        float v[16];
        {
            const int k = i < 64 ? map_c1(i) : -1;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                tc::tmem_ld16(tmem + Q_C1 + lane_t + c0, v);
                if (k >= 0) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_C0 + (c0 + o) * 35 + k, v[o]);
                }
            }
        }
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
            tc::tmem_ld16(tmem + Q_C2 + lane_t + c0, v);
            if (i >= 64) {
#pragma unroll
                for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_C1 + (c0 + o) * 64 + (i - 64), v[o]);
            }
        }
        tc::tmem_ld16(tmem + Q_C3 + lane_t, v);
        if (i < 64) {
#pragma unroll
            for (int o = 0; o < 6; ++o) atomicAdd(g_mlp + P_C2 + o * 64 + i, v[o]);
        }
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
            tc::tmem_ld16(tmem + Q_S1 + lane_t + c0, v);
            const int k = i < 64 ? map_s1(i) : -1;
            if (k >= 0) {
#pragma unroll
                for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_S0 + (c0 + o) * 19 + k, v[o]);
            }
        }
        tc::tmem_ld16(tmem + Q_S2 + lane_t, v);
        if (i < 32) atomicAdd(g_mlp + P_S1 + i, v[0]);
        if (full) {
            tc::tmem_ld16(tmem + Q_P2 + lane_t, v);
            if (i >= 32 && i < 64) {
#pragma unroll
                for (int o = 0; o < 3; ++o) atomicAdd(g_mlp + P_P1 + o * 32 + (i - 32), v[o]);
            }
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 16) {
                tc::tmem_ld16(tmem + Q_P1 + lane_t + c0, v);
                if (i >= 64 && i < 70) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) atomicAdd(g_mlp + P_P0 + (c0 + o) * 6 + (i - 64), v[o]);
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace n2m

using namespace n2m;

extern "C" {

/* debug: device buffer of >= 128 uint64 that block 0 of the pipelined MLP backward stamps with clock64(); NULL = off */
int n2m_s0_set_prof(void* buf) {
    unsigned long long* b = static_cast<unsigned long long*>(buf);
    cudaError_t e = cudaMemcpyToSymbol(g_prof, &b, sizeof(b));
    if (e != cudaSuccess) return fail("s0_set_prof", cudaGetErrorString(e));
    return 0;
}


uint32_t n2m_s0_wpack_bytes(void) { return W_BYTES; }
uint32_t n2m_s0_mlp_param_count(void) { return P_COUNT; }

int n2m_s0_pack_weights(const float* mlp_params, void* wpack, n2m_stream_t stream) {
    N2M_REQUIRE(mlp_params && wpack, "s0_pack_weights", "null pointer");
    k_pack_weights<<<div_up(W_BYTES / 2, 256u), 256, 0, as_stream(stream)>>>(mlp_params, static_cast<uint8_t*>(wpack));
    return check_launch("s0_pack_weights");
}

static int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

static bool g_fwd_compact = false;
/* experimental tuning hook (compiled, not yet measured): 1 = MLP forward with the compact shared-memory layout, three CTAs per SM */
int n2m_s0_set_mlp_fwd_compact(int on) { g_fwd_compact = on != 0; return 0; }

static int g_bwd_issuers = 1;
/* experimental tuning hook (compiled, not yet measured): issuing warps of the two-tile MLP backward, 1 (default) or 2 */
int n2m_s0_set_mlp_bwd_issuers(int n) { g_bwd_issuers = n == 2 ? 2 : 1; return 0; }
static bool g_bwd_pipelined = false;
/* 0 = single-tile backward kernel (default: measured faster, and it leaves room on the SM for a co-resident gather /
 * scatter kernel), 1 = two-tile pipelined kernel with issuer warp */
int n2m_s0_set_mlp_bwd_pipelined(int on) { g_bwd_pipelined = on != 0; return 0; }

/* one-time function attributes (dynamic shared memory opt-in); safe to call repeatedly */
int n2m_s0_init(void) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_fwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_fwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FC_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_bwd2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_bwd2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_BYTES);
    if (e != cudaSuccess) return fail("s0_init", cudaGetErrorString(e));
    num_sms();
    return 0;
}

int n2m_s0_mlp_fwd_part(const n2m_s0_params* p, const void* enc_tiles, const int32_t* counters, uint32_t Mcap, const void* wpack,
                        void* out, float* spec_sq_sum, uint32_t part, uint32_t nparts, n2m_stream_t stream) {
    N2M_REQUIRE(p && enc_tiles && counters && wpack && out, "s0_mlp_fwd", "null pointer");
    N2M_REQUIRE(Mcap % kTile == 0 && Mcap > 0, "s0_mlp_fwd", "Mcap must be a positive multiple of 128");
    N2M_REQUIRE(valid_parts(part, nparts), "s0_mlp_fwd", "nparts must be 1, 2, 4 or 8 and part < nparts");
    if (g_fwd_compact) {
        const uint32_t grid3 = min(Mcap / kTile, (uint32_t)(3 * num_sms()));
        k_mlp_fwd<true><<<grid3, 128, FC_BYTES, as_stream(stream)>>>(*p, static_cast<const uint8_t*>(enc_tiles), counters,
                                                                     static_cast<const uint8_t*>(wpack), static_cast<float4*>(out),
                                                                     spec_sq_sum, part, nparts);
        return check_launch("s0_mlp_fwd(compact)");
    }
    const uint32_t grid = min(Mcap / kTile, (uint32_t)(2 * num_sms()));
    k_mlp_fwd<false><<<grid, 128, F_BYTES, as_stream(stream)>>>(*p, static_cast<const uint8_t*>(enc_tiles), counters,
                                                                static_cast<const uint8_t*>(wpack), static_cast<float4*>(out), spec_sq_sum,
                                                                part, nparts);
    return check_launch("s0_mlp_fwd");
}

int n2m_s0_mlp_fwd(const n2m_s0_params* p, const void* enc_tiles, const int32_t* counters, uint32_t Mcap, const void* wpack,
                   void* out, float* spec_sq_sum, n2m_stream_t stream) {
    return n2m_s0_mlp_fwd_part(p, enc_tiles, counters, Mcap, wpack, out, spec_sq_sum, 0, 1, stream);
}

int n2m_s0_mlp_bwd_part(const n2m_s0_params* p, const void* enc_tiles, const void* dout, const int32_t* counters, uint32_t Mcap,
                        const void* wpack, void* denc_tiles, float* g_mlp, const float* loss_scale, uint32_t part, uint32_t nparts,
                        n2m_stream_t stream) {
    N2M_REQUIRE(p && enc_tiles && dout && counters && wpack && denc_tiles && g_mlp && loss_scale, "s0_mlp_bwd", "null pointer");
    N2M_REQUIRE(Mcap % kTile == 0 && Mcap > 0, "s0_mlp_bwd", "Mcap must be a positive multiple of 128");
    N2M_REQUIRE(valid_parts(part, nparts), "s0_mlp_bwd", "nparts must be 1, 2, 4 or 8 and part < nparts");
    if (g_bwd_pipelined) {
        const uint32_t grid2 = min((Mcap / kTile + 1) / 2, (uint32_t)num_sms());
        if (g_bwd_issuers == 2)
            k_mlp_bwd2<2><<<grid2, 320, P_BYTES, as_stream(stream)>>>(*p, static_cast<const uint8_t*>(enc_tiles), static_cast<const float4*>(dout),
                                                                      counters, static_cast<const uint8_t*>(wpack), static_cast<uint8_t*>(denc_tiles),
                                                                      g_mlp, loss_scale, part, nparts);
        else
        k_mlp_bwd2<1><<<grid2, 288, P_BYTES, as_stream(stream)>>>(*p, static_cast<const uint8_t*>(enc_tiles), static_cast<const float4*>(dout),
                                                               counters, static_cast<const uint8_t*>(wpack), static_cast<uint8_t*>(denc_tiles),
                                                               g_mlp, loss_scale, part, nparts);
        return check_launch("s0_mlp_bwd(pipelined)");
    }
    const uint32_t grid = min(Mcap / kTile, (uint32_t)num_sms());
    k_mlp_bwd<<<grid, 128, B_BYTES, as_stream(stream)>>>(*p, static_cast<const uint8_t*>(enc_tiles), static_cast<const float4*>(dout),
                                                         counters, static_cast<const uint8_t*>(wpack), static_cast<uint8_t*>(denc_tiles),
                                                         g_mlp, loss_scale, part, nparts);
    return check_launch("s0_mlp_bwd");
}

int n2m_s0_mlp_bwd(const n2m_s0_params* p, const void* enc_tiles, const void* dout, const int32_t* counters, uint32_t Mcap,
                   const void* wpack, void* denc_tiles, float* g_mlp, const float* loss_scale, n2m_stream_t stream) {
    return n2m_s0_mlp_bwd_part(p, enc_tiles, dout, counters, Mcap, wpack, denc_tiles, g_mlp, loss_scale, 0, 1, stream);
}

}  // extern "C"
