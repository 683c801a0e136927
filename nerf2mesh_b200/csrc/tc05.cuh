// tc05.cuh -- minimal hand-written tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX).
//
// Operand tiles live in shared memory in the no-swizzle ("interleave") canonical layout of the
// UMMA shared-memory descriptor: 8x8 fp16 core matrices of 128 contiguous bytes (8 rows x 16 B).
// All tiles in this library are stored CHUNK-MAJOR:
//      byte(r, c) = (c / 8) * chunk_bytes + (r / 8) * 128 + (r % 8) * 16 + (c % 8) * 2,
//      chunk_bytes = rows / 8 * 128
// so the same physical tile can be fed to the tensor core either
//   * K-major   (rows = M/N index, cols = K):  SBO = 128, LBO = chunk_bytes, or
//   * MN-major  (cols = M/N index, rows = K):  SBO = chunk_bytes, LBO = 128
// (descriptor field meaning per cute/arch/mma_sm100_desc.hpp + mma_traits_sm100.hpp:194,242).
// That is what lets one activation tile serve forward/dgrad (K-major) and wgrad (MN-major), and
// one weight tile serve forward (K-major B) and dgrad (MN-major B), without any transposition.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (r, c) in a chunk-major tile with `rows` rows (fp16)
__host__ __device__ __forceinline__ uint32_t tile_off(uint32_t r, uint32_t c, uint32_t rows) {
    return (c >> 3) * (rows << 4) + (r >> 3) * 128u + (r & 7u) * 16u + (c & 7u) * 2u;
}

// ---- descriptors --------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                       // descriptor version 1 (Blackwell)
    return d;                              // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

// kind::f16, fp16 x fp16 -> fp32; M in {64,128}, N multiple of 16 (M=128) in [16,256]
__host__ __device__ constexpr uint32_t instr_desc(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
    return (1u << 4)                       // c_format = F32
         | (0u << 7) | (0u << 10)          // a_format = b_format = F16
         | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16)
         | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- TMEM ---------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA ----------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a hardware-defined time; a poller of several barriers must not)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > 4000000u) __trap();
    }
}

// ---- registers -> TMEM: zero 16 consecutive 32-bit columns of the calling thread's lane (same lane rule as the loads below) ----
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 :: "r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM -> registers: each thread reads N consecutive fp32 columns of ITS lane (row) ------------
// warp w may only touch lanes [32*(w%4), 32*(w%4)+32): taddr = ((lane_base) << 16) | column
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 columns in one instruction (one wait instead of two): the epilogues are latency bound on LDTM round trips
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- helpers for issuing a whole GEMM over K ------------------------------------------------------
// A tile: rows_a rows (chunk-major), B tile: rows_b rows.  K-major operand: M/N = rows, K = cols.
// MN-major operand: M/N = cols, K = rows.  k_total multiple of 16.
struct Operand {
    uint32_t saddr;      // shared address of the tile element (0, first column used)
    uint32_t rows;       // rows of the physical tile (defines chunk_bytes)
    bool mn_major;
    __device__ __forceinline__ uint64_t desc(uint32_t k0) const {
        const uint32_t chunk = rows << 4;
        if (!mn_major)   // K along columns: 16 k = 2 chunks
            return smem_desc(saddr + (k0 >> 3) * chunk, /*lbo=*/chunk, /*sbo=*/128u);
        // K along rows: 16 k = 2 groups of 8 rows = 256 B
        return smem_desc(saddr + (k0 >> 3) * 128u, /*lbo=*/128u, /*sbo=*/chunk);
    }
};

__device__ __forceinline__ void gemm_issue(uint32_t d_tmem, const Operand& a, const Operand& b, uint32_t M, uint32_t N,
                                           uint32_t k_total, bool accumulate_first) {
    const uint32_t idesc = instr_desc(M, N, a.mn_major, b.mn_major);
    for (uint32_t k = 0; k < k_total; k += 16)
        mma_f16(d_tmem, a.desc(k), b.desc(k), idesc, accumulate_first || k > 0);
}

// ---- cheap issue path ------------------------------------------------------------------------------
// A single thread issues every tcgen05.mma of a CTA, so descriptor arithmetic is on the critical path
// (profiles/r1_ncu_summary.md: ~125 cycles per MMA with smem_desc() rebuilt each time).  The descriptor of k-step
// `s` differs from that of k-step 0 only in the 14-bit start-address field (low word), by a constant:
//   K-major : 16 k = 2 chunks          -> += 2 * chunk_bytes / 16 = rows * 2
//   MN-major: 16 k = 2 groups of 8 rows -> += 256 / 16 = 16
// so operands are described ONCE (outside the tile loop) and the unrolled issue loop is one IADD per operand.
struct OpDesc {
    uint32_t lo, hi, step;       // 64-bit descriptor of k-step 0 (split) and the per-k-step increment of `lo`
};
__device__ __forceinline__ OpDesc make_opdesc(const Operand& o) {
    const uint64_t d = o.desc(0);
    OpDesc r;
    r.lo = static_cast<uint32_t>(d);
    r.hi = static_cast<uint32_t>(d >> 32);
    r.step = o.mn_major ? 16u : (o.rows * 2u);
    return r;
}
__device__ __forceinline__ uint64_t desc_at(const OpDesc& o, uint32_t s) {
    return (static_cast<uint64_t>(o.hi) << 32) | static_cast<uint64_t>(o.lo + s * o.step);
}
// D[tmem] (+)= A * B over KSTEPS k-steps of 16; shapes and majors are compile-time so the loop unrolls fully
template <uint32_t N, uint32_t KSTEPS, bool A_MN, bool B_MN>
__device__ __forceinline__ void gemm_issue_fast(uint32_t d_tmem, const OpDesc& a, const OpDesc& b, bool accumulate_first) {
    constexpr uint32_t idesc = instr_desc(128, N, A_MN, B_MN);
#pragma unroll
    for (uint32_t s = 0; s < KSTEPS; ++s)
        mma_f16(d_tmem, desc_at(a, s), desc_at(b, s), idesc, accumulate_first || s > 0);
}

}  // namespace tc
