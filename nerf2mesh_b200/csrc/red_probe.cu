// red_probe.cu -- micro-benchmark of spread global reductions (fire-and-forget RED) by payload: what does one lane of
// red.global.add cost when every lane of a warp hits a different, random row of a table that is (a) L2-resident or
// (b) larger than L2?  Decides how the hash-gradient scatter should carry its payload (profiles/redbench.py).
#include "n2m_common.cuh"
#include "../../include/n2m_b200.h"
#include <cuda_fp16.h>

namespace n2m {
namespace {

__device__ __forceinline__ uint32_t mix(uint32_t x) {          // cheap integer hash -> pseudo-random row
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE 0: red.f32 (4 B)   1: red.v2.f32 (8 B)   2: red.v4.f32 (16 B)   3: red.v2.f16x2 (8 B = 4 halves)   4: red.v4.f16x2 (16 B)
template <int MODE>
__global__ void __launch_bounds__(256)
k_red_bench(float* __restrict__ table, uint32_t rows, uint32_t per_thread, uint32_t seed) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h = mix(tid * 2654435761u + seed);
    for (uint32_t i = 0; i < per_thread; ++i) {
        h = mix(h + i);
        const uint32_t row = h % rows;
        float* p = table + (size_t)row * 4;                      // 16-byte rows in every mode (same address pattern)
        if (MODE == 0) asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(1.0f) : "memory");
        if (MODE == 1) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(1.0f), "f"(2.0f) : "memory");
        if (MODE == 2) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(0.0f) : "memory");
        if (MODE == 3) asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" :: "l"(p), "r"(0x3c003c00u), "r"(0x3c003c00u) : "memory");
        if (MODE == 4) asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u), "r"(0x3c003c00u) : "memory");
    }
}

}  // namespace
}  // namespace n2m

using namespace n2m;

/* table: >= rows * 16 bytes; launches `blocks` x 256 threads, each issuing `per_thread` REDs to pseudo-random 16-byte rows */
extern "C" int n2m_red_bench(int mode, float* table, uint32_t rows, uint32_t blocks, uint32_t per_thread, uint32_t seed,
                             n2m_stream_t stream) {
    N2M_REQUIRE(table && rows > 0 && blocks > 0, "red_bench", "bad arguments");
    cudaStream_t st = as_stream(stream);
    switch (mode) {
        case 0: k_red_bench<0><<<blocks, 256, 0, st>>>(table, rows, per_thread, seed); break;
        case 1: k_red_bench<1><<<blocks, 256, 0, st>>>(table, rows, per_thread, seed); break;
        case 2: k_red_bench<2><<<blocks, 256, 0, st>>>(table, rows, per_thread, seed); break;
        case 3: k_red_bench<3><<<blocks, 256, 0, st>>>(table, rows, per_thread, seed); break;
        case 4: k_red_bench<4><<<blocks, 256, 0, st>>>(table, rows, per_thread, seed); break;
        default: return fail("red_bench", "mode must be 0..4");
    }
    return check_launch("red_bench");
}
