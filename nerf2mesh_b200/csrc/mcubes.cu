// mcubes.cu -- iso-surface extraction from the density volume: the first step of the stage-0 -> stage-1 hand-off
// (NeRFRenderer.export_stage0, nerf/renderer.py:526-529: `vertices, triangles = mcubes.marching_cubes(sigmas, density_thresh)` on a
// 512^3 volume that the reference first copies to the host; PyMCubes is a third-party CPU library, not vendored, version unpinned).
// Here the volume stays on the device (Stage0Trainer.density_volume) and the mesh is extracted by two kernels around one prefix sum:
//   k_mc_count : one thread per GRID POINT p = (x, y, z): the number of iso-crossings on the three edges the point owns (+x, +y, +z) and
//                the number of triangles of the cell whose minimum corner it is (case table of nerf2mesh_b200/mc_table.py)
//   (exclusive prefix sums of both counts: vertex / triangle offsets -- torch.cumsum, plumbing)
//   k_mc_emit  : the point writes its crossing vertices (linear interpolation along the edge, index coordinates as PyMCubes returns
//                them) and its cell's triangles; a triangle corner on cube edge e is the vertex of the grid point that OWNS that edge,
//                found as that point's vertex offset + the rank of the edge's axis among the point's crossings -- vertices are shared
//                between cells without any hashing or sorting, and the output order (points in x-major order, then axis; cells in
//                x-major order, then table order) is deterministic
// HBM-bound streaming work (8 B per grid point for the counts and offsets + the volume itself); no tensor cores.
#include "n2m_common.cuh"
#include "../../include/n2m_b200_mesh.h"

namespace n2m {
namespace {

struct Dims { uint32_t X, Y, Z; };

__device__ __forceinline__ size_t lin(const Dims& d, uint32_t x, uint32_t y, uint32_t z) { return ((size_t)x * d.Y + y) * d.Z + z; }

// crossings on the edges owned by point (x, y, z): bit a set when the edge along axis a exists and is crossed
__device__ __forceinline__ uint32_t own_crossings(const float* __restrict__ vol, const Dims& d, uint32_t x, uint32_t y, uint32_t z, float iso) {
    const bool in0 = vol[lin(d, x, y, z)] > iso;
    uint32_t m = 0;
    if (x + 1 < d.X && (vol[lin(d, x + 1, y, z)] > iso) != in0) m |= 1u;
    if (y + 1 < d.Y && (vol[lin(d, x, y + 1, z)] > iso) != in0) m |= 2u;
    if (z + 1 < d.Z && (vol[lin(d, x, y, z + 1)] > iso) != in0) m |= 4u;
    return m;
}

__device__ __forceinline__ uint32_t cell_case(const float* __restrict__ vol, const Dims& d, uint32_t x, uint32_t y, uint32_t z, float iso) {
    uint32_t mask = 0;
#pragma unroll
    for (uint32_t c = 0; c < 8; ++c)
        if (vol[lin(d, x + (c & 1u), y + ((c >> 1) & 1u), z + ((c >> 2) & 1u))] > iso) mask |= 1u << c;
    return mask;
}

__global__ void __launch_bounds__(256)
k_mc_count(const float* __restrict__ vol, Dims d, float iso, const int32_t* __restrict__ num_tris, uint8_t* __restrict__ vcount,
           uint8_t* __restrict__ tcount) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)d.X * d.Y * d.Z;
    if (p >= n) return;
    const uint32_t z = (uint32_t)(p % d.Z), y = (uint32_t)((p / d.Z) % d.Y), x = (uint32_t)(p / ((size_t)d.Z * d.Y));
    vcount[p] = (uint8_t)__popc(own_crossings(vol, d, x, y, z, iso));
    uint8_t nt = 0;
    if (x + 1 < d.X && y + 1 < d.Y && z + 1 < d.Z) nt = (uint8_t)__ldg(num_tris + cell_case(vol, d, x, y, z, iso));
    tcount[p] = nt;
}

__global__ void __launch_bounds__(256)
k_mc_emit(const float* __restrict__ vol, Dims d, float iso, const int8_t* __restrict__ tri_table, const uint8_t* __restrict__ tcount,
          const int32_t* __restrict__ voff, const int32_t* __restrict__ toff, float* __restrict__ vertices, int32_t* __restrict__ triangles) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)d.X * d.Y * d.Z;
    if (p >= n) return;
    const uint32_t z = (uint32_t)(p % d.Z), y = (uint32_t)((p / d.Z) % d.Y), x = (uint32_t)(p / ((size_t)d.Z * d.Y));
    // ---- vertices on the owned edges ----
    const uint32_t own = own_crossings(vol, d, x, y, z, iso);
    if (own) {
        const float f0 = vol[p];
        int32_t k = voff[p];
#pragma unroll
        for (uint32_t a = 0; a < 3; ++a) {
            if (!((own >> a) & 1u)) continue;
            const float f1 = vol[lin(d, x + (a == 0), y + (a == 1), z + (a == 2))];
            const float t = __fdiv_rn(iso - f0, f1 - f0);
            vertices[3 * (size_t)k + 0] = (float)x + (a == 0 ? t : 0.f);
            vertices[3 * (size_t)k + 1] = (float)y + (a == 1 ? t : 0.f);
            vertices[3 * (size_t)k + 2] = (float)z + (a == 2 ? t : 0.f);
            ++k;
        }
    }
    // ---- triangles of the cell ----
    const uint32_t nt = tcount[p];
    if (nt == 0) return;
    const uint32_t mask = cell_case(vol, d, x, y, z, iso);
    const int8_t* row = tri_table + 16 * mask;
    int32_t* out = triangles + 3 * (size_t)toff[p];
    for (uint32_t i = 0; i < 3 * nt; ++i) {
        const uint32_t e = (uint32_t)__ldg(row + i);
        // edge e = 4 * axis + (b1 + 2 * b2): owned by the point displaced by (b1, b2) on the other two axes (mc_table.py)
        const uint32_t axis = e >> 2, b1 = e & 1u, b2 = (e >> 1) & 1u;
        uint32_t ox = x, oy = y, oz = z;
        if (axis == 0) { oy += b1; oz += b2; } else if (axis == 1) { ox += b1; oz += b2; } else { ox += b1; oy += b2; }
        const uint32_t oc = own_crossings(vol, d, ox, oy, oz, iso);
        out[i] = voff[lin(d, ox, oy, oz)] + (int32_t)__popc(oc & ((1u << axis) - 1u));
    }
}

}  // namespace
}  // namespace n2m

using namespace n2m;

extern "C" {

int n2m_mc_count(const float* volume, uint32_t X, uint32_t Y, uint32_t Z, float iso, const int32_t* num_tris, uint8_t* vcount,
                 uint8_t* tcount, n2m_stream_t stream) {
    N2M_REQUIRE(volume && num_tris && vcount && tcount, "mc_count", "null pointer");
    N2M_REQUIRE(X >= 2 && Y >= 2 && Z >= 2, "mc_count", "the volume needs at least two samples per axis");
    const size_t n = (size_t)X * Y * Z;
    N2M_REQUIRE(n < (1ull << 31), "mc_count", "volume too large");
    k_mc_count<<<(uint32_t)div_up(n, (size_t)256), 256, 0, as_stream(stream)>>>(volume, Dims{X, Y, Z}, iso, num_tris, vcount, tcount);
    return check_launch("mc_count");
}

int n2m_mc_emit(const float* volume, uint32_t X, uint32_t Y, uint32_t Z, float iso, const int8_t* tri_table, const uint8_t* tcount,
                const int32_t* voff, const int32_t* toff, float* vertices, int32_t* triangles, n2m_stream_t stream) {
    N2M_REQUIRE(volume && tri_table && tcount && voff && toff && vertices && triangles, "mc_emit", "null pointer");
    const size_t n = (size_t)X * Y * Z;
    N2M_REQUIRE(X >= 2 && Y >= 2 && Z >= 2 && n < (1ull << 31), "mc_emit", "bad volume size");
    k_mc_emit<<<(uint32_t)div_up(n, (size_t)256), 256, 0, as_stream(stream)>>>(volume, Dims{X, Y, Z}, iso, tri_table, tcount, voff, toff,
                                                                             vertices, triangles);
    return check_launch("mc_emit");
}

}  // extern "C"
