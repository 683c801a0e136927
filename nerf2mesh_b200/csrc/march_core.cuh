// march_core.cuh -- the sequential ray-marcher core shared by raymarching.cu (operator form) and
// stage0.cu (fused train path).  See raymarching.cu's header comment for the arithmetic contract.
#pragma once
#include "n2m_common.cuh"
#include <cfloat>

namespace n2m {
namespace march {

constexpr float kSqrt3 = 1.7320508075688772f;

// ---------------------------------------------------------------------------------------------
// marcher core
// ---------------------------------------------------------------------------------------------
struct MarchCfg {
    float bound;
    float dt_gamma;
    float dt_min, dt_max;
    float Hf, rH, H3f, Hm1f, ncas;   // grid size as float, 1/H, H^3, H-1, cascades as float
    uint32_t H;
    bool contract;
    const uint8_t* __restrict__ bits;
};

__device__ __forceinline__ MarchCfg make_cfg(float bound, bool contract, float dt_gamma,
                                              uint32_t max_steps, uint32_t C, uint32_t H,
                                              const uint8_t* bits) {
    MarchCfg c;
    c.bound = bound;
    c.contract = contract;
    c.dt_gamma = dt_gamma;
    c.dt_min = 2 * kSqrt3 / max_steps;          // raymarching.cu:385
    c.dt_max = 2 * kSqrt3 * bound / H;          // raymarching.cu:386
    c.Hf = (float)H;
    c.rH = 1 / (float)H;
    c.H3f = H * H * H;                          // float, as in the reference (raymarching.cu:379)
    c.Hm1f = (float)(H - 1);
    c.ncas = (float)C;
    c.H = H;
    c.bits = bits;
    return c;
}

// cascade from the position's max-norm: exponent of frexpf clamped to [0, C-1] (raymarching.cu:42-47)
__device__ __forceinline__ int cascade_from_pos(float x, float y, float z, float ncas) {
    const float m = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(m, &e);
    return fminf(ncas - 1, fmaxf(0, e));
}

// cascade from the step size (raymarching.cu:49-54).  The reference multiplies by the double
// literal 0.5; halving is exact in either precision, so fp32 is bit-identical.
__device__ __forceinline__ int cascade_from_dt(float dt, float Hf, float ncas) {
    const float m = dt * Hf * 0.5f;
    int e;
    frexpf(m, &e);
    return fminf(ncas - 1, fmaxf(0, e));
}

struct Probe {
    float cx, cy, cz;   // (contracted) sample position that would be emitted
    float dt;           // step length at this t
    float mip_bound;
    int nx, ny, nz;     // cell in the cascade's H^3 grid
    bool emit;          // occupied (or forced by contraction)
};

// Everything the reference evaluates at the top of one loop iteration (raymarching.cu:397-432).
__device__ __forceinline__ Probe probe_at(const MarchCfg& c, float t, float ox, float oy, float oz,
                                          float dx, float dy, float dz) {
    Probe p;
    const float x = clampf(ox + t * dx, -c.bound, c.bound);
    const float y = clampf(oy + t * dy, -c.bound, c.bound);
    const float z = clampf(oz + t * dz, -c.bound, c.bound);

    p.dt = clampf(t * c.dt_gamma, c.dt_min, c.dt_max);

    const int level = max(cascade_from_pos(x, y, z, c.ncas), cascade_from_dt(p.dt, c.Hf, c.ncas));
    p.mip_bound = fminf(scalbnf(1.0f, level), c.bound);
    const float mip_rbound = 1 / p.mip_bound;

    p.cx = x; p.cy = y; p.cz = z;
    const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    const bool outer = c.contract && mag > 1;
    if (outer) {
        const float s = (2 - 1 / mag) / mag;    // L-inf contraction (raymarching.cu:415)
        p.cx *= s; p.cy *= s; p.cz *= s;
    }

    // nearest cell.  Reference: clamp(0.5 * (c*rb + 1) * H, 0, H-1) with a DOUBLE 0.5: both
    // products are exact in double (24 + 24 significant bits), so one fp32 multiply by H gives
    // the identical correctly-rounded float.
    p.nx = clampf(0.5f * (p.cx * mip_rbound + 1) * c.Hf, 0.0f, c.Hm1f);
    p.ny = clampf(0.5f * (p.cy * mip_rbound + 1) * c.Hf, 0.0f, c.Hm1f);
    p.nz = clampf(0.5f * (p.cz * mip_rbound + 1) * c.Hf, 0.0f, c.Hm1f);

    // bit index: float arithmetic on purpose (level * H3 is a float product in the reference,
    // the Morton code is converted to float for the add, raymarching.cu:426)
    const uint32_t index = level * c.H3f + morton3(p.nx, p.ny, p.nz);
    const bool occ = c.bits[index / 8] & (1 << (index % 8));
    p.emit = occ || outer;
    return p;
}

// Distance at which the ray leaves the current (empty) voxel (raymarching.cu:452-458).
__device__ __forceinline__ float exit_time(const MarchCfg& c, const Probe& p, float t,
                                           float dx, float dy, float dz,
                                           float rdx, float rdy, float rdz) {
    const float tx = (((p.nx + 0.5f + 0.5f * copysignf(1.0f, dx)) * c.rH * 2 - 1) * p.mip_bound - p.cx) * rdx;
    const float ty = (((p.ny + 0.5f + 0.5f * copysignf(1.0f, dy)) * c.rH * 2 - 1) * p.mip_bound - p.cy) * rdy;
    const float tz = (((p.nz + 0.5f + 0.5f * copysignf(1.0f, dz)) * c.rH * 2 - 1) * p.mip_bound - p.cz) * rdz;
    return t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
}

// Empty cell: advance t past the voxel's exit face in dt-sized hops (raymarching.cu:459-464).
__device__ __forceinline__ float hop_to_exit(const MarchCfg& c, const Probe& p, float t,
                                             float dx, float dy, float dz,
                                             float rdx, float rdy, float rdz) {
    const float tt = exit_time(c, p, t, dx, dy, dz, rdx, rdy, rdz);
    do {
        const float dt = clampf(t * c.dt_gamma, c.dt_min, c.dt_max);
        t += dt;
    } while (t < tt);
    return t;
}

// Sinks for the sequential march.
struct CountSink {            // counting pass: remember (t_before, dt) per emitted sample
    float2* slab;             // may be null
    __device__ __forceinline__ void put(uint32_t k, float t_before, float t_after, float dt,
                                        const Probe&, float, float, float) const {
        if (slab) slab[k] = make_float2(t_before, dt);
    }
};
struct WriteSink {            // sequential output pass (no slab): write the sample
    float* xyz; float* dir; float* ts;
    __device__ __forceinline__ void put(uint32_t k, float, float t_after, float dt,
                                        const Probe& p, float dx, float dy, float dz) const {
        xyz[3 * k + 0] = p.cx; xyz[3 * k + 1] = p.cy; xyz[3 * k + 2] = p.cz;
        dir[3 * k + 0] = dx;   dir[3 * k + 1] = dy;   dir[3 * k + 2] = dz;
        ts[2 * k + 0] = t_after; ts[2 * k + 1] = dt;
    }
};

template <typename Sink>
__device__ __forceinline__ uint32_t march_one(const MarchCfg& c, float t, float far, uint32_t limit,
                                              float ox, float oy, float oz,
                                              float dx, float dy, float dz,
                                              float rdx, float rdy, float rdz, const Sink& sink,
                                              float* t_out = nullptr) {
    uint32_t step = 0;
    while (t < far && step < limit) {
        const Probe p = probe_at(c, t, ox, oy, oz, dx, dy, dz);
        if (p.emit) {
            const float t_before = t;
            t += p.dt;
            sink.put(step, t_before, t, p.dt, p, dx, dy, dz);
            step++;
        } else {
            t = hop_to_exit(c, p, t, dx, dy, dz, rdx, rdy, rdz);
        }
    }
    if (t_out) *t_out = t;
    return step;
}


// Slab test against the AABB (raymarching.cu:92-145); returns false on a miss (near = far = FLT_MAX).
__device__ __forceinline__ bool near_far_aabb(const float* __restrict__ o3, const float* __restrict__ d3,
                                              const float* __restrict__ aabb, float min_near, float& near, float& far) {
    float lo = -FLT_MAX, hi = FLT_MAX;
    bool hit = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float o = o3[a];
        const float rd = 1 / d3[a];
        float t0 = (aabb[a] - o) * rd;
        float t1 = (aabb[a + 3] - o) * rd;
        if (t0 > t1) { const float s = t0; t0 = t1; t1 = s; }
        if (a == 0) { lo = t0; hi = t1; }
        else if (hit) {
            if (lo > t1 || t0 > hi) hit = false;
            else {
                if (t0 > lo) lo = t0;
                if (t1 < hi) hi = t1;
            }
        }
    }
    if (!hit) { near = FLT_MAX; far = FLT_MAX; return false; }
    if (lo < min_near) lo = min_near;
    near = lo; far = hi;
    return true;
}

}  // namespace march
}  // namespace n2m
