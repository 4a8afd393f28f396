// Per-level voxel location, hashing and trilinear interpolation shared by the stand-alone
// hash-encoder kernels (hashenc.cu) and the fused field kernels (field.cu).
//
// The arithmetic restates hash_encoding.py:24-46,78-94 of the reference one torch op per
// rounded fp32 operation (explicit __f*_rn intrinsics: no contraction, no fast reciprocal):
// integer slots are bit-exact by construction and so are the interpolated features.
#pragma once
#include "nsig_common.cuh"

namespace nsig {

constexpr uint32_t kPrimeY = 2654435761u;  // hash_encoding.py:16
constexpr uint32_t kPrimeZ = 805459861u;

struct Voxel {
    uint32_t hx0, hx1, hy0, hy1, hz0, hz1;  // per-axis hash terms of the lower / upper corner
    float wx, wy, wz;                       // interpolation weights
};

// x in the unit box; grid_size = fl(1/resolution) computed by the host in IEEE fp32.
__device__ __forceinline__ void locate_axis(float x, float grid_size, uint32_t& idx, float& w) {
    const float xc = fminf(fmaxf(x, 0.0f), 1.0f);            // hash_encoding.py:33-35
    const float fi = floorf(__fdiv_rn(xc, grid_size));       // :39  floor((x - 0)/grid_size)
    idx = (uint32_t)(int)fi;
    const float vmin = __fmul_rn(fi, grid_size);             // :40
    const float vmax = __fadd_rn(vmin, grid_size);           // :41
    w = __fdiv_rn(__fsub_rn(x, vmin), __fsub_rn(vmax, vmin)); // :87 (x unclamped)
}

__device__ __forceinline__ Voxel locate(float x, float y, float z, float grid_size) {
    Voxel v;
    uint32_t ix, iy, iz;
    locate_axis(x, grid_size, ix, v.wx);
    locate_axis(y, grid_size, iy, v.wy);
    locate_axis(z, grid_size, iz, v.wz);
    v.hx0 = ix;            v.hx1 = ix + 1u;
    v.hy0 = iy * kPrimeY;  v.hy1 = v.hy0 + kPrimeY;
    v.hz0 = iz * kPrimeZ;  v.hz1 = v.hz0 + kPrimeZ;
    return v;
}

// corner k = i*4 + j*2 + l with x the most significant bit (hash_encoding.py:8)
__device__ __forceinline__ uint32_t corner_slot(const Voxel& v, int k, uint32_t mask) {
    return (((k & 4) ? v.hx1 : v.hx0) ^ ((k & 2) ? v.hy1 : v.hy0) ^ ((k & 1) ? v.hz1 : v.hz0)) & mask;
}

__device__ __forceinline__ float lerp_ref(float a, float b, float w, float omw) {
    return __fadd_rn(__fmul_rn(a, omw), __fmul_rn(b, w));
}

// trilinear interpolation in the reference's order: x, then y, then z (hash_encoding.py:89-104)
__device__ __forceinline__ float2 trilerp(const float2 e[8], const Voxel& v) {
    const float ox = __fsub_rn(1.0f, v.wx), oy = __fsub_rn(1.0f, v.wy), oz = __fsub_rn(1.0f, v.wz);
    float2 r;
    {
        const float c00 = lerp_ref(e[0].x, e[4].x, v.wx, ox), c01 = lerp_ref(e[1].x, e[5].x, v.wx, ox);
        const float c10 = lerp_ref(e[2].x, e[6].x, v.wx, ox), c11 = lerp_ref(e[3].x, e[7].x, v.wx, ox);
        const float c0 = lerp_ref(c00, c10, v.wy, oy), c1 = lerp_ref(c01, c11, v.wy, oy);
        r.x = lerp_ref(c0, c1, v.wz, oz);
    }
    {
        const float c00 = lerp_ref(e[0].y, e[4].y, v.wx, ox), c01 = lerp_ref(e[1].y, e[5].y, v.wx, ox);
        const float c10 = lerp_ref(e[2].y, e[6].y, v.wx, ox), c11 = lerp_ref(e[3].y, e[7].y, v.wx, ox);
        const float c0 = lerp_ref(c00, c10, v.wy, oy), c1 = lerp_ref(c01, c11, v.wy, oy);
        r.y = lerp_ref(c0, c1, v.wz, oz);
    }
    return r;
}

// gather the 8 corners of one level and interpolate
__device__ __forceinline__ float2 encode_level(const float2* __restrict__ table, const Voxel& v, uint32_t mask) {
    float2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = __ldg(table + corner_slot(v, k, mask));
    return trilerp(e, v);
}

// d(loss)/d(table[slot_k]) contribution of output gradient g, in autograd's association order
// through hash_encoding.py:89-104: ((g * z-factor) * y-factor) * x-factor
__device__ __forceinline__ float corner_grad(const Voxel& v, int k, float g) {
    const float fz = (k & 1) ? v.wz : __fsub_rn(1.0f, v.wz);
    const float fy = (k & 2) ? v.wy : __fsub_rn(1.0f, v.wy);
    const float fx = (k & 4) ? v.wx : __fsub_rn(1.0f, v.wx);
    return __fmul_rn(__fmul_rn(__fmul_rn(g, fz), fy), fx);
}

// vectorised fp32 reduction into global memory (sm_90+): one L2 atomic for both features
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

struct TablePtrs {
    const float2* t[NSIG_MAX_LEVELS];
    float grid_size[NSIG_MAX_LEVELS];  // fl(1/resolution)
};

// ---------------------------------------------------------------------------------------------------
// Cheaper per-level geometry for the FUSED kernels (field.cu), whose features are rounded to fp16 MMA
// operands anyway.  Integer slots stay bit-exact, interpolation weights may differ from the reference's
// fp32 value in the last ulp:
//   * index: fl32(x / g) is obtained as fl32(fl64(x) * fl64(1/g)).  This is the SAME float for every input:
//     the double product is within 2^-52 (relative) of x/g, while x/g is either exactly representable
//     (g a power of two) or at relative distance >= 2^-49 from every fp32 rounding midpoint (a midpoint m has
//     25 significant bits with the last one set; m*g == x would need m*g to fit in 24 bits, impossible for a g
//     whose mantissa is not 1).  One DMUL + one conversion instead of an IEEE division (~10 instructions with
//     a guarded slow path).  Checked at every cell boundary of all 17 resolutions +-40 ulp and on 10^7 random points:
//     on the device through nsig_fused_hash_slots (tests/test_hash_gpu.py::test_fused_slots_equal_reference_order_slots),
//     in numpy in tests/test_host_cpu.py::test_fused_index_identity_double_multiply_equals_fp32_division.
//   * weight: (x - vmin) * fl(1/g) instead of (x - vmin) / (vmax - vmin)  (relative difference <= 2e-7).
//   * trilinear interpolation with FMA contraction (a*(1-w) + b*w -> fma(b, w, a*(1-w))).
// ---------------------------------------------------------------------------------------------------
struct LevelGeom {
    double rgd;    // 1.0 / (double)gs
    float gs;      // fl(1/resolution)   (hash_encoding.py:37)
    float inv_gs;  // fl(1/gs)
};

__host__ inline LevelGeom make_level_geom(float resolution) {
    LevelGeom L;
    L.gs = 1.0f / resolution;
    L.rgd = 1.0 / (double)L.gs;
    L.inv_gs = (float)L.rgd;
    return L;
}

__device__ __forceinline__ void locate_axis_fused(float x, const LevelGeom& L, uint32_t& idx, float& w) {
    const float xc = fminf(fmaxf(x, 0.0f), 1.0f);
    const float q = __double2float_rn(__dmul_rn((double)xc, L.rgd));  // == __fdiv_rn(xc, L.gs), see above
    const float fi = floorf(q);
    idx = (uint32_t)(int)fi;
    w = __fmul_rn(__fsub_rn(x, __fmul_rn(fi, L.gs)), L.inv_gs);
}

__device__ __forceinline__ Voxel locate_fused(float x, float y, float z, const LevelGeom& L) {
#ifdef NSIG_EXACT_GEOM  // A/B switch for profiling: the reference-order arithmetic of the stand-alone encoder
    return locate(x, y, z, L.gs);
#endif
    Voxel v;
    uint32_t ix, iy, iz;
    locate_axis_fused(x, L, ix, v.wx);
    locate_axis_fused(y, L, iy, v.wy);
    locate_axis_fused(z, L, iz, v.wz);
    v.hx0 = ix;            v.hx1 = ix + 1u;
    v.hy0 = iy * kPrimeY;  v.hy1 = v.hy0 + kPrimeY;
    v.hz0 = iz * kPrimeZ;  v.hz1 = v.hz0 + kPrimeZ;
    return v;
}

__device__ __forceinline__ float lerp_fma(float a, float b, float w, float omw) { return fmaf(b, w, a * omw); }

__device__ __forceinline__ float2 trilerp_fma(const float2 e[8], const Voxel& v) {
    const float ox = 1.0f - v.wx, oy = 1.0f - v.wy, oz = 1.0f - v.wz;
    float2 r;
    {
        const float c00 = lerp_fma(e[0].x, e[4].x, v.wx, ox), c01 = lerp_fma(e[1].x, e[5].x, v.wx, ox);
        const float c10 = lerp_fma(e[2].x, e[6].x, v.wx, ox), c11 = lerp_fma(e[3].x, e[7].x, v.wx, ox);
        r.x = lerp_fma(lerp_fma(c00, c10, v.wy, oy), lerp_fma(c01, c11, v.wy, oy), v.wz, oz);
    }
    {
        const float c00 = lerp_fma(e[0].y, e[4].y, v.wx, ox), c01 = lerp_fma(e[1].y, e[5].y, v.wx, ox);
        const float c10 = lerp_fma(e[2].y, e[6].y, v.wx, ox), c11 = lerp_fma(e[3].y, e[7].y, v.wx, ox);
        r.y = lerp_fma(lerp_fma(c00, c10, v.wy, oy), lerp_fma(c01, c11, v.wy, oy), v.wz, oz);
    }
    return r;
}

// (An x-pair merged variant - one 128-bit load for both x-neighbours when the lower x index is even - was measured
// SLOWER, 0.168 vs 0.137 ms per 555k samples: the L1 data pipe returns 32 lanes x 4 B per wavefront, so wider
// per-lane loads cost proportionally more wavefronts.  profiles/r01_field_kernels_v1.txt.)
__device__ __forceinline__ float2 encode_level_fused(const float2* __restrict__ table, const Voxel& v, uint32_t mask) {
    float2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = __ldg(table + corner_slot(v, k, mask));
#ifdef NSIG_EXACT_GEOM
    return trilerp(e, v);
#endif
    return trilerp_fma(e, v);
}

// half2 shadow tables (north_star kernel 2: "vectorised half2 loads"): entry = (f0, f1) * scale_l rounded to fp16,
// scale_l a power of two chosen per level so the largest magnitude sits just below 2^15 (no fp16 subnormals for
// U(+-1e-4) initialisations, no overflow for trained tables).  One 32-bit load per corner: the L1 data pipe moves
// 32 lanes x 4 B per wavefront, so this halves the wavefronts of the gather against the fp32 (8 B per lane) tables.
// Interpolation stays fp32; the result is multiplied by 1/scale_l (exact).  Absolute error per feature
// <= 2^-11 * max|corner|, the same size as the fp16 rounding the MMA operands get anyway.
__device__ __forceinline__ float2 encode_level_fused_h2(const __half2* __restrict__ table, const Voxel& v, uint32_t mask,
                                                        float inv_scale) {
    // (An x-pair variant - ONE aligned 8-byte load for both x-neighbours when ix is even, whose slots differ in bit 0 only -
    // was measured in round 2: 0.302 vs 0.221 ms per 1.07 M samples.  The divergent even/odd paths execute 12 instead of 8
    // load instructions per level and the L1 tag stage counts SECTORS, not instructions: tag requests stayed at 55 per
    // sample.  profiles/r02_experiments.txt.)
    float2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(table) + corner_slot(v, k, mask));
        e[k] = __half22float2(*reinterpret_cast<const __half2*>(&raw));
    }
    const float2 r = trilerp_fma(e, v);
    return make_float2(r.x * inv_scale, r.y * inv_scale);
}

struct FusedTablePtrs {
    const float2* t[NSIG_MAX_LEVELS];
    LevelGeom geom[NSIG_MAX_LEVELS];
    const __half2* th[NSIG_MAX_LEVELS];  // half2 shadow tables (all null: fp32 tables are gathered)
    const float* inv_scale;              // device float[NSIG_MAX_LEVELS] = 1 / scale_l
};

}  // namespace nsig
