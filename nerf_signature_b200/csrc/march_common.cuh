// Device-side core of the occupancy-grid marcher and the warp scan helpers, shared by the stand-alone
// raymarching kernels (raymarch.cu) and the fused frame renderer (render.cu).
//
// Bit-exactness: every float operation that feeds an integer decision (grid cell, occupancy, skip
// distance, loop exit) is written with explicit __f*_rn intrinsics in the operation order and FMA
// contraction the reference build produces (nvcc default -fmad=true, IEEE div), so the compiler cannot
// re-associate or re-contract it.
#pragma once
#include "nsig_common.cuh"

namespace nsig {

constexpr float kSqrt3 = 1.7320508075688772f;
constexpr float kRPi = 0.3183098861837907f;

// ---------------------------------------------------------------------------------------
// K1 near/far — reference raymarching.cu:92-145
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void near_far_one(float ox, float oy, float oz, float dx, float dy,
                                             float dz, const float* __restrict__ aabb,
                                             float min_near, float& near_out, float& far_out) {
    const float rdx = __fdiv_rn(1.0f, dx), rdy = __fdiv_rn(1.0f, dy), rdz = __fdiv_rn(1.0f, dz);
    float near = __fmul_rn(__fsub_rn(aabb[0], ox), rdx);
    float far = __fmul_rn(__fsub_rn(aabb[3], ox), rdx);
    if (near > far) { float c = near; near = far; far = c; }
    float near_y = __fmul_rn(__fsub_rn(aabb[1], oy), rdy);
    float far_y = __fmul_rn(__fsub_rn(aabb[4], oy), rdy);
    if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
    if (near > far_y || near_y > far) { near_out = far_out = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(__fsub_rn(aabb[2], oz), rdz);
    float far_z = __fmul_rn(__fsub_rn(aabb[5], oz), rdz);
    if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
    if (near > far_z || near_z > far) { near_out = far_out = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_out = near;
    far_out = far;
}

// ---------------------------------------------------------------------------------------
// Marching core
// ---------------------------------------------------------------------------------------
struct MarchCfg {
    float bound, dt_gamma, dt_min, dt_max;
    float rH;    // 1 / (float)H            (raymarching.cu:338)
    float H3f;   // (float)(H*H*H)          (raymarching.cu:339)
    float Hf;    // (float)H
    float Hm1f;  // (float)(H-1)
    float Cf;    // (float)C
    double Hd;   // (double)H
    float mb0, rmb0;  // level-0 mip_bound = fminf(1, bound) and its reciprocal (the only level when C == 1)
};

__device__ __forceinline__ MarchCfg make_cfg(float bound, float dt_gamma, uint32_t max_steps,
                                             uint32_t C, uint32_t H) {
    MarchCfg c;
    c.bound = bound;
    c.dt_gamma = dt_gamma;
    // dt_min = 2*SQRT3()/max_steps ; dt_max = 2*SQRT3()*(1<<(C-1))/H   (raymarching.cu:345-346)
    const float two_s3 = __fmul_rn(2.0f, kSqrt3);
    c.dt_min = __fdiv_rn(two_s3, (float)max_steps);
    c.dt_max = __fdiv_rn(__fmul_rn(two_s3, (float)(1 << (C - 1))), (float)H);
    c.rH = __fdiv_rn(1.0f, (float)H);
    c.H3f = (float)(H * H * H);
    c.Hf = (float)H;
    c.Hm1f = (float)(H - 1);
    c.Cf = (float)C;
    c.Hd = (double)H;
    c.mb0 = fminf(scalbnf(1.0f, 0), bound);
    c.rmb0 = __fdiv_rn(1.0f, c.mb0);
    return c;
}

struct RayConst {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float hsx, hsy, hsz;  // 0.5f * signf(d)
};

__device__ __forceinline__ RayConst load_ray(const float* __restrict__ rays_o,
                                             const float* __restrict__ rays_d, uint32_t n) {
    RayConst r;
    r.ox = rays_o[n * 3]; r.oy = rays_o[n * 3 + 1]; r.oz = rays_o[n * 3 + 2];
    r.dx = rays_d[n * 3]; r.dy = rays_d[n * 3 + 1]; r.dz = rays_d[n * 3 + 2];
    r.rdx = __fdiv_rn(1.0f, r.dx); r.rdy = __fdiv_rn(1.0f, r.dy); r.rdz = __fdiv_rn(1.0f, r.dz);
    r.hsx = __fmul_rn(0.5f, copysignf(1.0f, r.dx));
    r.hsy = __fmul_rn(0.5f, copysignf(1.0f, r.dy));
    r.hsz = __fmul_rn(0.5f, copysignf(1.0f, r.dz));
    return r;
}

// dt(t) = clamp(t*dt_gamma, dt_min, dt_max)   (raymarching.cu:365,396)
__device__ __forceinline__ float step_dt(float t, const MarchCfg& c) {
    return clampf(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max);
}

// (int) clamp(0.5 * (p * mip_rbound + 1) * H, 0.0f, (float)(H - 1))   (raymarching.cu:374-376):
// fp32 FMA, then two fp64 multiplies (the literal 0.5 is a double), rounded to fp32, clamped,
// truncated.
__device__ __forceinline__ int grid_coord(float p, float mip_rbound, const MarchCfg& c) {
    const float v = __fmaf_rn(p, mip_rbound, 1.0f);
    const double d = __dmul_rn(__dmul_rn(0.5, (double)v), c.Hd);
    return (int)clampf(__double2float_rn(d), 0.0f, c.Hm1f);
}

// distance along the ray to the exit face of cell `n` on one axis (raymarching.cu:389-391):
// (((n + 0.5f + 0.5f*sign(d)) * rH * 2 - 1) * mip_bound - p) * rd
__device__ __forceinline__ float exit_dist(int n, float hs, float p, float rd, float mip_bound,
                                           const MarchCfg& c) {
    const float a = __fadd_rn(__fadd_rn((float)n, 0.5f), hs);
    const float b = __fmul_rn(a, c.rH);
    const float e = __fadd_rn(__fmul_rn(b, 2.0f), -1.0f);  // b*2 is exact: FMA or not, same bits
    const float f = __fmaf_rn(e, mip_bound, -p);
    return __fmul_rn(f, rd);
}

struct Point {
    float x, y, z, dt;
    float tt;  // skip target when not occupied
    bool occ;
};

// One iteration body of the reference loop (raymarching.cu:357-399) evaluated at lattice point t.
__device__ __forceinline__ Point eval_point(float t, const RayConst& r, const MarchCfg& c,
                                            const uint8_t* __restrict__ grid) {
    Point p;
    p.x = clampf(__fmaf_rn(t, r.dx, r.ox), -c.bound, c.bound);
    p.y = clampf(__fmaf_rn(t, r.dy, r.oy), -c.bound, c.bound);
    p.z = clampf(__fmaf_rn(t, r.dz, r.oz), -c.bound, c.bound);
    p.dt = step_dt(t, c);

    // mip_from_pos / mip_from_dt (raymarching.cu:42-54).  With one cascade both are clamped to [0, C-1] = {0}: the
    // two frexpf, the fp64 product, scalbnf and the IEEE division cannot change the result, so the (warp-uniform)
    // C == 1 case takes level 0 and the precomputed level-0 bound (same operations, evaluated once in make_cfg).
    int level = 0;
    float mip_bound = c.mb0, mip_rbound = c.rmb0;
    if (c.Cf != 1.0f) {
        int ep, ed;
        frexpf(fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))), &ep);
        const int lp = (int)fminf(c.Cf - 1.0f, fmaxf(0.0f, (float)ep));
        const float mxd = __double2float_rn(__dmul_rn((double)__fmul_rn(p.dt, c.Hf), 0.5));
        frexpf(mxd, &ed);
        const int ld = (int)fminf(c.Cf - 1.0f, fmaxf(0.0f, (float)ed));
        level = max(lp, ld);
        mip_bound = fminf(scalbnf(1.0f, level), c.bound);
        mip_rbound = __fdiv_rn(1.0f, mip_bound);
    }

    const int nx = grid_coord(p.x, mip_rbound, c);
    const int ny = grid_coord(p.y, mip_rbound, c);
    const int nz = grid_coord(p.z, mip_rbound, c);

    // index = level * H3 + morton: evaluated in fp32 like the reference (H3 is a float there)
    const uint32_t m = morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const uint32_t index = (uint32_t)__fmaf_rn((float)level, c.H3f, (float)m);
    p.occ = (__ldg(grid + (index >> 3)) & (1u << (index & 7u))) != 0;

    const float tx = exit_dist(nx, r.hsx, p.x, r.rdx, mip_bound, c);
    const float ty = exit_dist(ny, r.hsy, p.y, r.rdy, mip_bound, c);
    const float tz = exit_dist(nz, r.hsz, p.z, r.rdz, mip_bound, c);
    p.tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    return p;
}

// One 32-point window of the per-ray lattice.  State (carried across windows):
//   t        first lattice point of the next window
//   carry_tt pending skip target of the last visited, unoccupied point
//   last_t   t after the previous emitted sample (raymarching.cu:424)
struct MarchState {
    float t, carry_tt, last_t;
};

struct Window {
    Point p;           // this lane's lattice point (valid where `emitted` has the lane's bit)
    uint32_t emitted;  // lanes whose point the reference's sequential loop visits AND finds occupied
    float delta_real;  // deltas[1] of this lane's sample: t after it minus t after the previous sample
};

// Evaluates lattice points t_k (k = lane) of the window starting at s.t, decides which of them the
// reference's sequential loop `while (t < far)` would visit, and advances the state.  Emits samples in
// the exact order and with the exact values of the reference.
__device__ __forceinline__ Window march_window(MarchState& s, const RayConst& r, const MarchCfg& c,
                                               const uint8_t* __restrict__ grid, float far, int lane,
                                               uint32_t lt_mask) {
    // lattice: t_{k+1} = t_k + dt(t_k), sequential fp32 adds exactly like the reference
    float my_t = s.t, tc = s.t;
    if (c.dt_gamma == 0.0f) {  // dt(t) == dt_min for every finite t: skip the clamp chain
        // Closed form of the sequential adds.  While t stays inside one binade every t_k is a multiple of
        // u = ulp(t) and fl(t_k + dt) = t_k + step with the SAME step = fl(t_0 + dt) - t_0 (dt rounded to a multiple of
        // u), unless dt sits exactly half-way between two multiples (then round-to-even alternates with t_k's parity).
        // So t_k = t_0 + k*step, exactly representable, one FMA per lane instead of a 32-deep dependent chain.  The
        // serial chain remains for windows that cross a power of two, for the tie case and for non-normal t.
        const float t1 = __fadd_rn(s.t, c.dt_min);
        const float step = __fsub_rn(t1, s.t);                      // exact (both multiples of u, within 2x of each other)
        const float hi = __fmaf_rn(32.0f, step, s.t);               // t_32 if the closed form holds
        const uint32_t b0 = __float_as_uint(s.t), bh = __float_as_uint(hi);
        const float u = __uint_as_float((b0 & 0x7f800000u) - (23u << 23));          // ulp of the binade of t_0
        const float resid = __fsub_rn(c.dt_min, step);              // exact: |resid| <= u/2, a multiple of ulp(dt)
        const bool closed = (b0 >> 23) == (bh >> 23) && (b0 >> 23) > 24u && (b0 >> 23) < 255u &&
                            fabsf(resid) != __fmul_rn(0.5f, u) && t1 > s.t;
        if (closed) {
            my_t = __fmaf_rn((float)lane, step, s.t);
            tc = hi;
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k == lane) my_t = tc;
                tc = __fadd_rn(tc, c.dt_min);
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (k == lane) my_t = tc;
            tc = __fadd_rn(tc, step_dt(tc, c));
        }
    }
    const bool inr = my_t < far;
    Window w;
    w.p.occ = false; w.p.tt = -INFINITY; w.p.dt = 0.f; w.p.x = w.p.y = w.p.z = 0.f;
    if (inr) w.p = eval_point(my_t, r, c, grid);

    // successor of this lattice point if the sequential loop visits it:
    //   occupied  -> next lattice point
    //   otherwise -> first lattice point with t >= tt (do { t += dt } while (t < tt))
    int nxt = lane + 1;
    {   // binary search (all lanes participate in the shuffles)
        const float key = (inr && !w.p.occ) ? w.p.tt : -INFINITY;
        const float t31 = __shfl_sync(NSIG_FULL_MASK, my_t, 31);
        int pos = 0;
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
            const float tv = __shfl_sync(NSIG_FULL_MASK, my_t, pos + sft - 1);
            if (tv < key) pos += sft;
        }
        if (t31 < key) pos = 32;
        if (inr && !w.p.occ) nxt = max(lane + 1, pos);
    }

    // entry point of this window: first lattice point with t >= carry_tt
    const int entry = __popc(__ballot_sync(NSIG_FULL_MASK, my_t < s.carry_tt));

    // visited set = orbit of `entry` under nxt; pointer doubling over 5 rounds
    uint32_t reach = 1u << lane;
    int jump = nxt;
#pragma unroll
    for (int rnd = 0; rnd < 5; ++rnd) {
        const uint32_t m2 = __shfl_sync(NSIG_FULL_MASK, reach, jump & 31);
        const int j2 = __shfl_sync(NSIG_FULL_MASK, jump, jump & 31);
        if (jump < 32) { reach |= m2; jump = j2; }
    }
    const uint32_t visited = (entry < 32) ? __shfl_sync(NSIG_FULL_MASK, reach, entry & 31) : 0u;
    const uint32_t occ_mask = __ballot_sync(NSIG_FULL_MASK, inr && w.p.occ);
    w.emitted = visited & occ_mask;

    if (visited) {
        const int last = 31 - __clz(visited);
        const float tt_last = __shfl_sync(NSIG_FULL_MASK, w.p.tt, last);
        s.carry_tt = ((occ_mask >> last) & 1u) ? -INFINITY : tt_last;
    }

    const float tnext = __fadd_rn(my_t, w.p.dt);
    const uint32_t before = w.emitted & lt_mask;
    const int prev = before ? (31 - __clz(before)) : 0;
    float lt = __shfl_sync(NSIG_FULL_MASK, tnext, prev);
    if (!before) lt = s.last_t;
    w.delta_real = __fsub_rn(tnext, lt);
    if (w.emitted) s.last_t = __shfl_sync(NSIG_FULL_MASK, tnext, 31 - __clz(w.emitted));
    s.t = tc;
    return w;
}

// Warp-cooperative march of one ray.  Emits (at most `limit`) samples in the exact order and with
// the exact values of the reference's sequential loop `while (t < far && step < limit)`.
// When WRITE, sample k goes to row k of xyzs/dirs/deltas (already offset to the ray's first row).
// Returns the number of samples.
template <bool WRITE>
__device__ __forceinline__ uint32_t warp_march(const RayConst& r, const MarchCfg& c,
                                               const uint8_t* __restrict__ grid, float t0, float far,
                                               uint32_t limit, float* __restrict__ xyzs,
                                               float* __restrict__ dirs, float* __restrict__ deltas,
                                               int lane) {
    MarchState s{t0, -INFINITY, t0};
    uint32_t count = 0;
    const uint32_t lt_mask = lanemask_lt();
    while (s.t < far && count < limit) {
        const Window w = march_window(s, r, c, grid, far, lane, lt_mask);
        if (WRITE) {
            const uint32_t rank = count + __popc(w.emitted & lt_mask);
            if (((w.emitted >> lane) & 1u) && rank < limit) {
                float* px = xyzs + (size_t)rank * 3;
                float* pd = dirs + (size_t)rank * 3;
                float* pl = deltas + (size_t)rank * 2;
                px[0] = w.p.x; px[1] = w.p.y; px[2] = w.p.z;
                pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                pl[0] = w.p.dt;
                pl[1] = w.delta_real;
            }
        }
        count = min(limit, count + (uint32_t)__popc(w.emitted));
    }
    return count;
}

// t0 = near + clamp(near*dt_gamma, dt_min, dt_max) * noise   (raymarching.cu:348-351; FMA-contracted)
__device__ __forceinline__ float perturbed_start(float t, float noise, const MarchCfg& c) {
    return __fmaf_rn(step_dt(t, c), noise, t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(NSIG_FULL_MASK, v, d);
    return v;
}
// inclusive scans across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(NSIG_FULL_MASK, v, d);
        if (lane >= d) v *= o;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(NSIG_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

}  // namespace nsig
