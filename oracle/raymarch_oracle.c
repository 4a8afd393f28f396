/*
 * raymarch_oracle.c — CPU restatement of the reference's raymarching kernels.
 *
 * TEST INFRASTRUCTURE ONLY: may be used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, never by the product path.
 *
 * Each function follows one kernel of /root/reference/raymarching/src/raymarching.cu, one
 * "thread" (ray) at a time, in the reference's own sequential order.  Floating-point
 * contraction follows the SASS of the reference built for sm_100a with nvcc defaults
 * (oracle/build_ref.py; cuobjdump -sass of kernel_march_rays<float>): fmaf() is used exactly
 * where that build emits FFMA on an inexact product, plain ops elsewhere.  Compile with
 * -ffp-contract=off so the host compiler adds no contraction of its own.
 *
 * Parity pin: the reference ships no tests or golden vectors.  This file is
 * pinned against the unmodified reference CUDA build (oracle/_ref/_raymarching.so) on the GPU
 * box by tests/test_ref_cuda_parity.py and by the fixtures under tests/golden/ that were
 * generated from it (tests/golden/make_golden_raymarch.py).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf_(float x) { return copysignf(1.0f, x); }

#define SQRT3 1.7320508075688772f
#define RPI 0.3183098861837907f

/* raymarching.cu:42-47 */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
/* raymarching.cu:49-54 */
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int exponent;
    frexpf(mx, &exponent);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)exponent));
}
/* raymarching.cu:56-72 */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D_(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert_(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:92-145 */
void oracle_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb,
                               uint32_t N, float min_near, float* nears, float* fars) {
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
        float near = (aabb[0] - ox) * rdx;
        float far = (aabb[3] - ox) * rdx;
        if (near > far) { float c = near; near = far; far = c; }
        float near_y = (aabb[1] - oy) * rdy;
        float far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz;
        float far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:163-198 (float tolerance only: libm vs CUDA atan2f/sqrtf) */
void oracle_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N,
                         float* coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = dx * dx + dy * dy + dz * dz;
        const float B = ox * dx + oy * dy + oz * dz;
        const float C = ox * ox + oy * oy + oz * oz - radius * radius;
        const float t = (-B + sqrtf(B * B - A * C)) / A;
        const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
        const float theta = atan2f(sqrtf(x * x + z * z), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = 2 * theta * RPI - 1;
        coords[n * 2 + 1] = phi * RPI;
    }
}

/* raymarching.cu:214-226 */
void oracle_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
    for (uint32_t n = 0; n < N; ++n)
        indices[n] = (int32_t)morton3D_((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1],
                                        (uint32_t)coords[n * 3 + 2]);
}
/* raymarching.cu:237-254 */
void oracle_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
    for (uint32_t n = 0; n < N; ++n) {
        const int32_t ind = indices[n];
        coords[n * 3] = (int32_t)morton3D_invert_((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int32_t)morton3D_invert_((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int32_t)morton3D_invert_((uint32_t)(ind >> 2));
    }
}
/* raymarching.cu:268-289 */
void oracle_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield) {
    for (uint32_t n = 0; n < N; ++n) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; ++i)
            bits |= (grid[(size_t)n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* ---- shared marching state: one loop iteration of raymarching.cu:357-399 ---- */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, dt_gamma, dt_min, dt_max, rH, H3;
    uint32_t C, H;
    const uint8_t* grid;
} march_t;

static inline void march_setup(march_t* m, const float* o, const float* d, const uint8_t* grid,
                               float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                               uint32_t H) {
    m->ox = o[0]; m->oy = o[1]; m->oz = o[2];
    m->dx = d[0]; m->dy = d[1]; m->dz = d[2];
    m->rdx = 1 / m->dx; m->rdy = 1 / m->dy; m->rdz = 1 / m->dz;
    m->bound = bound; m->dt_gamma = dt_gamma;
    m->rH = 1 / (float)H;
    m->H3 = (float)(H * H * H);
    m->dt_min = 2 * SQRT3 / (float)max_steps;
    m->dt_max = 2 * SQRT3 * (float)(1 << (C - 1)) / (float)H;
    m->C = C; m->H = H; m->grid = grid;
}

/* Evaluates the loop body at *t; returns 1 when the cell is occupied (caller emits and adds dt),
 * otherwise advances *t past the cell (the do/while of raymarching.cu:394-398) and returns 0. */
static inline int march_step(const march_t* m, float* t, float* x, float* y, float* z, float* dt) {
    const float tt0 = *t;
    *x = clampf(fmaf(tt0, m->dx, m->ox), -m->bound, m->bound);
    *y = clampf(fmaf(tt0, m->dy, m->oy), -m->bound, m->bound);
    *z = clampf(fmaf(tt0, m->dz, m->oz), -m->bound, m->bound);
    *dt = clampf(tt0 * m->dt_gamma, m->dt_min, m->dt_max);

    const int a = mip_from_pos(*x, *y, *z, (float)m->C);
    const int b = mip_from_dt(*dt, (float)m->H, (float)m->C);
    const int level = a > b ? a : b;

    const float mip_bound = fminf(scalbnf(1.0f, level), m->bound);
    const float mip_rbound = 1 / mip_bound;

    const int nx = (int)clampf((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
    const int ny = (int)clampf((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));
    const int nz = (int)clampf((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)m->H), 0.0f, (float)(m->H - 1));

    const uint32_t index = (uint32_t)fmaf((float)level, m->H3, (float)morton3D_((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = (m->grid[index / 8] & (1u << (index % 8))) != 0;
    if (occ) return 1;

    const float tx = fmaf(mip_bound, (((float)nx + 0.5f + 0.5f * signf_(m->dx)) * m->rH) * 2 - 1, -*x) * m->rdx;
    const float ty = fmaf(mip_bound, (((float)ny + 0.5f + 0.5f * signf_(m->dy)) * m->rH) * 2 - 1, -*y) * m->rdy;
    const float tz = fmaf(mip_bound, (((float)nz + 0.5f + 0.5f * signf_(m->dz)) * m->rH) * 2 - 1, -*z) * m->rdz;
    const float tt = tt0 + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tc = tt0;
    do {
        tc += clampf(tc * m->dt_gamma, m->dt_min, m->dt_max);
    } while (tc < tt);
    *t = tc;
    return 0;
}

/* raymarching.cu:312-480.  Rays are processed in index order, so ray n owns row n of `rays`
 * and offsets are the running sum — one of the orders the reference's atomics can produce. */
void oracle_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid,
                             float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C,
                             uint32_t H, uint32_t M, const float* nears, const float* fars,
                             float* xyzs, float* dirs, float* deltas, int32_t* rays,
                             int32_t* counter, const float* noises) {
    for (uint32_t n = 0; n < N; ++n) {
        march_t m;
        march_setup(&m, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H);
        const float near = nears[n], far = fars[n], noise = noises ? noises[n] : 0.0f;
        const float t0 = fmaf(clampf(near * dt_gamma, m.dt_min, m.dt_max), noise, near);

        float t = t0, x, y, z, dt;
        uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) {
            if (march_step(&m, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
        }
        const uint32_t point_index = (uint32_t)counter[0];
        counter[0] += (int32_t)num_steps;
        const uint32_t ray_index = (uint32_t)counter[1];
        counter[1] += 1;
        rays[ray_index * 3] = (int32_t)n;
        rays[ray_index * 3 + 1] = (int32_t)point_index;
        rays[ray_index * 3 + 2] = (int32_t)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps > M) continue;

        float* px = xyzs + (size_t)point_index * 3;
        float* pd = dirs + (size_t)point_index * 3;
        float* pl = deltas + (size_t)point_index * 2;
        t = t0;
        uint32_t step = 0;
        float last_t = t;
        while (t < far && step < num_steps) {
            if (march_step(&m, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* raymarching.cu:501-577 (expf stands in for __expf: float tolerance) */
void oracle_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                         const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                         float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1],
                       num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        const float* s = sigmas + offset; const float* c = rgbs + (size_t)offset * 3;
        const float* dl = deltas + (size_t)offset * 2;
        uint32_t step = 0;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        while (step < num_steps) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float weight = alpha * T;
            r += weight * c[0]; g += weight * c[1]; b += weight * c[2];
            t += dl[1];
            d += weight * t;
            ws += weight;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:602-682; grad buffers must be zero-filled by the caller */
void oracle_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                          const float* sigmas, const float* rgbs, const float* deltas,
                                          const int32_t* rays, const float* weights_sum,
                                          const float* image, uint32_t M, uint32_t N, float T_thresh,
                                          float* grad_sigmas, float* grad_rgbs) {
    for (uint32_t n = 0; n < N; ++n) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1],
                       num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps > M) continue;
        const float gws = grad_weights_sum[index];
        const float* gi = grad_image + (size_t)index * 3;
        const float* s = sigmas + offset; const float* c = rgbs + (size_t)offset * 3;
        const float* dl = deltas + (size_t)offset * 2;
        float* gs = grad_sigmas + offset; float* gc = grad_rgbs + (size_t)offset * 3;
        const float r_final = image[index * 3], g_final = image[index * 3 + 1],
                    b_final = image[index * 3 + 2], ws_final = weights_sum[index];
        uint32_t step = 0;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        while (step < num_steps) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float weight = alpha * T;
            r += weight * c[0]; g += weight * c[1]; b += weight * c[2];
            ws += weight;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * weight; gc[1] = gi[1] * weight; gc[2] = gi[2] * weight;
            gs[0] = dl[0] * (gi[0] * (T * c[0] - (r_final - r)) + gi[1] * (T * c[1] - (g_final - g)) +
                             gi[2] * (T * c[2] - (b_final - b)) + gws * (1 - ws_final));
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; gs++; gc += 3; step++;
        }
        (void)ws;
    }
}

/* raymarching.cu:701-805; xyzs/dirs/deltas must be zero-filled by the caller */
void oracle_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                       const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                       uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                       const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                       const float* noises) {
    (void)nears;
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        const float noise = noises ? noises[n] : 0.0f;
        march_t m;
        march_setup(&m, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, grid, bound, dt_gamma,
                    max_steps, C, H);
        float* px = xyzs + (size_t)n * n_step * 3;
        float* pd = dirs + (size_t)n * n_step * 3;
        float* pl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        const float far = fars[index];
        uint32_t step = 0;
        t = fmaf(clampf(t * dt_gamma, m.dt_min, m.dt_max), noise, t);
        float last_t = t, x, y, z, dt;
        while (t < far && step < n_step) {
            if (march_step(&m, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = m.dx; pd[1] = m.dy; pd[2] = m.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* raymarching.cu:819-905 */
void oracle_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                           float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                           float* weights_sum, float* depth, float* image) {
    for (uint32_t n = 0; n < n_alive; ++n) {
        const int32_t index = rays_alive[n];
        const float* s = sigmas + (size_t)n * n_step; const float* c = rgbs + (size_t)n * n_step * 3;
        const float* dl = deltas + (size_t)n * n_step * 2;
        float t = rays_t[index];
        float weight_sum = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += dl[1];
            d += weight * t;
            r += weight * c[0]; g += weight * c[1]; b += weight * c[2];
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; step++;
        }
        if (step < n_step) rays_alive[n] = -1;
        else rays_t[index] = t;
        weights_sum[index] = weight_sum; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}
