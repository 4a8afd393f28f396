// raymarch.cu — occupancy-grid ray marching and compositing for B200 (sm_100a).
//
// Replaces the reference's raymarching/src/raymarching.cu (10 thread-per-ray kernels) with a
// warp-per-ray design:
//   * march: a warp tests 32 lattice points of one ray per iteration (the sequence
//     of candidate t values is a fixed per-ray lattice), resolves which of them the reference's
//     sequential loop would actually visit with a pointer-doubling pass over warp shuffles, and
//     compacts the occupied ones with ballot/popc.  Sample offsets come from a deterministic
//     prefix sum over rays instead of the reference's atomicAdd (raymarching.cu:405-406).
//   * composite: a warp owns a ray, streams 32 samples per iteration with coalesced loads and
//     carries transmittance / colour with warp scans; early termination is decided per chunk.
//
// Bit-exactness: every float operation that feeds an integer decision (grid cell, occupancy,
// skip distance, loop exit) is written with explicit __f*_rn intrinsics in the operation order
// and FMA contraction the reference build produces (nvcc default -fmad=true, IEEE div), so the
// compiler cannot re-associate or re-contract it.
#include "march_common.cuh"
#include <cstdlib>

namespace nsig {


__global__ void __launch_bounds__(256)
k_near_far(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
           const float* __restrict__ aabb, uint32_t N, float min_near,
           float* __restrict__ nears, float* __restrict__ fars) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    float nr, fr;
    near_far_one(rays_o[n * 3], rays_o[n * 3 + 1], rays_o[n * 3 + 2], rays_d[n * 3],
                 rays_d[n * 3 + 1], rays_d[n * 3 + 2], aabb, min_near, nr, fr);
    nears[n] = nr;
    fars[n] = fr;
}

// ---------------------------------------------------------------------------------------
// K2 sph_from_ray — reference raymarching.cu:163-198 (API compatibility; bg_radius > 0 only)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius,
               uint32_t N, float* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float C = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * C)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[n * 2] = 2 * theta * kRPi - 1;
    coords[n * 2 + 1] = phi * kRPi;
}

// ---------------------------------------------------------------------------------------
// K3/K4 Morton — reference raymarching.cu:214-254
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_morton3D(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    indices[n] = (int)morton3D(coords[n * 3], coords[n * 3 + 1], coords[n * 3 + 2]);
}

__global__ void __launch_bounds__(256)
k_morton3D_invert(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= N) return;
    const int ind = indices[n];  // arithmetic shifts of the signed value, as in the reference
    coords[n * 3] = (int)morton3D_invert((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int)morton3D_invert((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int)morton3D_invert((uint32_t)(ind >> 2));
}

// ---------------------------------------------------------------------------------------
// K5 packbits — reference raymarching.cu:268-289.  One thread packs 4 output bytes from
// 8 float4 loads (128 contiguous bytes per thread) and stores them with one 32-bit store.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_packbits(const float* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t q = threadIdx.x + blockIdx.x * blockDim.x;  // group of 4 bytes
    const uint32_t n0 = q * 4;
    if (n0 >= N) return;
    const bool vec_ok = (n0 + 4 <= N) && ((((uintptr_t)grid) & 15) == 0) && ((((uintptr_t)bitfield) & 3) == 0);
    if (vec_ok) {
        const float4* g4 = reinterpret_cast<const float4*>(grid + (size_t)n0 * 8);
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float4 lo = ld_stream4(g4 + 2 * b), hi = ld_stream4(g4 + 2 * b + 1);
            uint32_t bits = (lo.x > thresh ? 1u : 0u) | (lo.y > thresh ? 2u : 0u) |
                            (lo.z > thresh ? 4u : 0u) | (lo.w > thresh ? 8u : 0u) |
                            (hi.x > thresh ? 16u : 0u) | (hi.y > thresh ? 32u : 0u) |
                            (hi.z > thresh ? 64u : 0u) | (hi.w > thresh ? 128u : 0u);
            word |= bits << (8 * b);
        }
        *reinterpret_cast<uint32_t*>(bitfield + n0) = word;
    } else {
        for (uint32_t n = n0; n < N && n < n0 + 4; ++n) {
            uint32_t bits = 0;
            for (int i = 0; i < 8; ++i) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (1u << i) : 0u;
            bitfield[n] = (uint8_t)bits;
        }
    }
}


constexpr int kMarchWarps = 8;  // rays per CTA

// pass 1: per-ray sample counts.  Warp-per-ray; the grid may be smaller than N / kMarchWarps (nsig_march_rays_train_limited:
// a march that shares the SMs with other kernels), then every warp walks rays n, n + stride, ...
__global__ void __launch_bounds__(kMarchWarps * 32)
k_march_count(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
              const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps,
              uint32_t N, uint32_t C, uint32_t H, const float* __restrict__ nears,
              const float* __restrict__ fars, const float* __restrict__ noises,
              uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const MarchCfg c = make_cfg(bound, dt_gamma, max_steps, C, H);
    for (uint32_t n = blockIdx.x * kMarchWarps + (threadIdx.x >> 5); n < N; n += gridDim.x * kMarchWarps) {
        const RayConst r = load_ray(rays_o, rays_d, n);
        const float t0 = perturbed_start(nears[n], noises ? noises[n] : 0.0f, c);
        const uint32_t cnt = warp_march<false>(r, c, grid, t0, fars[n], max_steps, nullptr, nullptr, nullptr, lane);
        if (lane == 0) counts[n] = cnt;
    }
}

// exclusive scan of counts (single CTA of 1024 threads); updates the global counters the way the
// reference's atomics do: counter[0] += sum(counts), counter[1] += N.
__global__ void __launch_bounds__(1024)
k_march_scan(const uint32_t* __restrict__ counts, uint32_t N, uint32_t* __restrict__ offsets,
             int* __restrict__ counter) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t running;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = (uint32_t)counter[0];
    __syncthreads();
    for (uint32_t base = 0; base < N; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < N) ? counts[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(NSIG_FULL_MASK, s, d);
            if (lane >= d) s += o;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(NSIG_FULL_MASK, ws, d);
                if (lane >= d) ws += o;
            }
            warp_sums[lane] = ws;  // inclusive
        }
        __syncthreads();
        const uint32_t warp_excl = (wid == 0) ? 0u : warp_sums[wid - 1];
        const uint32_t start = running;
        if (i < N) offsets[i] = start + warp_excl + s - v;
        __syncthreads();
        if (threadIdx.x == 0) running = start + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counter[0] = (int)running;
        counter[1] += (int)N;
    }
}

// pass 2: write samples and the rays table (id, offset, count); same ray-to-warp mapping as pass 1
__global__ void __launch_bounds__(kMarchWarps * 32)
k_march_write(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
              const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps,
              uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* __restrict__ nears,
              const float* __restrict__ fars, const float* __restrict__ noises,
              const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
              float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
              int* __restrict__ rays) {
    const int lane = threadIdx.x & 31;
    const MarchCfg c = make_cfg(bound, dt_gamma, max_steps, C, H);
    for (uint32_t n = blockIdx.x * kMarchWarps + (threadIdx.x >> 5); n < N; n += gridDim.x * kMarchWarps) {
        const uint32_t num_steps = counts[n], off = offsets[n];
        if (lane == 0) {
            rays[n * 3] = (int)n;
            rays[n * 3 + 1] = (int)off;
            rays[n * 3 + 2] = (int)num_steps;
        }
        if (num_steps == 0) continue;
        if (off + num_steps > M) {
            // reservation overflow: ray dropped (raymarching.cu:416).  Offsets only grow, so exactly one
            // dropped ray starts inside the buffer; it clears the tail the reference leaves zero-filled.
            for (uint32_t i = off + lane; i < M; i += 32) {
                xyzs[(size_t)i * 3] = xyzs[(size_t)i * 3 + 1] = xyzs[(size_t)i * 3 + 2] = 0.f;
                dirs[(size_t)i * 3] = dirs[(size_t)i * 3 + 1] = dirs[(size_t)i * 3 + 2] = 0.f;
                deltas[(size_t)i * 2] = deltas[(size_t)i * 2 + 1] = 0.f;
            }
            continue;
        }
        const RayConst r = load_ray(rays_o, rays_d, n);
        const float t0 = perturbed_start(nears[n], noises ? noises[n] : 0.0f, c);
        warp_march<true>(r, c, grid, t0, fars[n], num_steps, xyzs + (size_t)off * 3,
                         dirs + (size_t)off * 3, deltas + (size_t)off * 2, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-launch march: count -> offsets -> write in ONE kernel.  A CTA (8 rays) counts its rays, publishes the CTA total,
// obtains its exclusive prefix over all earlier CTAs by decoupled look-back (aggregate / inclusive-prefix words, one warp
// looking back 32 CTAs at a time) and marches its rays again to write them - the second walk re-reads occupancy bytes the
// first one just pulled into L1.  Offsets are the same exclusive scan in ray order as k_march_scan computes, so rays /
// xyzs / dirs / deltas / counter are bit-identical to the three-kernel path; what goes away is the single-CTA scan kernel
// and two kernel boundaries in a chain that sits on the step's critical path.  CTA ids are drawn from a ticket so that a
// CTA only ever waits for CTAs that started before it.
// state: uint64[n_blocks] (flag << 32 | value; flag 1 = CTA total, 2 = inclusive prefix), ticket: uint32 - zeroed by the host.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMarchWarps * 32)
k_march_fused(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
              float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
              const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
              unsigned long long* __restrict__ state, uint32_t* __restrict__ ticket, float* __restrict__ xyzs,
              float* __restrict__ dirs, float* __restrict__ deltas, int* __restrict__ rays, int* __restrict__ counter) {
    __shared__ uint32_t s_vb, s_cnt[kMarchWarps], s_excl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_vb = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t vb = s_vb, n_blocks = gridDim.x;
    const uint32_t n = vb * kMarchWarps + wid;
    const bool live = n < N;
    const MarchCfg c = make_cfg(bound, dt_gamma, max_steps, C, H);
    RayConst r{};
    float t0 = 0.f, far = 0.f;
    uint32_t cnt = 0;
    if (live) {
        r = load_ray(rays_o, rays_d, n);
        t0 = perturbed_start(nears[n], noises ? noises[n] : 0.0f, c);
        far = fars[n];
        cnt = warp_march<false>(r, c, grid, t0, far, max_steps, nullptr, nullptr, nullptr, lane);
    }
    if (lane == 0) s_cnt[wid] = cnt;
    __syncthreads();
    if (wid == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kMarchWarps; ++w) total += s_cnt[w];
        uint32_t excl = 0;
        if (vb == 0) {
            excl = (uint32_t)counter[0];                      // the reference's running sample counter (normally 0)
            if (lane == 0) {
                __threadfence();
                atomicExch(state, (2ull << 32) | (unsigned long long)(excl + total));
            }
        } else {
            if (lane == 0) {
                __threadfence();
                atomicExch(state + vb, (1ull << 32) | (unsigned long long)total);
            }
            // look back: lanes inspect CTAs vb-1-lane of the current window; stop at the first inclusive prefix
            int base = (int)vb - 1;
            while (true) {
                const int idx = base - lane;
                unsigned long long w = idx >= 0 ? *((volatile unsigned long long*)(state + idx)) : (2ull << 32);
                const uint32_t flag = (uint32_t)(w >> 32);
                const uint32_t ready = __ballot_sync(NSIG_FULL_MASK, flag != 0u);
                const uint32_t pref = __ballot_sync(NSIG_FULL_MASK, flag == 2u);
                // usable lanes: the contiguous run of published entries from lane 0 up to (and including) the first prefix
                const uint32_t first_gap = ~ready ? (uint32_t)__ffs(~ready) - 1u : 32u;
                const uint32_t first_pref = pref ? (uint32_t)__ffs(pref) - 1u : 32u;
                if (first_pref < first_gap) {                 // a prefix is reachable through published aggregates
                    uint32_t v = (uint32_t)lane <= first_pref ? (uint32_t)w : 0u;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(NSIG_FULL_MASK, v, d);
                    excl += v;
                    break;
                }
                if (first_gap == 32u) {                       // 32 aggregates, no prefix yet: take them and move on
                    uint32_t v = (uint32_t)w;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(NSIG_FULL_MASK, v, d);
                    excl += v;
                    base -= 32;
                }
                // else: an earlier CTA has not published yet - poll the same window again
            }
            if (lane == 0) {
                __threadfence();
                atomicExch(state + vb, (2ull << 32) | (unsigned long long)(excl + total));
            }
        }
        if (lane == 0) {
            s_excl = excl;
            if (vb == n_blocks - 1) {                          // the reference's atomics: counter[0] += samples, counter[1] += N
                counter[0] = (int)(excl + total);
                counter[1] += (int)N;
            }
        }
    }
    __syncthreads();
    if (!live) return;
    uint32_t off = s_excl;
    for (int w = 0; w < wid; ++w) off += s_cnt[w];
    if (lane == 0) {
        rays[n * 3] = (int)n;
        rays[n * 3 + 1] = (int)off;
        rays[n * 3 + 2] = (int)cnt;
    }
    if (cnt == 0) return;
    if (off + cnt > M) {   // reservation overflow: see k_march_write
        for (uint32_t i = off + lane; i < M; i += 32) {
            xyzs[(size_t)i * 3] = xyzs[(size_t)i * 3 + 1] = xyzs[(size_t)i * 3 + 2] = 0.f;
            dirs[(size_t)i * 3] = dirs[(size_t)i * 3 + 1] = dirs[(size_t)i * 3 + 2] = 0.f;
            deltas[(size_t)i * 2] = deltas[(size_t)i * 2 + 1] = 0.f;
        }
        return;
    }
    warp_march<true>(r, c, grid, t0, far, cnt, xyzs + (size_t)off * 3, dirs + (size_t)off * 3, deltas + (size_t)off * 2, lane);
}

// zero rows [counter[0], end) with end = min(align_up(counter[0]), M), or M when align == 0
__global__ void __launch_bounds__(256)
k_zero_padding(float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
               const int* __restrict__ counter, uint32_t align, uint32_t M) {
    const uint32_t m = (uint32_t)counter[0];
    uint32_t end = M;
    if (align) { end = m + align - m % align; if (end > M) end = M; }
    for (uint32_t i = m + threadIdx.x + blockIdx.x * blockDim.x; i < end; i += blockDim.x * gridDim.x) {
        xyzs[(size_t)i * 3] = xyzs[(size_t)i * 3 + 1] = xyzs[(size_t)i * 3 + 2] = 0.f;
        dirs[(size_t)i * 3] = dirs[(size_t)i * 3 + 1] = dirs[(size_t)i * 3 + 2] = 0.f;
        deltas[(size_t)i * 2] = deltas[(size_t)i * 2 + 1] = 0.f;
    }
}

// K9 inference march — reference raymarching.cu:701-805
__global__ void __launch_bounds__(kMarchWarps * 32)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
             const float* __restrict__ rays_t, const float* __restrict__ rays_o,
             const float* __restrict__ rays_d, float bound, float dt_gamma, uint32_t max_steps,
             uint32_t C, uint32_t H, const uint8_t* __restrict__ grid,
             const float* __restrict__ nears, const float* __restrict__ fars,
             float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
             const float* __restrict__ noises) {
    const uint32_t n = blockIdx.x * kMarchWarps + (threadIdx.x >> 5);
    if (n >= n_alive) return;
    const int lane = threadIdx.x & 31;
    const int index = rays_alive[n];
    const MarchCfg c = make_cfg(bound, dt_gamma, max_steps, C, H);
    const RayConst r = load_ray(rays_o, rays_d, (uint32_t)index);
    const float t0 = perturbed_start(rays_t[index], noises ? noises[n] : 0.0f, c);
    const size_t row = (size_t)n * n_step;
    warp_march<true>(r, c, grid, t0, fars[index], n_step, xyzs + row * 3, dirs + row * 3,
                     deltas + row * 2, lane);
}

// ---------------------------------------------------------------------------------------
// K7 composite forward (training) — reference raymarching.cu:501-577, warp per ray
// ---------------------------------------------------------------------------------------
constexpr int kCompWarps = 8;

// Optional epilogue of run_cuda's training branch (renderer_wtmk.py:298-303), one rounded fp32 op per torch op:
//   image = image + (1 - weights_sum)[:, None] * bg_color        (scalar bg_color)
//   depth = clamp(depth - nears, min=0) / (fars - nears)
struct CompositeBlend {
    float* image_out;      // [N,3] or null (no epilogue)
    float* depth_out;      // [N]
    const float* nears;    // [N]
    const float* fars;     // [N]
    float bg;
};

__device__ __forceinline__ void blend_outputs(const CompositeBlend& bl, uint32_t index, float ws, float d, float r,
                                              float g, float b) {
    const float k = __fmul_rn(__fsub_rn(1.0f, ws), bl.bg);
    bl.image_out[index * 3] = __fadd_rn(r, k);
    bl.image_out[index * 3 + 1] = __fadd_rn(g, k);
    bl.image_out[index * 3 + 2] = __fadd_rn(b, k);
    const float near = bl.nears[index], far = bl.fars[index];
    bl.depth_out[index] = __fdiv_rn(fmaxf(__fsub_rn(d, near), 0.0f), __fsub_rn(far, near));
}


__global__ void __launch_bounds__(kCompWarps * 32)
k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int* __restrict__ rays, uint32_t M,
                      uint32_t N, float T_thresh, float* __restrict__ weights_sum,
                      float* __restrict__ depth, float* __restrict__ image, CompositeBlend bl) {
    const uint32_t n = blockIdx.x * kCompWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1],
                   num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) {
        if (lane == 0) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = 0; image[index * 3 + 1] = 0; image[index * 3 + 2] = 0;
            if (bl.image_out) blend_outputs(bl, index, 0.f, 0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    float T = 1.0f, t_acc = 0.f;           // transmittance / accumulated real delta before the chunk
    float r = 0, g = 0, b = 0, ws = 0, d = 0;  // per-lane partial sums
    // the loads of chunk k+1 are issued before the scans of chunk k (a ray is a serial chain of ~10 chunks, each of
    // which would otherwise wait for its own L2/HBM round trip)
    float n_sigma = 0.f, n_r = 0.f, n_g = 0.f, n_b = 0.f;
    float2 n_dl = make_float2(0.f, 0.f);
    auto fetch = [&](uint32_t base) {
        const uint32_t s = base + lane;
        n_sigma = 0.f; n_dl = make_float2(0.f, 0.f); n_r = n_g = n_b = 0.f;
        if (s < num_steps) {
            const size_t i = (size_t)offset + s;
            n_sigma = sigmas[i];
            n_dl = *reinterpret_cast<const float2*>(deltas + i * 2);
            n_r = rgbs[i * 3]; n_g = rgbs[i * 3 + 1]; n_b = rgbs[i * 3 + 2];
        }
    };
    fetch(0);
    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t s = base + lane;
        const bool valid = s < num_steps;
        const float sigma = n_sigma, cr = n_r, cg = n_g, cb = n_b;
        const float2 dl = n_dl;
        if (base + 32 < num_steps) fetch(base + 32);
        const float alpha = 1.0f - __expf(-sigma * dl.x);
        const float one_m = valid ? (1.0f - alpha) : 1.0f;
        const float Tincl = T * warp_scan_mul(one_m, lane);       // T after this sample
        float Tbefore = __shfl_up_sync(NSIG_FULL_MASK, Tincl, 1);
        if (lane == 0) Tbefore = T;
        const float tcum = t_acc + warp_scan_add(dl.y, lane);
        // the reference accumulates sample k iff every earlier sample left T >= T_thresh
        const uint32_t dead = __ballot_sync(NSIG_FULL_MASK, valid && (Tincl < T_thresh));
        const int stop = dead ? (__ffs(dead) - 1) : 32;  // first lane that triggers the break
        if (valid && lane <= stop) {
            const float w = alpha * Tbefore;
            r += w * cr; g += w * cg; b += w * cb;
            d += w * tcum;
            ws += w;
        }
        if (dead) break;
        T = __shfl_sync(NSIG_FULL_MASK, Tincl, 31);
        t_acc = __shfl_sync(NSIG_FULL_MASK, tcum, 31);
    }
    r = warp_sum(r); g = warp_sum(g); b = warp_sum(b); ws = warp_sum(ws); d = warp_sum(d);
    if (lane == 0) {
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
        if (bl.image_out) blend_outputs(bl, index, ws, d, r, g, b);
    }
}

// ---------------------------------------------------------------------------------------
// K8 composite backward (training) — reference raymarching.cu:602-682, warp per ray
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCompWarps * 32)
k_composite_train_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image,
                      const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, const int* __restrict__ rays,
                      const float* __restrict__ weights_sum, const float* __restrict__ image,
                      uint32_t M, uint32_t N, float T_thresh, float* __restrict__ grad_sigmas,
                      float* __restrict__ grad_rgbs, float blend_bg) {
    const uint32_t n = blockIdx.x * kCompWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1],
                   num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    const float gr = grad_image[index * 3], gg = grad_image[index * 3 + 1], gb = grad_image[index * 3 + 2];
    // with the fused background blend grad_image is the gradient of image + (1 - ws) * bg: d/d ws gains -bg * sum(grad)
    const float gws = (grad_weights_sum ? grad_weights_sum[index] : 0.0f) - blend_bg * (gr + gg + gb);
    const float r_final = image[index * 3], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2];
    const float ws_term = gws * (1.0f - weights_sum[index]);
    float T = 1.0f, r0 = 0.f, g0 = 0.f, b0 = 0.f;  // state before the chunk
    float n_sigma = 0.f, n_d0 = 0.f, n_r = 0.f, n_g = 0.f, n_b = 0.f;   // next chunk's inputs (see the forward kernel)
    auto fetch = [&](uint32_t base) {
        const uint32_t s = base + lane;
        n_sigma = n_d0 = n_r = n_g = n_b = 0.f;
        if (s < num_steps) {
            const size_t i = (size_t)offset + s;
            n_sigma = sigmas[i];
            n_d0 = deltas[i * 2];
            n_r = rgbs[i * 3]; n_g = rgbs[i * 3 + 1]; n_b = rgbs[i * 3 + 2];
        }
    };
    fetch(0);
    for (uint32_t base = 0; base < num_steps; base += 32) {
        const uint32_t s = base + lane;
        const bool valid = s < num_steps;
        const size_t i = (size_t)offset + s;
        const float sigma = n_sigma, d0 = n_d0, cr = n_r, cg = n_g, cb = n_b;
        if (base + 32 < num_steps) fetch(base + 32);
        const float alpha = 1.0f - __expf(-sigma * d0);
        const float one_m = valid ? (1.0f - alpha) : 1.0f;
        const float Tincl = T * warp_scan_mul(one_m, lane);
        float Tbefore = __shfl_up_sync(NSIG_FULL_MASK, Tincl, 1);
        if (lane == 0) Tbefore = T;
        const float w = valid ? alpha * Tbefore : 0.f;
        const float racc = r0 + warp_scan_add(w * cr, lane);   // inclusive running colour
        const float gacc = g0 + warp_scan_add(w * cg, lane);
        const float bacc = b0 + warp_scan_add(w * cb, lane);
        const uint32_t dead = __ballot_sync(NSIG_FULL_MASK, valid && (Tincl < T_thresh));
        const int stop = dead ? (__ffs(dead) - 1) : 32;
        if (valid && lane <= stop) {
            grad_rgbs[i * 3] = gr * w; grad_rgbs[i * 3 + 1] = gg * w; grad_rgbs[i * 3 + 2] = gb * w;
            grad_sigmas[i] = d0 * (gr * (Tincl * cr - (r_final - racc)) +
                                   gg * (Tincl * cg - (g_final - gacc)) +
                                   gb * (Tincl * cb - (b_final - bacc)) + ws_term);
        }
        if (dead) {
            // samples after the terminating one get zero gradient (the reference relies on the caller's
            // zero-fill, raymarching.py:283-284; writing them here makes that fill unnecessary)
            for (uint32_t z = base + stop + 1 + lane; z < num_steps; z += 32) {
                const size_t zi = (size_t)offset + z;
                grad_sigmas[zi] = 0.f;
                grad_rgbs[zi * 3] = 0.f; grad_rgbs[zi * 3 + 1] = 0.f; grad_rgbs[zi * 3 + 2] = 0.f;
            }
            break;
        }
        T = __shfl_sync(NSIG_FULL_MASK, Tincl, 31);
        r0 = __shfl_sync(NSIG_FULL_MASK, racc, 31);
        g0 = __shfl_sync(NSIG_FULL_MASK, gacc, 31);
        b0 = __shfl_sync(NSIG_FULL_MASK, bacc, 31);
    }
}

// ---------------------------------------------------------------------------------------
// K10 inference composite — reference raymarching.cu:819-905 (n_step <= 8: thread per ray)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* __restrict__ rays_alive,
                 float* __restrict__ rays_t, const float* __restrict__ sigmas,
                 const float* __restrict__ rgbs, const float* __restrict__ deltas,
                 float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image) {
    const uint32_t n = threadIdx.x + blockIdx.x * blockDim.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    sigmas += (size_t)n * n_step;
    rgbs += (size_t)n * n_step * 3;
    deltas += (size_t)n * n_step * 2;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        if (deltas[0] == 0) break;  // ray ran out of samples
        const float alpha = 1.0f - __expf(-sigmas[0] * deltas[0]);
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t += deltas[1];
        d += weight * t;
        r += weight * rgbs[0]; g += weight * rgbs[1]; b += weight * rgbs[2];
        if (T < T_thresh) break;
        sigmas++; rgbs += 3; deltas += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = weight_sum; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

}  // namespace nsig

// =======================================================================================
// C ABI
// =======================================================================================
using namespace nsig;

extern "C" {

int nsig_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                            float min_near, float* nears, float* fars, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !aabb || !nears || !fars) return NSIG_EINVAL;
    k_near_far<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                      nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !coords) return NSIG_EINVAL;
    k_sph_from_ray<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!coords || !indices) return NSIG_EINVAL;
    k_morton3D<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!coords || !indices) return NSIG_EINVAL;
    k_morton3D_invert<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield,
                  nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!grid || !bitfield) return NSIG_EINVAL;
    k_packbits<<<div_up(div_up(N, 4), 256), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
    NSIG_LAUNCH_CHECK();
    return 0;
}

size_t nsig_march_rays_train_scratch_bytes(uint32_t N) { return (size_t)N * 2 * sizeof(uint32_t); }

static int march_rays_train_impl(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                 uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                                 float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                 void* scratch, uint32_t max_blocks, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !grid || !nears || !fars || !xyzs || !dirs || !deltas || !rays ||
        !counter || !scratch)
        return NSIG_EINVAL;
    if (C == 0 || C > 31 || H == 0 || H > 1024 || max_steps == 0) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* counts = (uint32_t*)scratch;
    uint32_t* offsets = counts + N;
    const uint32_t all_blocks = div_up(N, kMarchWarps);
    const uint32_t blocks = (max_blocks && max_blocks < all_blocks) ? max_blocks : all_blocks;
    // Measured (round 2, call Q): inside the training step the single-launch kernel is SLOWER than the three-kernel chain
    // (1.000 vs 0.967 ms per step; the march branch takes ~220 us either way next to the HBM-bound table Adam, and CTAs that
    // spin in the look-back keep SM slots the Adam and the remaining rays are waiting for), so it is opt-in: NSIG_MARCH_FUSED=1.
    static const bool fused = [] { const char* e = getenv("NSIG_MARCH_FUSED"); return e && e[0] == '1'; }();
    if (fused && blocks == all_blocks) {
        // scratch (2N words) as [ticket | pad | look-back state: one 64-bit word per CTA]
        const size_t state_bytes = 16 + (size_t)blocks * sizeof(unsigned long long);
        if (state_bytes <= (size_t)N * 2 * sizeof(uint32_t) && (((uintptr_t)scratch) & 7) == 0) {
            cudaError_t e = cudaMemsetAsync(scratch, 0, state_bytes, st);
            if (e != cudaSuccess) return (int)e;
            k_march_fused<<<blocks, kMarchWarps * 32, 0, st>>>(
                rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, noises,
                reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(scratch) + 16), counts, xyzs, dirs,
                deltas, rays, counter);
            NSIG_LAUNCH_CHECK();
            return 0;
        }
    }
    k_march_count<<<blocks, kMarchWarps * 32, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N,
                                                       C, H, nears, fars, noises, counts);
    NSIG_LAUNCH_CHECK();
    k_march_scan<<<1, 1024, 0, st>>>(counts, N, offsets, counter);
    NSIG_LAUNCH_CHECK();
    k_march_write<<<blocks, kMarchWarps * 32, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N,
                                                       C, H, M, nears, fars, noises, counts, offsets,
                                                       xyzs, dirs, deltas, rays);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                          float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                          uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                          float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                          void* scratch, nsig_stream_t stream) {
    return march_rays_train_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
                                 deltas, rays, counter, noises, scratch, 0, stream);
}

int nsig_march_rays_train_limited(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                  float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                  uint32_t M, const float* nears, const float* fars, float* xyzs, float* dirs,
                                  float* deltas, int32_t* rays, int32_t* counter, const float* noises,
                                  void* scratch, uint32_t max_blocks, nsig_stream_t stream) {
    return march_rays_train_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs,
                                 deltas, rays, counter, noises, scratch, max_blocks, stream);
}

int nsig_zero_sample_padding(float* xyzs, float* dirs, float* deltas, const int32_t* counter,
                             uint32_t align, uint32_t M, nsig_stream_t stream) {
    if (!xyzs || !dirs || !deltas || !counter) return NSIG_EINVAL;
    const uint32_t span = align ? align : M;
    const uint32_t blocks = div_up(span, 256) < 1184u ? div_up(span, 256) : 1184u;
    if (blocks == 0) return 0;
    k_zero_padding<<<blocks, 256, 0, (cudaStream_t)stream>>>(xyzs, dirs, deltas, counter, align, M);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                      const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                      float* weights_sum, float* depth, float* image,
                                      nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!sigmas || !rgbs || !deltas || !rays || !weights_sum || !depth || !image) return NSIG_EINVAL;
    k_composite_train_fwd<<<div_up(N, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image, CompositeBlend{});
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_composite_rays_train_blend_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                            const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                            float bg_color, const float* nears, const float* fars,
                                            float* weights_sum, float* depth, float* image, float* image_out,
                                            float* depth_out, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!sigmas || !rgbs || !deltas || !rays || !weights_sum || !depth || !image || !nears || !fars || !image_out ||
        !depth_out)
        return NSIG_EINVAL;
    k_composite_train_fwd<<<div_up(N, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image,
        CompositeBlend{image_out, depth_out, nears, fars, bg_color});
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_composite_rays_train_blend_backward(const float* grad_weights_sum, const float* grad_image_out,
                                             const float* sigmas, const float* rgbs, const float* deltas,
                                             const int32_t* rays, const float* weights_sum, const float* image,
                                             uint32_t M, uint32_t N, float T_thresh, float bg_color,
                                             float* grad_sigmas, float* grad_rgbs, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!grad_image_out || !sigmas || !rgbs || !deltas || !rays || !weights_sum || !image || !grad_sigmas || !grad_rgbs)
        return NSIG_EINVAL;
    k_composite_train_bwd<<<div_up(N, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        grad_weights_sum, grad_image_out, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh,
        grad_sigmas, grad_rgbs, bg_color);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                       const float* sigmas, const float* rgbs, const float* deltas,
                                       const int32_t* rays, const float* weights_sum, const float* image,
                                       uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                       float* grad_rgbs, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!grad_weights_sum || !grad_image || !sigmas || !rgbs || !deltas || !rays || !weights_sum ||
        !image || !grad_sigmas || !grad_rgbs)
        return NSIG_EINVAL;
    k_composite_train_bwd<<<div_up(N, kCompWarps), kCompWarps * 32, 0, (cudaStream_t)stream>>>(
        grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh,
        grad_sigmas, grad_rgbs, 0.0f);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma,
                    uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid, const float* nears,
                    const float* fars, float* xyzs, float* dirs, float* deltas, const float* noises,
                    nsig_stream_t stream) {
    if (n_alive == 0 || n_step == 0) return 0;
    if (!rays_alive || !rays_t || !rays_o || !rays_d || !grid || !nears || !fars || !xyzs || !dirs || !deltas)
        return NSIG_EINVAL;
    if (C == 0 || C > 31 || H == 0 || H > 1024 || max_steps == 0) return NSIG_EINVAL;
    k_march_rays<<<div_up(n_alive, kMarchWarps), kMarchWarps * 32, 0, (cudaStream_t)stream>>>(
        n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears,
        fars, xyzs, dirs, deltas, noises);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs, const float* deltas,
                        float* weights_sum, float* depth, float* image, nsig_stream_t stream) {
    if (n_alive == 0) return 0;
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image)
        return NSIG_EINVAL;
    k_composite_rays<<<div_up(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(
        n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
