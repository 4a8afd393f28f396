// field.cu — fused NeRFNetwork.forward / density and the watermark-mode backward for B200.
//
// Reference path (nerf/network_wtmk_tcnn.py:97-124): normalise x -> 16-level hash encode
// (hash_encoding.py) -> add message feature to channels 30,31 -> sigma MLP 32-64-16 ->
// trunc_exp -> SH4(d) ++ geo(15) -> colour MLP 32-64-64-16 -> sigmoid; ~400 torch kernels, a
// [M,32] fp32 feature tensor and several [M,16..64] activations round-tripping through HBM.
//
// Here one kernel does all of it.  A warp owns 32 samples (two 16-row MMA tiles).  The thread
// with quad coordinates (g, tig) gathers exactly the levels {tig, tig+4, tig+8, tig+12} of rows
// {g, g+8} of each tile, which is precisely the set of elements it owns in the m16n8k16
// A-fragment, so hash features go from the gather straight into tensor-core operand registers
// - no shared-memory transpose, no feature tensor in HBM.  Accumulator fragments of one layer
// are re-packed (ReLU, fp32->fp16) into the A fragments of the next layer in registers; weights
// (fp16, 10 240 values) sit in padded shared memory and every B fragment is reused by both row
// tiles.  The sigma net's output rows are permuted to [geo0..geo14, logit] so the geo features
// land in the colour net's second k-step without any data movement.
//
// Precision: fp16 operands, fp32 accumulation (tcnn's FullyFusedMLP accumulates in fp16 - the
// reference's own MLP arithmetic is unpinned); encoder arithmetic is the exact
// fp32 restatement in hash_common.cuh.
#include "field_common.cuh"

namespace nsig {


// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <bool COLOR, bool H2>
__global__ void __launch_bounds__(kFieldThreads, NSIG_FWD_MINB)
k_field_fwd(const FieldParams p, float* __restrict__ sigmas, float* __restrict__ rgbs, __half* __restrict__ feat_out,
            __half* __restrict__ geo_out, uint2* __restrict__ mask_out) {
    constexpr int MT = 2;
    extern __shared__ __align__(16) __half sm[];
    uint32_t M = p.M;
    if (p.M_dev) M = min(M, (uint32_t)max(*p.M_dev, 0));
    if (M == 0) return;
    stage_forward_weights(sm, p.sigma_w, p.color_w, COLOR);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    [[maybe_unused]] uint32_t* feat_s = reinterpret_cast<uint32_t*>(sm + kFwdHalfsPad) + warp * kFeatWords;
    const uint32_t rows_per_cta = kFieldWarps * 16 * MT;
    const uint32_t n_tiles = div_up(M, rows_per_cta);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row0 = tile * rows_per_cta + warp * 16 * MT;
        if (row0 >= M) continue;
        uint32_t fa[MT][2][4];
#ifndef NSIG_GATHER_V2
#ifndef NSIG_FWD_UNROLLED
        // per-row encode code emitted once and looped over the thread's 4 rows: 0.2317 -> 0.2180 ms per 1.07 M samples against the
        // fully unrolled body (-DNSIG_FWD_UNROLLED), which is instruction-fetch limited (profiles/r02_experiments.txt)
        encode_rows_rolled<MT, H2>(fa, p, M, row0, g, tig);
#else
        encode_rows<MT, H2>(fa, p, M, row0, g, tig);
#endif
        if (feat_out) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t r = frag_row<MT>(row0, mt, h, g);
                    if (r < M) {
                        uint32_t* dst = reinterpret_cast<uint32_t*>(feat_out + (size_t)r * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dst[tig + 4 * j] = fa[mt][j >> 1][2 * (j & 1) + h];
                    }
                }
        }
#else
        {
            const float* xr = p.xyzs + (size_t)min(row0 + lane, M - 1) * 3;  // x = (x + bound) / (2*bound)  (network_wtmk_tcnn.py:101)
            const float px = __fmul_rn(__fadd_rn(__ldg(xr), p.bound_add), p.bound_mul);
            const float py = __fmul_rn(__fadd_rn(__ldg(xr + 1), p.bound_add), p.bound_mul);
            const float pz = __fmul_rn(__fadd_rn(__ldg(xr + 2), p.bound_add), p.bound_mul);
            gather_tile<H2>(fa, p, px, py, pz, feat_s, lane);
        }
        if (feat_out) {  // the tile is row-major in shared memory: 64 contiguous bytes per row, 16 B per lane
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int chunk = i * 32 + lane, row = chunk >> 2, part = chunk & 3;
                if (row0 + row < M)
                    *reinterpret_cast<uint4*>(feat_out + (size_t)(row0 + row) * 32 + part * 8) =
                        *reinterpret_cast<const uint4*>(feat_s + row * kFeatStride + part * 4);
            }
        }
#endif
        // sigma net: 32 -> 64 (ReLU) -> 16
        uint32_t h1[MT][4][4];
        [[maybe_unused]] uint32_t mk[2][MT][2] = {};   // ReLU sign masks: [0] = m1s | m1c << 8, [1] = m2c (COLOR && mask_out)
        {
            float c[MT][8][4];
            layer<MT, 2, 8>(c, fa, sm + oWs0, kS32, g, tig);
            relu_to_a<MT, 8>(h1, c);
            if (COLOR && mask_out) act_mask_bits<MT>(mk[0], h1, 0);
        }
        float so[MT][2][4];
        layer<MT, 4, 2>(so, h1, sm + oWs1, kS64, g, tig);
        // logit = permuted column 15: thread tig==3, second element of n-tile 1
        if (tig == 3) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t r = frag_row<MT>(row0, mt, h, g);
                    if (r < M) sigmas[r] = __fmul_rn(p.density_scale, expf(so[mt][1][2 * h + 1]));  // trunc_exp (activation.py:8-10)
                }
        }
        if (geo_out) {  // density(): geo_feat [M,15] fp16
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t r = frag_row<MT>(row0, mt, h, g);
                    if (r < M) {
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int col = nt * 8 + 2 * tig + e;
                                if (col < 15) geo_out[(size_t)r * 15 + col] = __float2half_rn(so[mt][nt][2 * h + e]);
                            }
                    }
                }
        }
        if (COLOR) {
            uint32_t ca[MT][2][4];
            sh_rows<MT>(ca, p.dirs, M, row0, g, tig);
            geo_to_a<MT>(ca, so, tig);
            uint32_t h2[MT][4][4];
            {
                float c[MT][8][4];
                layer<MT, 2, 8>(c, ca, sm + oWc0, kS32, g, tig);
                relu_to_a<MT, 8>(h1, c);
                if (mask_out) act_mask_bits<MT>(mk[0], h1, 8);
                layer<MT, 4, 8>(c, h1, sm + oWc1, kS64, g, tig);
                relu_to_a<MT, 8>(h2, c);
                if (mask_out) act_mask_bits<MT>(mk[1], h2, 0);
            }
            if (mask_out) {   // 8 bytes per (row, quad thread): {m1s | m1c << 8, m2c}; a quad writes 32 contiguous bytes per row
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t r = frag_row<MT>(row0, mt, h, g);
                        if (r < M) mask_out[(size_t)r * 4 + tig] = make_uint2(mk[0][mt][h], mk[1][mt][h]);
                    }
            }
            float co[MT][1][4];
            layer<MT, 4, 1>(co, h2, sm + oWc2, kS64, g, tig);
            if (tig < 2) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t r = frag_row<MT>(row0, mt, h, g);
                        if (r < M) {
                            const float s0 = 1.0f / (1.0f + expf(-co[mt][0][2 * h]));
                            if (tig == 0) {
                                rgbs[(size_t)r * 3] = s0;
                                rgbs[(size_t)r * 3 + 1] = 1.0f / (1.0f + expf(-co[mt][0][2 * h + 1]));
                            } else {
                                rgbs[(size_t)r * 3 + 2] = s0;
                            }
                        }
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// colour branch only: NeRFNetwork.color(x, d, geo_feat) (network_wtmk_tcnn.py:146-176), used by the
// non-cuda_ray renderer.  geo_feat [M,15] fp16 comes from nsig_field_density.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFieldThreads)
k_color_fwd(const float* __restrict__ dirs, const __half* __restrict__ geo, uint32_t M,
            const __half* __restrict__ color_w, float* __restrict__ rgbs) {
    constexpr int MT = 2;
    extern __shared__ __align__(16) __half sm[];
    {   // only the colour matrices are needed; the sigma slots are left unused
        const __half zero = __float2half(0.0f);
        for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
            const int r = i >> 5, c = i & 31;
            sm[oWc0 + r * kS32 + c] = (c == 31) ? zero : color_w[i];
        }
        for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) sm[oWc1 + (i >> 6) * kS64 + (i & 63)] = color_w[2048 + i];
        for (int i = threadIdx.x; i < 8 * 64; i += blockDim.x) sm[oWc2 + (i >> 6) * kS64 + (i & 63)] = color_w[2048 + 4096 + i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const uint32_t rows_per_cta = kFieldWarps * 16 * MT;
    const uint32_t n_tiles = div_up(M, rows_per_cta);
    const float* dirs_ = dirs;
    const uint32_t M_ = M;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t row0 = tile * rows_per_cta + warp * 16 * MT;
        if (row0 >= M) continue;
        uint32_t ca[MT][2][4];
        sh_rows<MT>(ca, dirs_, M_, row0, g, tig);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t r = min(frag_row<MT>(row0, mt, h, g), M - 1);
                const __half* gr = geo + (size_t)r * 15;
                const __half z = __float2half(0.0f);
                __half2 lo = __halves2half2(gr[2 * tig], gr[2 * tig + 1]);
                __half2 hi = __halves2half2(gr[2 * tig + 8], (tig == 3) ? z : gr[2 * tig + 9]);
                ca[mt][1][h] = *reinterpret_cast<uint32_t*>(&lo);
                ca[mt][1][2 + h] = *reinterpret_cast<uint32_t*>(&hi);
            }
        uint32_t h1[MT][4][4], h2[MT][4][4];
        {
            float c[MT][8][4];
            layer<MT, 2, 8>(c, ca, sm + oWc0, kS32, g, tig);
            relu_to_a<MT, 8>(h1, c);
            layer<MT, 4, 8>(c, h1, sm + oWc1, kS64, g, tig);
            relu_to_a<MT, 8>(h2, c);
        }
        float co[MT][1][4];
        layer<MT, 4, 1>(co, h2, sm + oWc2, kS64, g, tig);
        if (tig < 2) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t r = frag_row<MT>(row0, mt, h, g);
                    if (r < M) {
                        const float s0 = 1.0f / (1.0f + expf(-co[mt][0][2 * h]));
                        if (tig == 0) {
                            rgbs[(size_t)r * 3] = s0;
                            rgbs[(size_t)r * 3 + 1] = 1.0f / (1.0f + expf(-co[mt][0][2 * h + 1]));
                        } else {
                            rgbs[(size_t)r * 3 + 2] = s0;
                        }
                    }
                }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Fused frame renderer: the inference branch of NeRFRenderer.run_cuda (renderer_wtmk.py:323-372) as ONE
// persistent kernel.  The reference drives march_rays -> network -> composite_rays -> alive-ray compaction
// from the host, up to ~10^3 iterations per 4096-ray chunk with a device-to-host sync each.
// Here a warp owns a ray from start to finish: it marches the occupancy grid 32 lattice points at a time
// (march_window, bit-exact lattice), stages occupied samples in shared memory, evaluates the field on 32
// samples at once with the same tensor-core path as k_field_fwd (hash gather -> A fragments -> MLPs),
// composites them with warp scans and stops at the reference's termination rule (T < T_thresh before a
// sample, renderer_wtmk.py:361 -> raymarching.cu:868-883) or when the ray leaves the box.  No sample ever
// touches HBM; rays are handed out through an atomic counter so long and short rays balance.
//
// Per-ray results equal the reference loop's up to fp32 rounding: the reference accumulates the march
// parameter through composite_rays' `t += deltas[1]` between chunks, this kernel keeps the exact lattice.
// ---------------------------------------------------------------------------------------
struct RenderParams {
    FieldParams f;  // tables, S, MLP weights, bound, density_scale (xyzs/dirs/M unused)
    const float* rays_o;
    const float* rays_d;
    uint32_t N;
    const uint8_t* grid;
    float dt_gamma;
    uint32_t max_steps, C, H;
    const float* aabb;
    float min_near, T_thresh;
    const float* noises;      // optional [N] (perturb)
    uint32_t* work_counter;   // zero-initialised by the caller
    float* weights_sum;       // [N]
    float* depth;             // [N]   sum w * t (not normalised)
    float* image;             // [N,3] (no background)
    float* nears;             // optional [N]
    float* fars;              // optional [N]
    uint32_t* sample_count;   // optional: total samples evaluated (atomicAdd)
};

constexpr int kRenderWarps = 4;
constexpr int kStageSlots = 64;
// per-warp staging (floats): xyz[64][3] + dt[64] + dreal[64] + sigma[32] + rgb[32][3]
constexpr int kStageFloats = kStageSlots * 5 + 32 * 4;

#ifndef NSIG_RENDER_MINB
#define NSIG_RENDER_MINB 4   // resident CTAs per SM the frame renderer is compiled for: 128 registers (40 B of spills) and
                             // 16 warps/SM beat 188 registers at 8 warps/SM: 800x800 frame 25.0 -> 23.5 ms
#endif
template <bool H2>
__global__ void __launch_bounds__(kRenderWarps * 32, NSIG_RENDER_MINB)
k_render_rays(const RenderParams p) {
    constexpr int MT = 2;
    extern __shared__ __align__(16) __half sm[];
    stage_forward_weights(sm, p.f.sigma_w, p.f.color_w, true);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    float* st_xyz = reinterpret_cast<float*>(sm + kFwdHalfsPad) + warp * kStageFloats;
    float* st_dt = st_xyz + kStageSlots * 3;
    float* st_dr = st_dt + kStageSlots;
    float* st_sig = st_dr + kStageSlots;
    float* st_rgb = st_sig + 32;
    [[maybe_unused]] uint32_t* feat_s = reinterpret_cast<uint32_t*>(reinterpret_cast<float*>(sm + kFwdHalfsPad) + kRenderWarps * kStageFloats) +
                       warp * kFeatWords;
    const MarchCfg c = make_cfg(p.f.bound_add, p.dt_gamma, p.max_steps, p.C, p.H);
    const uint32_t lt_mask = lanemask_lt();
    uint32_t evaluated = 0;

    for (;;) {
        uint32_t n = 0;
        if (lane == 0) n = atomicAdd(p.work_counter, 1u);
        n = __shfl_sync(NSIG_FULL_MASK, n, 0);
        if (n >= p.N) break;
        const RayConst r = load_ray(p.rays_o, p.rays_d, n);
        float near, far;
        near_far_one(r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, p.aabb, p.min_near, near, far);
        if (lane == 0) {
            if (p.nears) p.nears[n] = near;
            if (p.fars) p.fars[n] = far;
        }
        // the colour net's SH inputs depend on the ray only: build its first k-step once
        uint32_t sh_lo, sh_hi;
        {
            float o[16];
            sh4(r.dx, r.dy, r.dz, o);
            float lo0 = o[0], lo1 = o[1], hi0 = o[8], hi1 = o[9];
#pragma unroll
            for (int t = 1; t < 4; ++t)
                if (tig == t) { lo0 = o[2 * t]; lo1 = o[2 * t + 1]; hi0 = o[2 * t + 8]; hi1 = o[2 * t + 9]; }
            sh_lo = pack_h2(lo0, lo1);
            sh_hi = pack_h2(hi0, hi1);
        }
        const float t0 = perturbed_start(near, p.noises ? p.noises[n] : 0.0f, c);
        MarchState ms{t0, -INFINITY, t0};
        float T = 1.0f, t_acc = near;  // composite_rays starts its t from rays_t = nears (renderer_wtmk.py:345)
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_d = 0.f;
        uint32_t total = 0, nbuf = 0;
        for (;;) {
            // ---- march until 32 samples are staged or the ray ends ----
            while (nbuf < 32 && ms.t < far && total + nbuf < p.max_steps) {
                const Window w = march_window(ms, r, c, p.grid, far, lane, lt_mask);
                if ((w.emitted >> lane) & 1u) {
                    const uint32_t slot = nbuf + __popc(w.emitted & lt_mask);
                    st_xyz[slot * 3] = w.p.x; st_xyz[slot * 3 + 1] = w.p.y; st_xyz[slot * 3 + 2] = w.p.z;
                    st_dt[slot] = w.p.dt;
                    st_dr[slot] = w.delta_real;
                }
                nbuf += (uint32_t)__popc(w.emitted);
            }
            __syncwarp();
            const uint32_t nb = min(min(nbuf, 32u), p.max_steps - total);
            if (nb == 0) break;
            // ---- field on rows [0, nb) ----
            {
                uint32_t fa[MT][2][4];
#ifndef NSIG_GATHER_V2
#ifndef NSIG_RENDER_UNROLLED
                auto pos = [&](int mt, int h, float (&x)[3]) {
                    const uint32_t row = min((uint32_t)(mt * 16 + h * 8 + g), nb - 1);
#pragma unroll
                    for (int a = 0; a < 3; ++a) x[a] = __fmul_rn(__fadd_rn(st_xyz[row * 3 + a], p.f.bound_add), p.f.bound_mul);
                };
                encode_positions_rolled<MT, H2>(fa, p.f, pos, g, tig);
#else
                float xn[MT][2][3];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t row = min((uint32_t)(mt * 16 + h * 8 + g), nb - 1);
#pragma unroll
                        for (int a = 0; a < 3; ++a)
                            xn[mt][h][a] = __fmul_rn(__fadd_rn(st_xyz[row * 3 + a], p.f.bound_add), p.f.bound_mul);
                    }
                encode_positions<MT, H2>(fa, p.f, xn, g, tig);
#endif
#else
                {
                    const uint32_t row = min((uint32_t)lane, nb - 1);
                    const float px = __fmul_rn(__fadd_rn(st_xyz[row * 3], p.f.bound_add), p.f.bound_mul);
                    const float py = __fmul_rn(__fadd_rn(st_xyz[row * 3 + 1], p.f.bound_add), p.f.bound_mul);
                    const float pz = __fmul_rn(__fadd_rn(st_xyz[row * 3 + 2], p.f.bound_add), p.f.bound_mul);
                    gather_tile<H2>(fa, p.f, px, py, pz, feat_s, lane);
                }
#endif
                uint32_t h1[MT][4][4];
                {
                    float cc[MT][8][4];
                    layer<MT, 2, 8>(cc, fa, sm + oWs0, kS32, g, tig);
                    relu_to_a<MT, 8>(h1, cc);
                }
                float so[MT][2][4];
                layer<MT, 4, 2>(so, h1, sm + oWs1, kS64, g, tig);
                if (tig == 3) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            st_sig[mt * 16 + h * 8 + g] = __fmul_rn(p.f.density_scale, expf(so[mt][1][2 * h + 1]));
                }
                uint32_t ca[MT][2][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    ca[mt][0][0] = sh_lo; ca[mt][0][1] = sh_lo; ca[mt][0][2] = sh_hi; ca[mt][0][3] = sh_hi;
                }
                geo_to_a<MT>(ca, so, tig);
                uint32_t h2[MT][4][4];
                {
                    float cc[MT][8][4];
                    layer<MT, 2, 8>(cc, ca, sm + oWc0, kS32, g, tig);
                    relu_to_a<MT, 8>(h1, cc);
                    layer<MT, 4, 8>(cc, h1, sm + oWc1, kS64, g, tig);
                    relu_to_a<MT, 8>(h2, cc);
                }
                float co[MT][1][4];
                layer<MT, 4, 1>(co, h2, sm + oWc2, kS64, g, tig);
                if (tig < 2) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int row = mt * 16 + h * 8 + g;
                            const float s0 = 1.0f / (1.0f + expf(-co[mt][0][2 * h]));
                            if (tig == 0) {
                                st_rgb[row * 3] = s0;
                                st_rgb[row * 3 + 1] = 1.0f / (1.0f + expf(-co[mt][0][2 * h + 1]));
                            } else {
                                st_rgb[row * 3 + 2] = s0;
                            }
                        }
                }
            }
            __syncwarp();
            // ---- composite (raymarching.cu:850-890): lane k owns sample k ----
            bool terminated;
            {
                const bool valid = (uint32_t)lane < nb;
                const float sigma = valid ? st_sig[lane] : 0.f;
                const float dt = valid ? st_dt[lane] : 0.f;
                const float dr = valid ? st_dr[lane] : 0.f;
                const float alpha = 1.0f - __expf(-sigma * dt);
                const float one_m = valid ? (1.0f - alpha) : 1.0f;
                const float Tincl = T * warp_scan_mul(one_m, lane);
                float Tbefore = __shfl_up_sync(NSIG_FULL_MASK, Tincl, 1);
                if (lane == 0) Tbefore = T;
                const float tcum = t_acc + warp_scan_add(dr, lane);
                // the reference accumulates a sample, then stops if the transmittance BEFORE it was below T_thresh
                const uint32_t dead = __ballot_sync(NSIG_FULL_MASK, valid && (Tbefore < p.T_thresh));
                const int stop = dead ? (__ffs(dead) - 1) : 32;
                if (valid && lane <= stop) {
                    const float w = alpha * Tbefore;
                    acc_r += w * st_rgb[lane * 3]; acc_g += w * st_rgb[lane * 3 + 1]; acc_b += w * st_rgb[lane * 3 + 2];
                    acc_d += w * tcum;
                    acc_w += w;
                }
                terminated = dead != 0;
                T = __shfl_sync(NSIG_FULL_MASK, Tincl, 31);
                t_acc = __shfl_sync(NSIG_FULL_MASK, tcum, 31);
            }
            total += nb;
            evaluated += nb;
            if (terminated) break;
            // ---- keep the samples that did not fit into this batch ----
            const uint32_t left = nbuf - nb;
            float kx = 0.f, ky = 0.f, kz = 0.f, kdt = 0.f, kdr = 0.f;
            if ((uint32_t)lane < left) {
                const uint32_t src = nb + lane;
                kx = st_xyz[src * 3]; ky = st_xyz[src * 3 + 1]; kz = st_xyz[src * 3 + 2];
                kdt = st_dt[src]; kdr = st_dr[src];
            }
            __syncwarp();
            if ((uint32_t)lane < left) {
                st_xyz[lane * 3] = kx; st_xyz[lane * 3 + 1] = ky; st_xyz[lane * 3 + 2] = kz;
                st_dt[lane] = kdt; st_dr[lane] = kdr;
            }
            nbuf = left;
            __syncwarp();
        }
        acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
        acc_w = warp_sum(acc_w); acc_d = warp_sum(acc_d);
        if (lane == 0) {
            p.weights_sum[n] = acc_w;
            p.depth[n] = acc_d;
            p.image[(size_t)n * 3] = acc_r; p.image[(size_t)n * 3 + 1] = acc_g; p.image[(size_t)n * 3 + 2] = acc_b;
        }
    }
    if (p.sample_count && lane == 0 && evaluated) atomicAdd(p.sample_count, evaluated);
}

// ---------------------------------------------------------------------------------------
// backward (dgrad through both MLPs, scatter of the message-table gradient)
// ---------------------------------------------------------------------------------------
struct FieldBwdParams {
    const float* xyzs;
    const float* dirs;
    uint32_t M;
    float bound_add, bound_mul;
    const __half* feat;        // [M,32] saved by the forward
    const float* grad_sigmas;  // [M]
    const float* grad_rgbs;    // [M,3]
    const __half* sigma_w;
    const __half* color_w;
    float msg_grid_size;
    uint32_t mask;
    float* G;          // [T,2] gradient of the pre-summed message table (may be null)
    float* grad_feat;  // [M,32] fp32 full encoder-output gradient (may be null)
    const int32_t* M_dev;
    float density_scale;
    float* grad_sigma_w;  // [3072] fp32 weight gradients (accumulated; WGRAD only)
    float* grad_color_w;  // [7168]
};

// power-of-two scale that brings |v| into [0.5, 1): keeps the fp16 gradient chain in range
__device__ __forceinline__ float pow2_scale(float vmax) {
    if (!(vmax > 0.0f) || !isfinite(vmax)) return 1.0f;
    int e;
    frexpf(vmax, &e);
    e = max(-100, min(100, e));
    return scalbnf(1.0f, -e);
}

// ---- weight gradients (clean-model training, nerf/network_hash.py:154-161 trains both MLPs) --------------
// dW[out][in] = sum over rows of dOut[row][out] * In[row][in]: the contraction runs over the row index, which is
// the M index of BOTH operands as the forward/dgrad passes hold them (A-fragment layout).  movmatrix transposes
// the 8x8 blocks in registers, which turns an A fragment of dOut into the A fragment of dOut^T and an A fragment of
// In into the B fragment [k = row][n = in]; one m16n8k16 MMA per (16 outs x 8 ins) block then contracts the tile's
// 16 rows.  Products are accumulated per CTA in shared memory (fp32) and flushed once with global atomics.
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
    uint32_t r;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// rescale a packed fp16 pair by a power of two (exact until underflow)
__device__ __forceinline__ uint32_t scale_h2(uint32_t v, float ratio) {
    __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&v), __float2half2_rn(ratio));
    return *reinterpret_cast<uint32_t*>(&h);
}

// X: layer input  [16 rows x 16*KS_IN],  A-fragment layout (reg 2*hk + h: rows h*8+g, cols ks*16 + hk*8 + 2tig..)
// Y: output grad  [16 rows x 16*KS_OUT], same layout, rows pre-scaled to the tile's common power-of-two scale
// acc: shared fp32 [out][in] (row-major, leading dimension 16*KS_IN); ROT: output row r accumulates into param row
// (r + ROT) & 15 (the sigma net's second matrix is staged with its rows rotated by one).
template <int KS_IN, int KS_OUT, int ROT = 0>
__device__ __forceinline__ void wgrad_tile(float* __restrict__ acc, const uint32_t (&X)[KS_IN][4],
                                           const uint32_t (&Y)[KS_OUT][4], float inv_tile, int g, int tig) {
    constexpr int LD = 16 * KS_IN;
    uint32_t XT[KS_IN][4];
#pragma unroll
    for (int ks = 0; ks < KS_IN; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) XT[ks][r] = movmatrix_trans(X[ks][r]);
#pragma unroll
    for (int mo = 0; mo < KS_OUT; ++mo) {
        uint32_t a[4];
        a[0] = movmatrix_trans(Y[mo][0]);  // (outs 0-7,  rows 0-7)
        a[1] = movmatrix_trans(Y[mo][2]);  // (outs 8-15, rows 0-7)
        a[2] = movmatrix_trans(Y[mo][1]);  // (outs 0-7,  rows 8-15)
        a[3] = movmatrix_trans(Y[mo][3]);  // (outs 8-15, rows 8-15)
#pragma unroll
        for (int ni = 0; ni < 2 * KS_IN; ++ni) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(c, a, XT[ni >> 1][2 * (ni & 1)], XT[ni >> 1][2 * (ni & 1) + 1]);
            const int col = ni * 8 + 2 * tig;
            int r0 = mo * 16 + g, r1 = r0 + 8;
            if (ROT) { r0 = (r0 & ~15) | ((r0 + ROT) & 15); r1 = (r1 & ~15) | ((r1 + ROT) & 15); }
            atomicAdd(acc + r0 * LD + col, c[0] * inv_tile);
            atomicAdd(acc + r0 * LD + col + 1, c[1] * inv_tile);
            atomicAdd(acc + r1 * LD + col, c[2] * inv_tile);
            atomicAdd(acc + r1 * LD + col + 1, c[3] * inv_tile);
        }
    }
}

// fragments of a gradient tile brought from per-row scales to the tile's common scale
template <int KS>
__device__ __forceinline__ void to_tile_scale(uint32_t (&dst)[KS][4], const uint32_t (&src)[KS][4], float ratio0,
                                              float ratio1) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        dst[ks][0] = scale_h2(src[ks][0], ratio0);
        dst[ks][1] = scale_h2(src[ks][1], ratio1);
        dst[ks][2] = scale_h2(src[ks][2], ratio0);
        dst[ks][3] = scale_h2(src[ks][3], ratio1);
    }
}

// offsets of the five matrices inside the flat fp32 accumulators (= the FusedMLP `params` layout)
constexpr int aWs0 = 0, aWs1 = 2048, aWc0 = 3072, aWc1 = 3072 + 2048, aWc2 = 3072 + 2048 + 4096;
constexpr int kWgradFloats = NSIG_SIGMA_PARAMS + NSIG_COLOR_PARAMS;

template <bool FULL_GRAD, bool WGRAD>
__global__ void __launch_bounds__(kFieldThreads)
k_field_bwd(const FieldBwdParams p) {
    constexpr int MT = 1;
    extern __shared__ __align__(16) __half sm[];
    uint32_t M = p.M;
    if (p.M_dev) M = min(M, (uint32_t)max(*p.M_dev, 0));
    if (M == 0) return;
    stage_forward_weights(sm, p.sigma_w, p.color_w, true);
    stage_backward_weights(sm, p.sigma_w, p.color_w);
    float* wacc = reinterpret_cast<float*>(sm + kBwdHalfsPad);
    if (WGRAD)
        for (int i = threadIdx.x; i < kWgradFloats; i += blockDim.x) wacc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const uint32_t rows_per_cta = kFieldWarps * 16 * MT;
    const uint32_t n_tiles = div_up(M, rows_per_cta);
    const float* dirs_ = p.dirs;
    const uint32_t M_ = M;
    // Inputs of a tile (per thread): incoming gradients, saved features (A fragments), view directions of its rows and
    // the position of the row it scatters.  They are loaded one tile AHEAD: the loads of tile i+1 are issued before the
    // MLP work of tile i, so their L2/HBM latency (30 % of the warp time in ncu's source view of the non-prefetching
    // version, profiles/r01_experiments_v5.txt) is covered by ~300 tensor-core instructions instead of being waited for.
    struct TileIn {
        float gs[MT][2], gc[MT][2][2], dv[MT][2][3], sx[(MT * 2 + 3) / 4][3];
        uint32_t fa[MT][2][4];
    };
    auto load_tile = [&](uint32_t tile, TileIn& t) {
        const uint32_t row0 = tile * rows_per_cta + warp * 16 * MT;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t r = frag_row<MT>(row0, mt, h, g);
                t.gs[mt][h] = 0.f; t.gc[mt][h][0] = t.gc[mt][h][1] = 0.f;
                if (r < M) {
                    if (tig == 3) t.gs[mt][h] = __ldg(p.grad_sigmas + r);
                    if (tig == 0) { t.gc[mt][h][0] = __ldg(p.grad_rgbs + (size_t)r * 3); t.gc[mt][h][1] = __ldg(p.grad_rgbs + (size_t)r * 3 + 1); }
                    if (tig == 1) t.gc[mt][h][0] = __ldg(p.grad_rgbs + (size_t)r * 3 + 2);
                }
                const uint32_t rc = min(r, M - 1);
                const uint32_t* src = reinterpret_cast<const uint32_t*>(p.feat + (size_t)rc * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) t.fa[mt][j >> 1][2 * (j & 1) + h] = __ldg(src + tig + 4 * j);
#pragma unroll
                for (int a = 0; a < 3; ++a) t.dv[mt][h][a] = __ldg(dirs_ + (size_t)rc * 3 + a);
            }
        // the row whose 8 corners this thread scatters: quad thread `tig` takes row (mt,h) = 4q + tig
#pragma unroll
        for (int q = 0; q < (MT * 2 + 3) / 4; ++q) {
            const int sel = q * 4 + tig;
            const uint32_t r = row0 + (sel >> 1) * 16 + (sel & 1) * 8 + g;
#pragma unroll
            for (int a = 0; a < 3; ++a) t.sx[q][a] = (sel < MT * 2 && r < M) ? __ldg(p.xyzs + (size_t)r * 3 + a) : 0.f;
        }
    };
    TileIn nxt;
    if (blockIdx.x < n_tiles) load_tile(blockIdx.x, nxt);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileIn cur = nxt;
        if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x, nxt);
        const uint32_t row0 = tile * rows_per_cta + warp * 16 * MT;
        if (row0 >= M) continue;
        // incoming gradients of this thread's rows; skip the tile when the whole warp has none
        float gs_in[MT][2], gc_in[MT][2][2];
        bool any = false;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                gs_in[mt][h] = cur.gs[mt][h]; gc_in[mt][h][0] = cur.gc[mt][h][0]; gc_in[mt][h][1] = cur.gc[mt][h][1];
                any |= (gs_in[mt][h] != 0.f) | (gc_in[mt][h][0] != 0.f) | (gc_in[mt][h][1] != 0.f);
            }
        if (!__any_sync(NSIG_FULL_MASK, any)) {
            if (FULL_GRAD && p.grad_feat) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t r = frag_row<MT>(row0, mt, h, g);
                        if (r < M)
#pragma unroll
                            for (int nt = 0; nt < 4; ++nt)
                                *reinterpret_cast<float2*>(p.grad_feat + (size_t)r * 32 + nt * 8 + 2 * tig) = make_float2(0.f, 0.f);
                    }
            }
            continue;
        }
        // ---- recompute the forward activations from the saved encoder output ----
        uint32_t fa[MT][2][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int e = 0; e < 4; ++e) fa[mt][ks][e] = cur.fa[mt][ks][e];
        uint32_t h1s[MT][4][4], h1c[MT][4][4], h2c[MT][4][4], ca[MT][2][4];
        float so[MT][2][4], co[MT][1][4];
        {
            float c[MT][8][4];
            layer<MT, 2, 8>(c, fa, sm + oWs0, kS32, g, tig);
            relu_to_a<MT, 8>(h1s, c);
            layer<MT, 4, 2>(so, h1s, sm + oWs1, kS64, g, tig);
            sh_rows_vals<MT>(ca, cur.dv, tig);
            geo_to_a<MT>(ca, so, tig);
            layer<MT, 2, 8>(c, ca, sm + oWc0, kS32, g, tig);
            relu_to_a<MT, 8>(h1c, c);
            layer<MT, 4, 8>(c, h1c, sm + oWc1, kS64, g, tig);
            relu_to_a<MT, 8>(h2c, c);
            layer<MT, 4, 1>(co, h2c, sm + oWc2, kS64, g, tig);
        }
        // ---- output-activation gradients, normalised per row by a power of two ----
        float d_rgb[MT][2][2], d_logit[MT][2], inv_scale[MT][2], row_sc[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // sigmoid'(z) = s(1-s)
                const float s0 = 1.0f / (1.0f + expf(-co[mt][0][2 * h])), s1 = 1.0f / (1.0f + expf(-co[mt][0][2 * h + 1]));
                d_rgb[mt][h][0] = (tig < 2) ? gc_in[mt][h][0] * s0 * (1.0f - s0) : 0.f;
                d_rgb[mt][h][1] = (tig == 0) ? gc_in[mt][h][1] * s1 * (1.0f - s1) : 0.f;
                // trunc_exp backward: g * exp(clamp(x, -15, 15))  (activation.py:14-16)
                const float logit = so[mt][1][2 * h + 1];
                d_logit[mt][h] = (tig == 3) ? gs_in[mt][h] * p.density_scale * expf(fminf(fmaxf(logit, -15.0f), 15.0f)) : 0.f;
                float vmax = fmaxf(fmaxf(fabsf(d_rgb[mt][h][0]), fabsf(d_rgb[mt][h][1])), fabsf(d_logit[mt][h]));
                vmax = fmaxf(vmax, __shfl_xor_sync(NSIG_FULL_MASK, vmax, 1));
                vmax = fmaxf(vmax, __shfl_xor_sync(NSIG_FULL_MASK, vmax, 2));
                const float sc = pow2_scale(vmax);
                row_sc[mt][h] = (vmax > 0.0f && isfinite(vmax)) ? sc : INFINITY;  // rows without gradient do not set the tile scale
                inv_scale[mt][h] = 1.0f / sc;
                d_rgb[mt][h][0] *= sc; d_rgb[mt][h][1] *= sc; d_logit[mt][h] *= sc;
            }
        // common power-of-two scale of the tile's rows (weight gradients contract over rows)
        float wr0 = 0.f, wr1 = 0.f, inv_tile = 0.f;
        if (WGRAD) {
            float s_tile = fminf(row_sc[0][0], row_sc[0][1]);
            s_tile = fminf(s_tile, __shfl_xor_sync(NSIG_FULL_MASK, s_tile, 4));
            s_tile = fminf(s_tile, __shfl_xor_sync(NSIG_FULL_MASK, s_tile, 8));
            s_tile = fminf(s_tile, __shfl_xor_sync(NSIG_FULL_MASK, s_tile, 16));
            if (isfinite(s_tile)) {
                inv_tile = 1.0f / s_tile;
                wr0 = isfinite(row_sc[0][0]) ? s_tile / row_sc[0][0] : 0.f;
                wr1 = isfinite(row_sc[0][1]) ? s_tile / row_sc[0][1] : 0.f;
            }
        }
        // ---- colour net dgrad ----
        uint32_t da[MT][1][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            da[mt][0][0] = pack_h2(d_rgb[mt][0][0], d_rgb[mt][0][1]);
            da[mt][0][1] = pack_h2(d_rgb[mt][1][0], d_rgb[mt][1][1]);
            da[mt][0][2] = 0u; da[mt][0][3] = 0u;
        }
        uint32_t dh[MT][4][4];
        float dgeo[MT][2][4];
        {
            float c[MT][8][4];
            if (WGRAD) {   // dWc2 = d out^T x h2
                uint32_t y[1][4];
                to_tile_scale<1>(y, da[0], wr0, wr1);
                wgrad_tile<4, 1>(wacc + aWc2, h2c[0], y, inv_tile, g, tig);
            }
            layer<MT, 1, 8>(c, da, sm + oWc2T, kS16, g, tig);        // d h2 = d out x W2
            grad_to_a<MT, 8>(dh, c, h2c);
            if (WGRAD) {   // dWc1 = d z2^T x h1
                uint32_t y[4][4];
                to_tile_scale<4>(y, dh[0], wr0, wr1);
                wgrad_tile<4, 4>(wacc + aWc1, h1c[0], y, inv_tile, g, tig);
            }
            layer<MT, 4, 8>(c, dh, sm + oWc1T, kS64, g, tig);        // d h1 = d h2 x W1
            grad_to_a<MT, 8>(dh, c, h1c);
            if (WGRAD) {   // dWc0 = d z1^T x [SH, geo]
                uint32_t y[4][4];
                to_tile_scale<4>(y, dh[0], wr0, wr1);
                wgrad_tile<2, 4>(wacc + aWc0, ca[0], y, inv_tile, g, tig);
            }
            layer<MT, 4, 2>(dgeo, dh, sm + oWc0T, kS64, g, tig);     // d geo = (d h1 x W0)[:, 16:32]
        }
        // ---- sigma net dgrad: d out' = [d geo0..14, d logit] ----
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            da[mt][0][0] = pack_h2(dgeo[mt][0][0], dgeo[mt][0][1]);
            da[mt][0][1] = pack_h2(dgeo[mt][0][2], dgeo[mt][0][3]);
            da[mt][0][2] = pack_h2(dgeo[mt][1][0], (tig == 3) ? d_logit[mt][0] : dgeo[mt][1][1]);
            da[mt][0][3] = pack_h2(dgeo[mt][1][2], (tig == 3) ? d_logit[mt][1] : dgeo[mt][1][3]);
        }
        {
            float c[MT][8][4];
            if (WGRAD) {   // dWs1 = d out'^T x h1s (staged rows are rotated by one: out' row r is param row (r+1)%16)
                uint32_t y[1][4];
                to_tile_scale<1>(y, da[0], wr0, wr1);
                wgrad_tile<4, 1, 1>(wacc + aWs1, h1s[0], y, inv_tile, g, tig);
            }
            layer<MT, 1, 8>(c, da, sm + oWs1T, kS16, g, tig);        // d h1s = d out' x W1'
            grad_to_a<MT, 8>(dh, c, h1s);
            if (WGRAD) {   // dWs0 = d z1s^T x feat
                uint32_t y[4][4];
                to_tile_scale<4>(y, dh[0], wr0, wr1);
                wgrad_tile<2, 4>(wacc + aWs0, fa[0], y, inv_tile, g, tig);
            }
        }
        float gmsg[MT][2][2];  // d feature 30,31 (valid on tig == 3)
        if (FULL_GRAD) {
            float c[MT][4][4];
            layer<MT, 4, 4>(c, dh, sm + oWs0T, kS64, g, tig);        // d feat = d h1s x W0
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t r = frag_row<MT>(row0, mt, h, g);
                    gmsg[mt][h][0] = c[mt][3][2 * h] * inv_scale[mt][h];
                    gmsg[mt][h][1] = c[mt][3][2 * h + 1] * inv_scale[mt][h];
                    if (r < M && p.grad_feat)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
                            *reinterpret_cast<float2*>(p.grad_feat + (size_t)r * 32 + nt * 8 + 2 * tig) =
                                make_float2(c[mt][nt][2 * h] * inv_scale[mt][h], c[mt][nt][2 * h + 1] * inv_scale[mt][h]);
                }
        } else {
            float c[MT][1][4];
            layer<MT, 4, 1, 3>(c, dh, sm + oWs0T, kS64, g, tig);     // only channels 24..31
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    gmsg[mt][h][0] = c[mt][0][2 * h] * inv_scale[mt][h];
                    gmsg[mt][h][1] = c[mt][0][2 * h + 1] * inv_scale[mt][h];
                }
        }
        // ---- scatter into G: quad thread `tig` handles all 8 corners of row (mt,h) = tig ----
        if (p.G) {
#pragma unroll
            for (int q = 0; q < (MT * 2 + 3) / 4; ++q) {
                float gx = 0.f, gy = 0.f;
                uint32_t myrow = 0xffffffffu;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int dst = mt * 2 + h - q * 4;
                        if (dst >= 0 && dst < 4) {
                            const float vx = __shfl_sync(NSIG_FULL_MASK, gmsg[mt][h][0], (g << 2) | 3);
                            const float vy = __shfl_sync(NSIG_FULL_MASK, gmsg[mt][h][1], (g << 2) | 3);
                            if (tig == dst) { gx = vx; gy = vy; myrow = frag_row<MT>(row0, mt, h, g); }
                        }
                    }
                if (myrow < M && (gx != 0.f || gy != 0.f)) {
                    float xn[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                        xn[a] = __fmul_rn(__fadd_rn(cur.sx[q][a], p.bound_add), p.bound_mul);   // prefetched position of myrow
                    const Voxel v = locate(xn[0], xn[1], xn[2], p.msg_grid_size);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        red_add_v2(p.G + (size_t)corner_slot(v, k, p.mask) * 2, corner_grad(v, k, gx), corner_grad(v, k, gy));
                }
            }
        }
    }
    if (WGRAD) {  // one global atomic per weight and CTA
        __syncthreads();
        for (int i = threadIdx.x; i < kWgradFloats; i += blockDim.x) {
            const float v = wacc[i];
            if (v != 0.0f) atomicAdd((i < NSIG_SIGMA_PARAMS ? p.grad_sigma_w + i : p.grad_color_w + (i - NSIG_SIGMA_PARAMS)), v);
        }
    }
}


// ---------------------------------------------------------------------------------------
// Watermark-mode backward from SAVED ReLU MASKS (the hot path: only dL/dS is needed).
// k_field_bwd above recomputes the five forward GEMMs of a tile from the saved encoder output just to learn (a) which hidden
// units were active and (b) sigmoid'(z), exp(logit).  (b) follows from the forward's own outputs (rgb, sigma), and (a) is 3 x 64
// bits per sample, which k_field_fwd writes next to them (relu_mask_bits: 32 B per sample instead of the 64 B of features).
// What is left is the five dgrad GEMMs: 60 instead of 136 m16n8k16 MMAs per 16 rows, no SH, no feature reload.
// Bit-identical masks by construction (same accumulators, same rounding rule), so dL/dS differs from k_field_bwd's only by
// exp(logit) being read back as sigma / density_scale (1 ulp).
// ---------------------------------------------------------------------------------------
struct FieldBwdMaskParams {
    const float* xyzs;
    uint32_t M;
    float bound_add, bound_mul;
    const uint2* masks;        // [M][4] (row, quad thread): x = m1s | m1c << 8, y = m2c (bit layout: act_mask_bits)
    const float* sigmas;       // [M]   forward outputs
    const float* rgbs;         // [M,3]
    const float* grad_sigmas;  // [M]
    const float* grad_rgbs;    // [M,3]
    const __half* sigma_w;
    const __half* color_w;
    float msg_grid_size;
    uint32_t mask;
    float* G;
    const int32_t* M_dev;
    float density_scale;
};

__global__ void __launch_bounds__(kFieldThreads)
k_field_bwd_masks(const FieldBwdMaskParams p) {
    constexpr int MT = 1;
    extern __shared__ __align__(16) __half sm[];
    uint32_t M = p.M;
    if (p.M_dev) M = min(M, (uint32_t)max(*p.M_dev, 0));
    if (M == 0) return;
    stage_backward_weights(sm, p.sigma_w, p.color_w);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const uint32_t rows_per_cta = kFieldWarps * 16 * MT;
    const uint32_t n_tiles = div_up(M, rows_per_cta);
    const float inv_ds = 1.0f / p.density_scale;
    struct TileIn {
        float gs[2], gc[2][2], act[2][2], sx[3];   // act: tig 3 -> sigma; tig 0 -> (r, g); tig 1 -> b
        uint2 mk[2];
    };
    auto load_tile = [&](uint32_t tile, TileIn& t) {
        const uint32_t row0 = tile * rows_per_cta + warp * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t r = row0 + h * 8 + g;
            t.gs[h] = 0.f; t.gc[h][0] = t.gc[h][1] = 0.f; t.act[h][0] = t.act[h][1] = 0.f;
            t.mk[h] = make_uint2(0u, 0u);
            if (r < M) {
                if (tig == 3) { t.gs[h] = __ldg(p.grad_sigmas + r); t.act[h][0] = __ldg(p.sigmas + r); }
                if (tig == 0) {
                    t.gc[h][0] = __ldg(p.grad_rgbs + (size_t)r * 3); t.gc[h][1] = __ldg(p.grad_rgbs + (size_t)r * 3 + 1);
                    t.act[h][0] = __ldg(p.rgbs + (size_t)r * 3); t.act[h][1] = __ldg(p.rgbs + (size_t)r * 3 + 1);
                }
                if (tig == 1) { t.gc[h][0] = __ldg(p.grad_rgbs + (size_t)r * 3 + 2); t.act[h][0] = __ldg(p.rgbs + (size_t)r * 3 + 2); }
                t.mk[h] = __ldg(p.masks + (size_t)r * 4 + tig);
            }
        }
        // the row whose 8 corners this thread scatters: quad thread `tig` takes row (h = tig & 1) of the tile's row pair
        const uint32_t rs = row0 + (tig & 1) * 8 + g;
#pragma unroll
        for (int a = 0; a < 3; ++a) t.sx[a] = (tig < 2 && rs < M) ? __ldg(p.xyzs + (size_t)rs * 3 + a) : 0.f;
    };
    TileIn nxt;
    if (blockIdx.x < n_tiles) load_tile(blockIdx.x, nxt);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileIn cur = nxt;
        if (tile + gridDim.x < n_tiles) load_tile(tile + gridDim.x, nxt);
        const uint32_t row0 = tile * rows_per_cta + warp * 16;
        if (row0 >= M) continue;
        bool any = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) any |= (cur.gs[h] != 0.f) | (cur.gc[h][0] != 0.f) | (cur.gc[h][1] != 0.f);
        if (!__any_sync(NSIG_FULL_MASK, any)) continue;
        // ---- output-activation gradients, normalised per row by a power of two ----
        float d_rgb[2][2], d_logit[2], inv_scale[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float s0 = cur.act[h][0], s1 = cur.act[h][1];                // sigmoid'(z) = s (1 - s)
            d_rgb[h][0] = (tig < 2) ? cur.gc[h][0] * s0 * (1.0f - s0) : 0.f;
            d_rgb[h][1] = (tig == 0) ? cur.gc[h][1] * s1 * (1.0f - s1) : 0.f;
            // trunc_exp backward: g * exp(clamp(x, -15, 15)) (activation.py:14-16), with exp(x) = sigma / density_scale
            const float e = fminf(fmaxf(s0 * inv_ds, 3.0590232050182579e-7f), 3269017.3724721107f);
            d_logit[h] = (tig == 3) ? cur.gs[h] * p.density_scale * e : 0.f;
            float vmax = fmaxf(fmaxf(fabsf(d_rgb[h][0]), fabsf(d_rgb[h][1])), fabsf(d_logit[h]));
            vmax = fmaxf(vmax, __shfl_xor_sync(NSIG_FULL_MASK, vmax, 1));
            vmax = fmaxf(vmax, __shfl_xor_sync(NSIG_FULL_MASK, vmax, 2));
            const float sc = pow2_scale(vmax);
            inv_scale[h] = 1.0f / sc;
            d_rgb[h][0] *= sc; d_rgb[h][1] *= sc; d_logit[h] *= sc;
        }
        uint32_t m1[MT][2], m2[MT][2];   // m1 = m1s | m1c << 8, m2 = m2c (bit layout: act_mask_bits)
#pragma unroll
        for (int h = 0; h < 2; ++h) { m1[0][h] = cur.mk[h].x; m2[0][h] = cur.mk[h].y; }
        // ---- colour net dgrad ----
        uint32_t da[MT][1][4];
        da[0][0][0] = pack_h2(d_rgb[0][0], d_rgb[0][1]);
        da[0][0][1] = pack_h2(d_rgb[1][0], d_rgb[1][1]);
        da[0][0][2] = 0u; da[0][0][3] = 0u;
        uint32_t dh[MT][4][4];
        float dgeo[MT][2][4];
        {
            float c[MT][8][4];
            layer<MT, 1, 8>(c, da, sm + oWc2T, kS16, g, tig);        // d h2 = d out x W2
            grad_to_a_bits<MT, 8>(dh, c, m2, 0);
            layer<MT, 4, 8>(c, dh, sm + oWc1T, kS64, g, tig);        // d h1 = d h2 x W1
            grad_to_a_bits<MT, 8>(dh, c, m1, 8);
            layer<MT, 4, 2>(dgeo, dh, sm + oWc0T, kS64, g, tig);     // d geo = (d h1 x W0)[:, 16:32]
        }
        // ---- sigma net dgrad: d out' = [d geo0..14, d logit] ----
        da[0][0][0] = pack_h2(dgeo[0][0][0], dgeo[0][0][1]);
        da[0][0][1] = pack_h2(dgeo[0][0][2], dgeo[0][0][3]);
        da[0][0][2] = pack_h2(dgeo[0][1][0], (tig == 3) ? d_logit[0] : dgeo[0][1][1]);
        da[0][0][3] = pack_h2(dgeo[0][1][2], (tig == 3) ? d_logit[1] : dgeo[0][1][3]);
        float gmsg[2][2];  // d feature 30,31 (valid on tig == 3)
        {
            float c[MT][8][4];
            layer<MT, 1, 8>(c, da, sm + oWs1T, kS16, g, tig);        // d h1s = d out' x W1'
            grad_to_a_bits<MT, 8>(dh, c, m1, 0);
            float c2[MT][1][4];
            layer<MT, 4, 1, 3>(c2, dh, sm + oWs0T, kS64, g, tig);    // d feat, channels 24..31 only
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                gmsg[h][0] = c2[0][0][2 * h] * inv_scale[h];
                gmsg[h][1] = c2[0][0][2 * h + 1] * inv_scale[h];
            }
        }
        // ---- scatter into G: quad threads 0 and 1 take rows g and g + 8 ----
        float gx = 0.f, gy = 0.f;
        uint32_t myrow = 0xffffffffu;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float vx = __shfl_sync(NSIG_FULL_MASK, gmsg[h][0], (g << 2) | 3);
            const float vy = __shfl_sync(NSIG_FULL_MASK, gmsg[h][1], (g << 2) | 3);
            if (tig == h) { gx = vx; gy = vy; myrow = row0 + h * 8 + g; }
        }
        if (myrow < M && (gx != 0.f || gy != 0.f)) {
            float xn[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) xn[a] = __fmul_rn(__fadd_rn(cur.sx[a], p.bound_add), p.bound_mul);
            const Voxel v = locate(xn[0], xn[1], xn[2], p.msg_grid_size);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                red_add_v2(p.G + (size_t)corner_slot(v, k, p.mask) * 2, corner_grad(v, k, gx), corner_grad(v, k, gy));
        }
    }
}

}  // namespace nsig

using namespace nsig;

extern "C" {

int nsig_field_forward(const float* xyzs, const float* dirs, uint32_t M, float bound, const float* const* tables,
                       const float* resolutions, uint32_t log2_T, const float* S, float msg_resolution,
                       const void* sigma_w, const void* color_w, float density_scale, const int32_t* M_dev,
                       float* sigmas, float* rgbs, void* feat_out, void* masks_out, const void* const* tables_h2,
                       const float* h2_inv_scale, nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!dirs || !color_w || !sigmas || !rgbs) return NSIG_EINVAL;
    FieldParams p;
    const int rc = fill_field_params(p, xyzs, dirs, M, bound, tables, resolutions, log2_T, S, msg_resolution,
                                     sigma_w, color_w, M_dev, density_scale, tables_h2, h2_inv_scale);
    if (rc) return rc;
    const size_t smem = kFieldFwdSmem;
    cudaStream_t st = (cudaStream_t)stream;
    __half* feat = reinterpret_cast<__half*>(feat_out);
    uint2* masks = reinterpret_cast<uint2*>(masks_out);
    if (((uintptr_t)masks_out) & 7) return NSIG_EINVAL;
    if (tables_h2)
        k_field_fwd<true, true><<<field_grid(k_field_fwd<true, true>, smem, M, kFieldWarps * 32), kFieldThreads, smem, st>>>(
            p, sigmas, rgbs, feat, nullptr, masks);
    else
        k_field_fwd<true, false><<<field_grid(k_field_fwd<true, false>, smem, M, kFieldWarps * 32), kFieldThreads, smem, st>>>(
            p, sigmas, rgbs, feat, nullptr, masks);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_field_density(const float* xyzs, uint32_t M, float bound, const float* const* tables,
                       const float* resolutions, uint32_t log2_T, const float* S, float msg_resolution,
                       const void* sigma_w, float density_scale, float* sigmas, void* geo_feat,
                       const void* const* tables_h2, const float* h2_inv_scale, nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!sigmas) return NSIG_EINVAL;
    FieldParams p;
    const int rc = fill_field_params(p, xyzs, nullptr, M, bound, tables, resolutions, log2_T, S, msg_resolution,
                                     sigma_w, nullptr, nullptr, density_scale, tables_h2, h2_inv_scale);
    if (rc) return rc;
    const size_t smem = kFieldFwdSmem;
    cudaStream_t st = (cudaStream_t)stream;
    __half* geo = reinterpret_cast<__half*>(geo_feat);
    if (tables_h2)
        k_field_fwd<false, true><<<field_grid(k_field_fwd<false, true>, smem, M, kFieldWarps * 32), kFieldThreads, smem, st>>>(
            p, sigmas, nullptr, nullptr, geo, nullptr);
    else
        k_field_fwd<false, false><<<field_grid(k_field_fwd<false, false>, smem, M, kFieldWarps * 32), kFieldThreads, smem, st>>>(
            p, sigmas, nullptr, nullptr, geo, nullptr);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_color_forward(const float* dirs, const void* geo_feat, uint32_t M, const void* color_w, float* rgbs,
                       nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!dirs || !geo_feat || !color_w || !rgbs) return NSIG_EINVAL;
    const size_t smem = kFwdHalfs * sizeof(__half);
    k_color_fwd<<<field_grid(k_color_fwd, smem, M, kFieldWarps * 32), kFieldThreads, smem, (cudaStream_t)stream>>>(
        dirs, reinterpret_cast<const __half*>(geo_feat), M, reinterpret_cast<const __half*>(color_w), rgbs);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_render_rays(const float* rays_o, const float* rays_d, uint32_t N, const float* aabb, float min_near,
                     float bound, const uint8_t* grid, uint32_t C, uint32_t H, float dt_gamma, uint32_t max_steps,
                     float T_thresh, const float* noises, const float* const* tables, const float* resolutions,
                     uint32_t log2_T, const float* S, float msg_resolution, const void* sigma_w, const void* color_w,
                     float density_scale, uint32_t* work_counter, float* weights_sum, float* depth, float* image,
                     float* nears, float* fars, uint32_t* sample_count, const void* const* tables_h2,
                     const float* h2_inv_scale, nsig_stream_t stream) {
    if (N == 0) return 0;
    if (!rays_o || !rays_d || !aabb || !grid || !color_w || !work_counter || !weights_sum || !depth || !image)
        return NSIG_EINVAL;
    if (C == 0 || C > 31 || H == 0 || H > 1024 || max_steps == 0) return NSIG_EINVAL;
    RenderParams p;
    const int rc = fill_field_params(p.f, rays_o /*unused*/, rays_d, 0, bound, tables, resolutions, log2_T, S,
                                     msg_resolution, sigma_w, color_w, nullptr, density_scale, tables_h2, h2_inv_scale);
    if (rc) return rc;
    p.rays_o = rays_o; p.rays_d = rays_d; p.N = N; p.grid = grid; p.dt_gamma = dt_gamma;
    p.max_steps = max_steps; p.C = C; p.H = H; p.aabb = aabb; p.min_near = min_near; p.T_thresh = T_thresh;
    p.noises = noises; p.work_counter = work_counter; p.weights_sum = weights_sum; p.depth = depth; p.image = image;
    p.nears = nears; p.fars = fars; p.sample_count = sample_count;
    const size_t smem = kFwdHalfsPad * sizeof(__half) + (size_t)kRenderWarps * (kStageFloats * sizeof(float) + kFeatWordsAlloc * 4);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t want = div_up(N, kRenderWarps);
    int per_sm = 3;
    if (tables_h2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_render_rays<true>, kRenderWarps * 32, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_render_rays<false>, kRenderWarps * 32, smem);
    const uint32_t cap = (uint32_t)sms * (uint32_t)(per_sm > 0 ? per_sm : 1);  // persistent: one resident wave
    const uint32_t blocks = want < cap ? want : cap;
    if (tables_h2) k_render_rays<true><<<blocks, kRenderWarps * 32, smem, (cudaStream_t)stream>>>(p);
    else k_render_rays<false><<<blocks, kRenderWarps * 32, smem, (cudaStream_t)stream>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_field_backward(const float* xyzs, const float* dirs, uint32_t M, float bound, const void* feat,
                        const float* grad_sigmas, const float* grad_rgbs, const void* sigma_w, const void* color_w,
                        float density_scale, const int32_t* M_dev, float msg_resolution, uint32_t log2_T,
                        float* G, float* grad_feat, float* grad_sigma_w, float* grad_color_w,
                        nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!xyzs || !dirs || !feat || !grad_sigmas || !grad_rgbs || !sigma_w || !color_w) return NSIG_EINVAL;
    if ((grad_sigma_w == nullptr) != (grad_color_w == nullptr)) return NSIG_EINVAL;
    if ((((uintptr_t)sigma_w) | ((uintptr_t)color_w)) & 15) return NSIG_EINVAL;   // 16-byte weight staging
    if (log2_T < 1 || log2_T > 30 || !(bound > 0.0f)) return NSIG_EINVAL;
    if (G && !(msg_resolution > 0.0f)) return NSIG_EINVAL;
    FieldBwdParams p;
    p.xyzs = xyzs; p.dirs = dirs; p.M = M;
    p.bound_add = bound; p.bound_mul = 1.0f / (2.0f * bound);
    p.feat = reinterpret_cast<const __half*>(feat);
    p.grad_sigmas = grad_sigmas; p.grad_rgbs = grad_rgbs;
    p.sigma_w = reinterpret_cast<const __half*>(sigma_w);
    p.color_w = reinterpret_cast<const __half*>(color_w);
    p.msg_grid_size = (msg_resolution > 0.0f) ? 1.0f / msg_resolution : 0.0f;
    p.mask = (1u << log2_T) - 1u;
    p.G = G; p.grad_feat = grad_feat;
    p.M_dev = M_dev; p.density_scale = density_scale;
    p.grad_sigma_w = grad_sigma_w; p.grad_color_w = grad_color_w;
    const bool wgrad = grad_sigma_w != nullptr;
    const size_t smem = kBwdHalfsPad * sizeof(__half) + (wgrad ? kWgradFloats * sizeof(float) : 0);
    static bool attr_set = false;
    if (!attr_set) {
        const int big = (int)(kBwdHalfsPad * sizeof(__half) + kWgradFloats * sizeof(float));
        cudaFuncSetAttribute(k_field_bwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_field_bwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute(k_field_bwd<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (wgrad) {  // weight gradients imply the full feature gradient path (clean-model training)
        k_field_bwd<true, true><<<field_grid(k_field_bwd<true, true>, smem, M, kFieldWarps * 16), kFieldThreads, smem, st>>>(p);
    } else if (grad_feat) {
        k_field_bwd<true, false><<<field_grid(k_field_bwd<true, false>, smem, M, kFieldWarps * 16), kFieldThreads, smem, st>>>(p);
    } else {
        k_field_bwd<false, false><<<field_grid(k_field_bwd<false, false>, smem, M, kFieldWarps * 16), kFieldThreads, smem, st>>>(p);
    }
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_field_backward_masks(const float* xyzs, uint32_t M, float bound, const void* masks, const float* sigmas,
                              const float* rgbs, const float* grad_sigmas, const float* grad_rgbs, const void* sigma_w,
                              const void* color_w, float density_scale, const int32_t* M_dev, float msg_resolution,
                              uint32_t log2_T, float* G, nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!xyzs || !masks || !sigmas || !rgbs || !grad_sigmas || !grad_rgbs || !sigma_w || !color_w || !G) return NSIG_EINVAL;
    if ((((uintptr_t)sigma_w) | ((uintptr_t)color_w)) & 15) return NSIG_EINVAL;   // 16-byte weight staging
    if (((uintptr_t)masks) & 7) return NSIG_EINVAL;
    if (log2_T < 1 || log2_T > 30 || !(bound > 0.0f) || !(msg_resolution > 0.0f) || !(density_scale > 0.0f)) return NSIG_EINVAL;
    FieldBwdMaskParams p;
    p.xyzs = xyzs; p.M = M; p.bound_add = bound; p.bound_mul = 1.0f / (2.0f * bound);
    p.masks = reinterpret_cast<const uint2*>(masks);
    p.sigmas = sigmas; p.rgbs = rgbs; p.grad_sigmas = grad_sigmas; p.grad_rgbs = grad_rgbs;
    p.sigma_w = reinterpret_cast<const __half*>(sigma_w);
    p.color_w = reinterpret_cast<const __half*>(color_w);
    p.msg_grid_size = 1.0f / msg_resolution;
    p.mask = (1u << log2_T) - 1u;
    p.G = G; p.M_dev = M_dev; p.density_scale = density_scale;
    const size_t smem = kBwdHalfsPad * sizeof(__half);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_field_bwd_masks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    k_field_bwd_masks<<<field_grid(k_field_bwd_masks, smem, M, kFieldWarps * 16), kFieldThreads, smem, (cudaStream_t)stream>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
