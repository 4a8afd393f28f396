// grid.cu — the occupancy-grid side of the hot path as fused sm_100a kernels:
//
//   * k_grid_sweep      NeRFRenderer.update_extra_state's density sweep (nerf/renderer_wtmk.py:456-514):
//                       cell -> centre + jitter -> normalise -> 16-level hash encode (+ message feature) ->
//                       sigma MLP on tensor cores -> trunc_exp -> EMA-max into density_grid, and the running sum
//                       for mean_density.  The reference builds a [128^3,3] meshgrid, a Morton index tensor, two
//                       random tensors and runs ~450 torch kernels per cascade; here it is one launch for all
//                       cascades and no intermediate touches HBM.
//   * k_grid_finalize   the EMA/max + mean of the partial update (renderer_wtmk.py:521-524), where cells can be
//                       drawn more than once and the sweep therefore records sigma in a temporary grid.
//   * k_grid_pack       packbits with the threshold min(mean_density, density_thresh) taken from the device-side
//                       sum (renderer_wtmk.py:524-530) - no .item() between the sweep and the bitfield.
//   * k_cells_*         the partial update's cell selection (renderer_wtmk.py:489-501): N uniform cells + N cells
//                       drawn uniformly from the currently occupied ones, without torch.nonzero's host sync.
//   * k_mark_untrained  NeRFRenderer.mark_untrained_grid (renderer_wtmk.py:380-442): frustum coverage of every
//                       cell over all training cameras in one launch.
//   * k_get_rays        get_rays (nerf/utils_wtmk_disen.py:59-143): pose + intrinsics + pixel ids -> rays.
//
// Arithmetic that feeds integer decisions or is compared against the reference (cell centres, jitter scaling,
// EMA, thresholds) is written one rounded fp32 operation per torch operation.
#include "field_common.cuh"

namespace nsig {

// ---- Philox4x32-10 (counter-based; one call per cell, nothing to store or advance) -------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x, hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }  // [0,1), 24 bits

struct SweepParams {
    FieldParams f;        // tables, S, sigma weights, bound, density_scale (xyzs/dirs/M unused)
    float* grid;          // [C, H^3]
    float* tmp;           // optional [C, H^3]: record sigma here (atomic max) instead of updating `grid`
    const int32_t* cells; // optional [C, n] Morton indices; null = every cell, in Morton order
    const float* noise;   // optional [C, n, 3] in [0,1)
    uint32_t n, C, H;
    double bound;
    float decay;
    uint32_t seed_lo, seed_hi;
    double* sum;          // optional: += sum of max(grid, 0) over the visited cells (full update only)
};

// Cell centre + jitter in the reference's operation order (renderer_wtmk.py:470-479):
//   xyzs = 2 * coords.float() / (H - 1) - 1          (torch divides by a scalar as a multiply with fl(1/s))
//   cas_xyzs = xyzs * (bound_c - half)               (python double, rounded to fp32 by the multiply)
//   cas_xyzs += (rand * 2 - 1) * half
__device__ __forceinline__ float cell_axis(uint32_t c, float u, float r_hm1, float scale, float half) {
    const float centre = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c), r_hm1), 1.0f), scale);
    const float jitter = __fmul_rn(__fsub_rn(__fmul_rn(u, 2.0f), 1.0f), half);
    return __fadd_rn(centre, jitter);
}

template <bool H2>
__global__ void __launch_bounds__(kFieldThreads, NSIG_FWD_MINB)
k_grid_sweep(const SweepParams p) {
    constexpr int MT = 2;
    extern __shared__ __align__(16) __half sm[];
    __shared__ double s_sum[kFieldWarps];
    stage_forward_weights(sm, p.f.sigma_w, nullptr, false);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    [[maybe_unused]] uint32_t* feat_s = reinterpret_cast<uint32_t*>(sm + kFwdHalfsPad) + warp * kFeatWords;
    const uint32_t H3 = p.H * p.H * p.H;
    const uint32_t tiles_per_cas = div_up(p.n, 32u);
    const uint32_t total = p.C * tiles_per_cas;
    const float r_hm1 = __fdiv_rn(1.0f, (float)(p.H - 1));
    double acc = 0.0;
    for (uint32_t wt = blockIdx.x * kFieldWarps + warp; wt < total; wt += gridDim.x * kFieldWarps) {
        const uint32_t cas = wt / tiles_per_cas, i0 = (wt - cas * tiles_per_cas) * 32u;
        const double bound_c = fmin(ldexp(1.0, (int)cas), p.bound);   // min(2 ** cas, self.bound)
        const double half_d = bound_c / (double)p.H;
        const float scale = (float)(bound_c - half_d), half = (float)half_d;
        // position of cell i of this cascade (centre + jitter), normalised to the unit box
        auto cell_pos = [&](uint32_t i, uint32_t& idx, float (&pn)[3]) {
            idx = p.cells ? (uint32_t)p.cells[(size_t)cas * p.n + i] : i;
            float u[3];
            if (p.noise) {
                const float* nz = p.noise + ((size_t)cas * p.n + i) * 3;
                u[0] = __ldg(nz); u[1] = __ldg(nz + 1); u[2] = __ldg(nz + 2);
            } else {
                const uint4 r = philox4x32_10(make_uint4(i, cas, 0x6e736967u, 0u), make_uint2(p.seed_lo, p.seed_hi));
                u[0] = u01(r.x); u[1] = u01(r.y); u[2] = u01(r.z);
            }
            const uint32_t c3[3] = {morton3D_invert(idx), morton3D_invert(idx >> 1), morton3D_invert(idx >> 2)};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float x = cell_axis(c3[a], u[a], r_hm1, scale, half);
                pn[a] = __fmul_rn(__fadd_rn(x, p.f.bound_add), p.f.bound_mul);  // network_wtmk_tcnn.py:129
            }
        };
        uint32_t fa[MT][2][4];
        uint32_t cell[MT][2];  // cells of the rows whose sigma this thread ends up holding
#ifndef NSIG_GATHER_V2
        {
            float xn[MT][2][3];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) cell_pos(min(i0 + mt * 16 + h * 8 + g, p.n - 1), cell[mt][h], xn[mt][h]);
            encode_positions<MT, H2>(fa, p.f, xn, g, tig);
        }
#else
        {   // lane r owns cell i0 + r of the tile: position -> warp-cooperative gather
            uint32_t idx;
            float pn[3];
            cell_pos(min(i0 + (uint32_t)lane, p.n - 1), idx, pn);
            gather_tile<H2>(fa, p.f, pn[0], pn[1], pn[2], feat_s, lane);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t i = min(i0 + mt * 16 + h * 8 + g, p.n - 1);
                    cell[mt][h] = p.cells ? (uint32_t)p.cells[(size_t)cas * p.n + i] : i;
                }
        }
#endif
        uint32_t h1[MT][4][4];
        {
            float c[MT][8][4];
            layer<MT, 2, 8>(c, fa, sm + oWs0, kS32, g, tig);
            relu_to_a<MT, 8>(h1, c);
        }
        float so[MT][1][4];
        layer<MT, 4, 1, 1>(so, h1, sm + oWs1, kS64, g, tig);  // only the n-tile holding the logit (permuted column 15)
        if (tig == 3) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (i0 + mt * 16 + h * 8 + g >= p.n) continue;
                    const float s = __fmul_rn(p.f.density_scale, expf(so[mt][0][2 * h + 1]));
                    const size_t at = (size_t)cas * H3 + cell[mt][h];
                    if (p.tmp) {  // partial update: a cell may be drawn several times; keep the largest sample
                        if (s >= 0.0f) atomicMax(reinterpret_cast<int*>(p.tmp + at), __float_as_int(s));
                    } else {      // full update: every cell exactly once -> EMA in place (renderer_wtmk.py:521-523)
                        const float old = p.grid[at];
                        float v = old;
                        if (old >= 0.0f && s >= 0.0f) { v = fmaxf(__fmul_rn(old, p.decay), s); p.grid[at] = v; }
                        acc += (double)fmaxf(v, 0.0f);
                    }
                }
        }
    }
    if (p.sum && !p.tmp) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(NSIG_FULL_MASK, acc, o);
        if (lane == 0) s_sum[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kFieldWarps; ++w) t += s_sum[w];
            atomicAdd(p.sum, t);
        }
    }
}

// EMA/max + mean over the whole grid for the partial update (renderer_wtmk.py:521-524)
__global__ void __launch_bounds__(256)
k_grid_finalize(float* __restrict__ grid, const float* __restrict__ tmp, uint32_t n, float decay, double* __restrict__ sum) {
    __shared__ double s_sum[8];
    double acc = 0.0;
    for (uint32_t i = (blockIdx.x * 256u + threadIdx.x) * 4u; i < n; i += gridDim.x * 1024u) {
        if (i + 4 <= n) {
            float4 g4 = *reinterpret_cast<const float4*>(grid + i);
            const float4 t4 = ld_stream4(reinterpret_cast<const float4*>(tmp + i));
            float* gv = reinterpret_cast<float*>(&g4);
            const float* tv = reinterpret_cast<const float*>(&t4);
            bool dirty = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (gv[k] >= 0.0f && tv[k] >= 0.0f) { gv[k] = fmaxf(__fmul_rn(gv[k], decay), tv[k]); dirty = true; }
                acc += (double)fmaxf(gv[k], 0.0f);
            }
            if (dirty) *reinterpret_cast<float4*>(grid + i) = g4;
        } else {
            for (uint32_t k = i; k < n; ++k) {
                float gvv = grid[k];
                const float t = tmp[k];
                if (gvv >= 0.0f && t >= 0.0f) { gvv = fmaxf(__fmul_rn(gvv, decay), t); grid[k] = gvv; }
                acc += (double)fmaxf(gvv, 0.0f);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(NSIG_FULL_MASK, acc, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_sum[w];
        atomicAdd(sum, t);
    }
}

// packbits (raymarching.cu:268-289) with thresh = min(mean_density, density_thresh) read from the device
__global__ void __launch_bounds__(256)
k_grid_pack(const float* __restrict__ grid, uint32_t n_bytes, const double* __restrict__ sum, uint32_t n_cells,
            float density_thresh, uint8_t* __restrict__ bitfield, float* __restrict__ stats) {
    const float mean = (float)(*sum / (double)n_cells);
    const float thresh = fminf(mean, density_thresh);
    const uint32_t q = threadIdx.x + blockIdx.x * blockDim.x;
    if (q == 0 && stats) { stats[0] = mean; stats[1] = thresh; }
    const uint32_t n0 = q * 4;
    if (n0 >= n_bytes) return;
    if (n0 + 4 <= n_bytes) {
        const float4* g4 = reinterpret_cast<const float4*>(grid + (size_t)n0 * 8);
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float4 lo = ld_stream4(g4 + 2 * b), hi = ld_stream4(g4 + 2 * b + 1);
            const uint32_t bits = (lo.x > thresh ? 1u : 0u) | (lo.y > thresh ? 2u : 0u) | (lo.z > thresh ? 4u : 0u) |
                                  (lo.w > thresh ? 8u : 0u) | (hi.x > thresh ? 16u : 0u) | (hi.y > thresh ? 32u : 0u) |
                                  (hi.z > thresh ? 64u : 0u) | (hi.w > thresh ? 128u : 0u);
            word |= bits << (8 * b);
        }
        *reinterpret_cast<uint32_t*>(bitfield + n0) = word;
    } else {
        for (uint32_t n = n0; n < n_bytes; ++n) {
            uint32_t bits = 0;
            for (int i = 0; i < 8; ++i) bits |= (grid[(size_t)n * 8 + i] > thresh) ? (1u << i) : 0u;
            bitfield[n] = (uint8_t)bits;
        }
    }
}

// ---- partial-update cell selection (renderer_wtmk.py:489-501) ------------------------------------------------------
// occupied = nonzero(density_grid[cas] > 0) as an ordered compaction: per-block counts -> scan -> write.
constexpr uint32_t kOccBlock = 1024;  // cells per block

__global__ void __launch_bounds__(256)
k_occ_count(const float* __restrict__ grid, uint32_t H3, uint32_t* __restrict__ block_counts) {
    const uint32_t cas = blockIdx.y, base = blockIdx.x * kOccBlock + threadIdx.x * 4;
    uint32_t c = 0;
    if (base + 4 <= H3) {
        const float4 v = *reinterpret_cast<const float4*>(grid + (size_t)cas * H3 + base);
        c = (v.x > 0.f) + (v.y > 0.f) + (v.z > 0.f) + (v.w > 0.f);
    } else {
        for (uint32_t k = base; k < H3; ++k) c += grid[(size_t)cas * H3 + k] > 0.f;
    }
    const uint32_t total = __syncthreads_count(0) * 0u + c;  // (keeps `c` live across the barrier below)
    __shared__ uint32_t s[8];
    uint32_t w = total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(NSIG_FULL_MASK, w, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += s[i];
        block_counts[cas * gridDim.x + blockIdx.x] = t;
    }
}

// one CTA per cascade: exclusive scan of the block counts (in place), total -> n_occ[cas]
__global__ void __launch_bounds__(1024)
k_occ_scan(uint32_t* __restrict__ block_counts, uint32_t n_blocks, uint32_t* __restrict__ n_occ) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    uint32_t* bc = block_counts + (size_t)blockIdx.x * n_blocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? bc[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(NSIG_FULL_MASK, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(NSIG_FULL_MASK, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t incl = x + (warp ? s_warp[warp - 1] : 0u);
        if (i < n_blocks) bc[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_occ[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(256)
k_occ_write(const float* __restrict__ grid, uint32_t H3, const uint32_t* __restrict__ block_offsets,
            int32_t* __restrict__ occ) {
    const uint32_t cas = blockIdx.y, base = blockIdx.x * kOccBlock + threadIdx.x * 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool f[4] = {false, false, false, false};
    for (int k = 0; k < 4; ++k)
        if (base + k < H3) f[k] = grid[(size_t)cas * H3 + base + k] > 0.f;
    const uint32_t c = f[0] + f[1] + f[2] + f[3];
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(NSIG_FULL_MASK, x, o); if (lane >= o) x += y; }
    __shared__ uint32_t s[8];
    if (lane == 31) s[warp] = x;
    __syncthreads();
    uint32_t pre = 0;
    for (int w = 0; w < warp; ++w) pre += s[w];
    uint32_t at = block_offsets[cas * gridDim.x + blockIdx.x] + pre + x - c;
    int32_t* dst = occ + (size_t)cas * H3;
    for (int k = 0; k < 4; ++k)
        if (f[k]) dst[at++] = (int32_t)(base + k);
}

// cells[cas, 0:n_uniform)  = morton3D(randint(0, H, 3));   cells[cas, n_uniform:n_uniform+n_occupied) = occupied[randint]
__global__ void __launch_bounds__(256)
k_cells_sample(const int32_t* __restrict__ occ, const uint32_t* __restrict__ n_occ, uint32_t H, uint32_t n_uniform,
               uint32_t n_occupied, uint32_t seed_lo, uint32_t seed_hi, int32_t* __restrict__ cells) {
    const uint32_t cas = blockIdx.y, i = blockIdx.x * 256u + threadIdx.x, n = n_uniform + n_occupied;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(i, cas, 0x63656c6cu, 0u), make_uint2(seed_lo, seed_hi));
    const uint32_t H3 = H * H * H;
    uint32_t idx;
    const uint32_t cnt = n_occ[cas];
    if (i < n_uniform || cnt == 0) {
        idx = morton3D(__umulhi(r.x, H), __umulhi(r.y, H), __umulhi(r.z, H));
    } else {
        idx = (uint32_t)occ[(size_t)cas * H3 + __umulhi(r.w, cnt)];
    }
    cells[(size_t)cas * n + i] = (int32_t)idx;
}

// ---- mark_untrained_grid (renderer_wtmk.py:380-442) -----------------------------------------------------------------
constexpr int kPoseChunk = 64;

__global__ void __launch_bounds__(256)
k_mark_untrained(const float* __restrict__ poses, uint32_t B, float cx_fx, float cy_fy, uint32_t C, uint32_t H,
                 double bound, float* __restrict__ grid) {
    __shared__ float sp[kPoseChunk][12];  // R (row-major 3x3) then t
    const uint32_t H3 = H * H * H, cas = blockIdx.y, idx = blockIdx.x * 256u + threadIdx.x;
    const double bound_c = fmin(ldexp(1.0, (int)cas), bound);
    const double half_d = bound_c / (double)H;
    const float scale = (float)(bound_c - half_d), margin = (float)(half_d * 2.0);
    const float r_hm1 = __fdiv_rn(1.0f, (float)(H - 1));
    float w[3] = {0.f, 0.f, 0.f};
    if (idx < H3) {
        const uint32_t c3[3] = {morton3D_invert(idx), morton3D_invert(idx >> 1), morton3D_invert(idx >> 2)};
#pragma unroll
        for (int a = 0; a < 3; ++a)
            w[a] = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn(2.0f, (float)c3[a]), r_hm1), 1.0f), scale);
    }
    bool seen = false;
    for (uint32_t b0 = 0; b0 < B; b0 += kPoseChunk) {
        const uint32_t nb = min((uint32_t)kPoseChunk, B - b0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nb * 12; i += 256) {
            const uint32_t b = i / 12, e = i % 12;
            sp[b][e] = e < 9 ? poses[(size_t)(b0 + b) * 16 + (e / 3) * 4 + (e % 3)] : poses[(size_t)(b0 + b) * 16 + (e - 9) * 4 + 3];
        }
        __syncthreads();
        if (idx < H3 && !seen) {
            for (uint32_t b = 0; b < nb; ++b) {
                const float dx = __fsub_rn(w[0], sp[b][9]), dy = __fsub_rn(w[1], sp[b][10]), dz = __fsub_rn(w[2], sp[b][11]);
                // cam = d @ R : cam_j = sum_i d_i R[i][j]
                const float camx = fmaf(dz, sp[b][6], fmaf(dy, sp[b][3], __fmul_rn(dx, sp[b][0])));
                const float camy = fmaf(dz, sp[b][7], fmaf(dy, sp[b][4], __fmul_rn(dx, sp[b][1])));
                const float camz = fmaf(dz, sp[b][8], fmaf(dy, sp[b][5], __fmul_rn(dx, sp[b][2])));
                const bool in = camz > 0.0f && fabsf(camx) < __fadd_rn(__fmul_rn(cx_fx, camz), margin) &&
                                fabsf(camy) < __fadd_rn(__fmul_rn(cy_fy, camz), margin);
                if (in) { seen = true; break; }
            }
        }
    }
    if (idx < H3 && !seen) grid[(size_t)cas * H3 + idx] = -1.0f;
}

// ---- get_rays (utils_wtmk_disen.py:59-143) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_get_rays(const float* __restrict__ poses, uint32_t B, float r_fx, float r_fy, float cx, float cy, uint32_t W,
           const int64_t* __restrict__ inds, int64_t inds_batch_stride, uint32_t N, float* __restrict__ rays_o,
           float* __restrict__ rays_d) {
    const uint32_t n = blockIdx.x * 256u + threadIdx.x, b = blockIdx.y;
    if (n >= N) return;
    const uint32_t pix = inds ? (uint32_t)inds[(size_t)b * inds_batch_stride + n] : n;
    const float i = __fadd_rn((float)(pix % W), 0.5f), j = __fadd_rn((float)(pix / W), 0.5f);
    const float xs = __fmul_rn(__fsub_rn(i, cx), r_fx), ys = __fmul_rn(__fsub_rn(j, cy), r_fy);  // zs = 1
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xs, xs), __fmul_rn(ys, ys)), 1.0f));
    const float dx = __fdiv_rn(xs, nrm), dy = __fdiv_rn(ys, nrm), dz = __fdiv_rn(1.0f, nrm);
    const float* P = poses + (size_t)b * 16;
    float* o = rays_o + ((size_t)b * N + n) * 3;
    float* d = rays_d + ((size_t)b * N + n) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {  // directions @ R^T
        d[r] = fmaf(dz, P[r * 4 + 2], fmaf(dy, P[r * 4 + 1], __fmul_rn(dx, P[r * 4])));
        o[r] = P[r * 4 + 3];
    }
}

}  // namespace nsig

using namespace nsig;

extern "C" {

int nsig_grid_sweep(float* density_grid, float* tmp_grid, const int32_t* cells, uint32_t n, const float* noise,
                    uint64_t seed, uint32_t C, uint32_t H, double bound, float decay, const float* const* tables,
                    const float* resolutions, uint32_t log2_T, const float* S, float msg_resolution,
                    const void* sigma_w, float density_scale, double* sum, const void* const* tables_h2,
                    const float* h2_inv_scale, nsig_stream_t stream) {
    if (n == 0 || C == 0) return 0;
    if (!density_grid || H < 2 || H > 1024 || C > 31) return NSIG_EINVAL;
    if (!cells && n != H * H * H) return NSIG_EINVAL;
    SweepParams p;
    const int rc = fill_field_params(p.f, density_grid /*unused*/, nullptr, 0, (float)bound, tables, resolutions, log2_T,
                                     S, msg_resolution, sigma_w, nullptr, nullptr, density_scale, tables_h2, h2_inv_scale);
    if (rc) return rc;
    p.grid = density_grid; p.tmp = tmp_grid; p.cells = cells; p.noise = noise; p.n = n; p.C = C; p.H = H;
    p.bound = bound; p.decay = decay; p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.sum = sum;
    const size_t smem = kFieldFwdSmem;
    const uint32_t tiles = C * div_up(n, 32u);
    if (tables_h2)
        k_grid_sweep<true><<<field_grid(k_grid_sweep<true>, smem, tiles * 32u, kFieldWarps * 32), kFieldThreads, smem,
                             (cudaStream_t)stream>>>(p);
    else
        k_grid_sweep<false><<<field_grid(k_grid_sweep<false>, smem, tiles * 32u, kFieldWarps * 32), kFieldThreads, smem,
                              (cudaStream_t)stream>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_grid_finalize(float* density_grid, const float* tmp_grid, uint32_t n_cells, float decay, double* sum,
                       nsig_stream_t stream) {
    if (n_cells == 0) return 0;
    if (!density_grid || !tmp_grid || !sum) return NSIG_EINVAL;
    if ((((uintptr_t)density_grid) | ((uintptr_t)tmp_grid)) & 15) return NSIG_EINVAL;
    const uint32_t blocks = min(div_up(n_cells, 1024u), 148u * 8u);
    k_grid_finalize<<<blocks, 256, 0, (cudaStream_t)stream>>>(density_grid, tmp_grid, n_cells, decay, sum);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_grid_pack(const float* density_grid, uint32_t n_bytes, const double* sum, uint32_t n_cells,
                   float density_thresh, uint8_t* bitfield, float* stats, nsig_stream_t stream) {
    if (n_bytes == 0) return 0;
    if (!density_grid || !sum || !bitfield || n_cells == 0) return NSIG_EINVAL;
    if ((((uintptr_t)density_grid) & 15) || (((uintptr_t)bitfield) & 3)) return NSIG_EINVAL;
    k_grid_pack<<<div_up(div_up(n_bytes, 4u), 256u), 256, 0, (cudaStream_t)stream>>>(density_grid, n_bytes, sum, n_cells,
                                                                                    density_thresh, bitfield, stats);
    NSIG_LAUNCH_CHECK();
    return 0;
}

size_t nsig_grid_sample_cells_scratch_bytes(uint32_t C, uint32_t H) {
    const size_t H3 = (size_t)H * H * H;
    return (size_t)C * (div_up((uint32_t)H3, kOccBlock) + 1) * sizeof(uint32_t) + (size_t)C * H3 * sizeof(int32_t);
}

int nsig_grid_sample_cells(const float* density_grid, uint32_t C, uint32_t H, uint32_t n_uniform, uint32_t n_occupied,
                           uint64_t seed, int32_t* cells, void* scratch, nsig_stream_t stream) {
    if (C == 0 || n_uniform + n_occupied == 0) return 0;
    if (!density_grid || !cells || !scratch || H < 2 || H > 1024 || C > 31) return NSIG_EINVAL;
    if (((uintptr_t)density_grid) & 15) return NSIG_EINVAL;
    const uint32_t H3 = H * H * H, nblk = div_up(H3, kOccBlock);
    uint32_t* block_counts = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* n_occ = block_counts + (size_t)C * nblk;
    int32_t* occ = reinterpret_cast<int32_t*>(n_occ + C);
    cudaStream_t st = (cudaStream_t)stream;
    k_occ_count<<<dim3(nblk, C), 256, 0, st>>>(density_grid, H3, block_counts);
    NSIG_LAUNCH_CHECK();
    k_occ_scan<<<C, 1024, 0, st>>>(block_counts, nblk, n_occ);
    NSIG_LAUNCH_CHECK();
    k_occ_write<<<dim3(nblk, C), 256, 0, st>>>(density_grid, H3, block_counts, occ);
    NSIG_LAUNCH_CHECK();
    const uint32_t n = n_uniform + n_occupied;
    k_cells_sample<<<dim3(div_up(n, 256u), C), 256, 0, st>>>(occ, n_occ, H, n_uniform, n_occupied, (uint32_t)seed,
                                                            (uint32_t)(seed >> 32), cells);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_mark_untrained_grid(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t C,
                             uint32_t H, double bound, float* density_grid, nsig_stream_t stream) {
    if (C == 0) return 0;
    if (!density_grid || (B && !poses) || H < 2 || H > 1024 || C > 31) return NSIG_EINVAL;
    const uint32_t H3 = H * H * H;
    // cx / fx and cy / fy are python doubles rounded to fp32 by the tensor multiply (renderer_wtmk.py:430-431)
    k_mark_untrained<<<dim3(div_up(H3, 256u), C), 256, 0, (cudaStream_t)stream>>>(
        poses, B, (float)((double)cx / (double)fx), (float)((double)cy / (double)fy), C, H, bound, density_grid);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_get_rays(const float* poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                  const int64_t* inds, int64_t inds_batch_stride, uint32_t N, float* rays_o, float* rays_d,
                  nsig_stream_t stream) {
    if (B == 0 || N == 0) return 0;
    if (!poses || !rays_o || !rays_d || W == 0 || H == 0 || !(fx != 0.0f) || !(fy != 0.0f)) return NSIG_EINVAL;
    if (!inds && N != H * W) return NSIG_EINVAL;
    k_get_rays<<<dim3(div_up(N, 256u), B), 256, 0, (cudaStream_t)stream>>>(poses, B, 1.0f / fx, 1.0f / fy, cx, cy, W, inds,
                                                                          inds_batch_stride, N, rays_o, rays_d);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
