// field_tc.cu — the watermark-mode field backward on Blackwell's 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same mathematics as k_field_bwd<false,false> (field.cu): recompute the two 64-wide MLPs of
// nerf/network_wtmk_tcnn.py:52-88 from the saved encoder output, back-propagate d(sigma), d(rgb) through them and
// scatter d(loss)/dS - the gradient of feature channels 30,31 - into the pre-summed message table
// (hash_encoding_wtmk_bit.py:99-116).  What changes is where the ten small GEMMs of a tile run:
//
//   * a CTA of 128 threads owns a tile of 128 samples; thread t owns row t - TMEM lane t - from start to finish;
//   * every layer is ONE tcgen05.mma chain (M = 128, N in {16, 64}, K in {16, 32, 64}) issued by one thread:
//     the A operand (this layer's input activations, fp16) is written by the 128 row owners into shared memory in the
//     UMMA canonical K-major layout, the B operand (the layer's weights, staged once per CTA, both orientations) sits
//     next to it, the fp32 accumulator lives in 64 TMEM columns;
//   * tcgen05.commit -> mbarrier tells the row owners the accumulator is complete; each reads ITS OWN row back with
//     tcgen05.ld.32x32b, applies ReLU / the ReLU mask / sigmoid' / trunc_exp' in registers and writes the next layer's A
//     row.  ReLU masks are kept as 64-bit sign masks (2 registers per layer), so nothing but the row's scalars stays
//     live between layers;
//   * the chain of a tile is serial (10 dependent layers), so latency is hidden by running several CTAs per SM
//     (4 x 52 KB of shared memory, 4 x 64 TMEM columns), each on its own tile.
//
// mma.sync version for comparison (ncu, profiles/r01_experiments_v5.txt): tensor pipe 43 %, issue 58 %, top stall =
// dependent HMMA chains; every 16 x 8 x 16 MMA is a warp instruction plus two shared-memory B-fragment loads.  Here a
// whole layer is 1-4 instructions of one thread.
//
// Operand layout (no swizzle, "interleaved" 8 x 16-byte core matrices), matrix [R x K] halfs:
//     byte(r, k) = (r / 8) * (K * 16) + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2
// i.e. leading-dimension byte offset (K direction) = 128, stride byte offset (row-group direction) = K * 16.
// A row owner writes its row as K/8 16-byte chunks: a quarter warp covers 128 contiguous bytes - conflict free.
//
// Numerics: fp16 operands, fp32 accumulation (TMEM), per-row power-of-two gradient scaling as in field.cu.
#include "field_common.cuh"

#include <cstdlib>

namespace nsig {
namespace tc {

constexpr int kRows = 128;
constexpr uint32_t kTmemCols = 64;

// shared-memory map (bytes)
constexpr uint32_t oBs0 = 0;                    // [64 x 32]  Ws0            n = hidden,  k = feature
constexpr uint32_t oBs1 = oBs0 + 64 * 32 * 2;   // [16 x 64]  Ws1 permuted   n = [geo0..14, logit]
constexpr uint32_t oBc0 = oBs1 + 16 * 64 * 2;   // [64 x 32]  Wc0 (input 31 = 0)
constexpr uint32_t oBc1 = oBc0 + 64 * 32 * 2;   // [64 x 64]  Wc1
constexpr uint32_t oBc2 = oBc1 + 64 * 64 * 2;   // [16 x 64]  Wc2 (rows >= 3 zero)
constexpr uint32_t oBc2T = oBc2 + 16 * 64 * 2;  // [64 x 16]  n = h2,  k = rgb output (k >= 3 zero)
constexpr uint32_t oBc1T = oBc2T + 64 * 16 * 2; // [64 x 64]  n = h1,  k = h2
constexpr uint32_t oBc0T = oBc1T + 64 * 64 * 2; // [16 x 64]  n = geo (15 = 0), k = h1
constexpr uint32_t oBs1T = oBc0T + 16 * 64 * 2; // [64 x 16]  n = h1s, k = [geo0..14, logit]
constexpr uint32_t oBs0T = oBs1T + 64 * 16 * 2; // [16 x 64]  n = feature 16..31, k = h1s
constexpr uint32_t oA = oBs0T + 16 * 64 * 2;    // [128 x 64] activation operand of the current layer
constexpr uint32_t kSmemBytes = oA + kRows * 64 * 2;

__host__ __device__ constexpr uint32_t canon(uint32_t r, uint32_t k, uint32_t K) {
    return (r >> 3) * (K << 4) + (k >> 3) * 128u + (r & 7u) * 16u + (k & 7u) * 2u;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: start address, LBO, SBO in 16-byte units, descriptor version 1, no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t K) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(((K << 4) >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

// D[128 x N] = A[128 x K] * B[N x K]^T : K/16 instructions of the calling thread, then commit to the mbarrier
template <int N, int K>
__device__ __forceinline__ void issue_layer(uint32_t sA, uint32_t sB, uint32_t tmem_d, uint32_t mbar) {
    // instruction descriptor: D = fp32 (bit 4), A = B = fp16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
#pragma unroll
    for (int k = 0; k < K / 16; ++k)
        mma_f16_ss(tmem_d, make_desc(sA + k * 256, K), make_desc(sB + k * 256, K), idesc, k > 0 ? 1u : 0u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

#define NSIG_TMEM_LD16(taddr, v)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"  \
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),         \
                   "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])    \
                 : "r"(taddr))

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();   // a lost completion must fail loudly, never hang the GPU
    }
}

// one 16-byte chunk (8 halfs) of this thread's row of the A operand
__device__ __forceinline__ void store_a(uint8_t* sA, int t, int chunk, int K, uint4 v) {
    *reinterpret_cast<uint4*>(sA + (t >> 3) * (K << 4) + chunk * 128 + (t & 7) * 16) = v;
}

// stage a [R x K] operand matrix; src(r, k) yields the element
template <typename F>
__device__ __forceinline__ void stage(uint8_t* dst, int R, int K, F src) {
    for (int i = threadIdx.x; i < R * K; i += kRows) {
        const int r = i / K, k = i - r * K;
        *reinterpret_cast<__half*>(dst + canon(r, k, K)) = src(r, k);
    }
}

struct BwdTcParams {
    const float* xyzs;
    const float* dirs;
    uint32_t M;
    float bound_add, bound_mul;
    const __half* feat;
    const float* grad_sigmas;
    const float* grad_rgbs;
    const __half* sigma_w;
    const __half* color_w;
    float msg_grid_size;
    uint32_t mask;
    float* G;
    const int32_t* M_dev;
    float density_scale;
};

struct RowIn {
    uint4 f[4];      // saved encoder output of the row: 32 halfs
    float dv[3];     // view direction
    float gs, gc[3]; // incoming gradients
    float sx[3];     // position (for the scatter)
};

__global__ void __launch_bounds__(kRows, 4)
k_field_bwd_tc(const BwdTcParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t mbar_storage;
    __shared__ uint32_t tmem_slot;
    uint32_t M = p.M;
    if (p.M_dev) M = min(M, (uint32_t)max(*p.M_dev, 0));
    if (M == 0) return;
    const int t = threadIdx.x, warp = t >> 5;
    const __half hz = __float2half(0.0f);
    // ---- weights, both orientations, in operand layout (once per persistent CTA) ----
    {
        const __half* sw = p.sigma_w;
        const __half* cw = p.color_w;
        stage(sm + oBs0, 64, 32, [&](int n, int k) { return sw[n * 32 + k]; });
        stage(sm + oBs1, 16, 64, [&](int n, int k) { return sw[2048 + ((n + 1) & 15) * 64 + k]; });
        stage(sm + oBc0, 64, 32, [&](int n, int k) { return k == 31 ? hz : cw[n * 32 + k]; });
        stage(sm + oBc1, 64, 64, [&](int n, int k) { return cw[2048 + n * 64 + k]; });
        stage(sm + oBc2, 16, 64, [&](int n, int k) { return n < 3 ? cw[6144 + n * 64 + k] : hz; });
        stage(sm + oBc2T, 64, 16, [&](int n, int k) { return k < 3 ? cw[6144 + k * 64 + n] : hz; });
        stage(sm + oBc1T, 64, 64, [&](int n, int k) { return cw[2048 + k * 64 + n]; });
        stage(sm + oBc0T, 16, 64, [&](int n, int k) { return n < 15 ? cw[k * 32 + 16 + n] : hz; });
        stage(sm + oBs1T, 64, 16, [&](int n, int k) { return sw[2048 + ((k + 1) & 15) * 64 + n]; });
        stage(sm + oBs0T, 16, 64, [&](int n, int k) { return sw[k * 32 + 16 + n]; });
    }
    const uint32_t mbar = smem_u32(&mbar_storage);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t my_tmem = tmem_d + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
    const uint32_t sbase = smem_u32(sm);
    uint8_t* sA = sm + oA;
    uint32_t parity = 0;

    // hand the A operand written by all row owners to the tensor core, run one layer, wait for its accumulator
    auto run_layer = [&](auto issue) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); // my TMEM reads are done before the barrier
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue();
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    const uint32_t n_tiles = div_up(M, (uint32_t)kRows);
    auto load_row = [&](uint32_t tile, RowIn& in) {
        const uint32_t r = tile * kRows + t;
        const uint32_t rc = min(r, M - 1);
        const uint4* f = reinterpret_cast<const uint4*>(p.feat + (size_t)rc * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) in.f[c] = __ldg(f + c);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            in.dv[a] = __ldg(p.dirs + (size_t)rc * 3 + a);
            in.sx[a] = __ldg(p.xyzs + (size_t)rc * 3 + a);
        }
        const bool live = r < M;
        in.gs = live ? __ldg(p.grad_sigmas + r) : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) in.gc[a] = live ? __ldg(p.grad_rgbs + (size_t)r * 3 + a) : 0.f;
    };

    RowIn nxt;
    if (blockIdx.x < n_tiles) load_row(blockIdx.x, nxt);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RowIn cur = nxt;
        if (tile + gridDim.x < n_tiles) load_row(tile + gridDim.x, nxt);   // next tile's inputs fly during this tile's chain
        const bool has_grad = (cur.gs != 0.f) | (cur.gc[0] != 0.f) | (cur.gc[1] != 0.f) | (cur.gc[2] != 0.f);
        if (!__syncthreads_or(has_grad)) continue;   // padding / fully terminated tile

        float v[32];
        uint32_t m1s[2], m1c[2], m2c[2];
        // ReLU, remember the sign mask, round to fp16 and write the row as the next A operand (K = 64); the 64 accumulator
        // columns are read in two halves of 32 to keep the register footprint at 4 CTAs per SM
        auto relu_store = [&](uint32_t (&msk)[2]) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                NSIG_TMEM_LD16(my_tmem + hf * 32, v);
                NSIG_TMEM_LD16(my_tmem + hf * 32 + 16, (&v[16]));
                tmem_wait_ld();
                uint32_t bits = 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = c * 8 + 2 * j;
                        const float a = fmaxf(v[i], 0.f), b = fmaxf(v[i + 1], 0.f);
                        bits |= (a > 0.f ? 1u : 0u) << i;
                        bits |= (b > 0.f ? 1u : 0u) << (i + 1);
                        w[j] = pack_h2(a, b);
                    }
                    store_a(sA, t, hf * 4 + c, 64, make_uint4(w[0], w[1], w[2], w[3]));
                }
                msk[hf] = bits;
            }
        };
        // gradient accumulators x ReLU mask -> fp16 A operand (K = 64)
        auto mask_store = [&](const uint32_t (&msk)[2]) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                NSIG_TMEM_LD16(my_tmem + hf * 32, v);
                NSIG_TMEM_LD16(my_tmem + hf * 32 + 16, (&v[16]));
                tmem_wait_ld();
                const uint32_t bits = msk[hf];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = c * 8 + 2 * j;
                        const float a = ((bits >> i) & 1u) ? v[i] : 0.f;
                        const float b = ((bits >> (i + 1)) & 1u) ? v[i + 1] : 0.f;
                        w[j] = pack_h2(a, b);
                    }
                    store_a(sA, t, hf * 4 + c, 64, make_uint4(w[0], w[1], w[2], w[3]));
                }
            }
        };

        // ---- forward recompute ----
        // L1: h1s = relu(feat x Ws0^T)
#pragma unroll
        for (int c = 0; c < 4; ++c) store_a(sA, t, c, 32, cur.f[c]);
        run_layer([&] { issue_layer<64, 32>(sbase + oA, sbase + oBs0, tmem_d, mbar); });
        // L2: [geo0..14, logit] = h1s x Ws1'^T
        relu_store(m1s);
        run_layer([&] { issue_layer<16, 64>(sbase + oA, sbase + oBs1, tmem_d, mbar); });
        // L3: h1c = relu([SH4(d), geo, 0] x Wc0^T)
        NSIG_TMEM_LD16(my_tmem, v);
        tmem_wait_ld();
        const float logit = v[15];
        {
            float sh[16];
            sh4(cur.dv[0], cur.dv[1], cur.dv[2], sh);
            store_a(sA, t, 0, 32, make_uint4(pack_h2(sh[0], sh[1]), pack_h2(sh[2], sh[3]), pack_h2(sh[4], sh[5]), pack_h2(sh[6], sh[7])));
            store_a(sA, t, 1, 32, make_uint4(pack_h2(sh[8], sh[9]), pack_h2(sh[10], sh[11]), pack_h2(sh[12], sh[13]), pack_h2(sh[14], sh[15])));
            store_a(sA, t, 2, 32, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
            store_a(sA, t, 3, 32, make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], 0.f)));
        }
        run_layer([&] { issue_layer<64, 32>(sbase + oA, sbase + oBc0, tmem_d, mbar); });
        // L4: h2c = relu(h1c x Wc1^T)
        relu_store(m1c);
        run_layer([&] { issue_layer<64, 64>(sbase + oA, sbase + oBc1, tmem_d, mbar); });
        // L5: rgb logits = h2c x Wc2^T
        relu_store(m2c);
        run_layer([&] { issue_layer<16, 64>(sbase + oA, sbase + oBc2, tmem_d, mbar); });

        // ---- output-activation gradients, normalised per row by a power of two (fp16 range) ----
        NSIG_TMEM_LD16(my_tmem, v);
        tmem_wait_ld();
        float d_rgb[3], d_logit, inv_scale;
        {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float s = 1.0f / (1.0f + expf(-v[a]));          // sigmoid'(z) = s (1 - s)
                d_rgb[a] = cur.gc[a] * s * (1.0f - s);
            }
            // trunc_exp backward: g * exp(clamp(x, -15, 15))  (activation.py:14-16)
            d_logit = cur.gs * p.density_scale * expf(fminf(fmaxf(logit, -15.0f), 15.0f));
            const float vmax = fmaxf(fmaxf(fabsf(d_rgb[0]), fabsf(d_rgb[1])), fmaxf(fabsf(d_rgb[2]), fabsf(d_logit)));
            float sc = 1.0f;
            if (vmax > 0.0f && isfinite(vmax)) { int e; frexpf(vmax, &e); sc = scalbnf(1.0f, -max(-100, min(100, e))); }
            inv_scale = 1.0f / sc;
            d_rgb[0] *= sc; d_rgb[1] *= sc; d_rgb[2] *= sc; d_logit *= sc;
        }
        // ---- colour net dgrad ----
        // B5: d h2 = d out x Wc2
        store_a(sA, t, 0, 16, make_uint4(pack_h2(d_rgb[0], d_rgb[1]), pack_h2(d_rgb[2], 0.f), 0u, 0u));
        store_a(sA, t, 1, 16, make_uint4(0u, 0u, 0u, 0u));
        run_layer([&] { issue_layer<64, 16>(sbase + oA, sbase + oBc2T, tmem_d, mbar); });
        // B4: d h1 = (d h2 . relu') x Wc1
        mask_store(m2c);
        run_layer([&] { issue_layer<64, 64>(sbase + oA, sbase + oBc1T, tmem_d, mbar); });
        // B3: d geo = ((d h1 . relu') x Wc0)[:, 16:31]
        mask_store(m1c);
        run_layer([&] { issue_layer<16, 64>(sbase + oA, sbase + oBc0T, tmem_d, mbar); });
        // ---- sigma net dgrad: d out' = [d geo0..14, d logit] ----
        NSIG_TMEM_LD16(my_tmem, v);
        tmem_wait_ld();
        store_a(sA, t, 0, 16, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
        store_a(sA, t, 1, 16, make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], d_logit)));
        run_layer([&] { issue_layer<64, 16>(sbase + oA, sbase + oBs1T, tmem_d, mbar); });
        // B1: d feat[16..31] = (d h1s . relu') x Ws0[:, 16:32]; the message feature was ADDED to channels 30,31
        mask_store(m1s);
        run_layer([&] { issue_layer<16, 64>(sbase + oA, sbase + oBs0T, tmem_d, mbar); });
        NSIG_TMEM_LD16(my_tmem, v);
        tmem_wait_ld();
        const float gx = v[14] * inv_scale, gy = v[15] * inv_scale;
        // ---- scatter into G: this thread's row, 8 corners (hash_encoding_wtmk_bit.py backward through the trilerp) ----
        if (p.G && (gx != 0.f || gy != 0.f)) {
            float xn[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) xn[a] = __fmul_rn(__fadd_rn(cur.sx[a], p.bound_add), p.bound_mul);
            const Voxel vx = locate(xn[0], xn[1], xn[2], p.msg_grid_size);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                red_add_v2(p.G + (size_t)corner_slot(vx, k, p.mask) * 2, corner_grad(vx, k, gx), corner_grad(vx, k, gy));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols));
}


// ---------------------------------------------------------------------------------------------------------------
// The same tcgen05 backward fed by SAVED ReLU MASKS (k_field_fwd's mask_out, see k_field_bwd_masks in field.cu): no MLP
// recomputation, five UMMA layers per 128-row tile instead of ten, 32 KB of shared memory per CTA (backward weights + one
// A operand) so six CTAs - six independent tiles - are resident per SM to hide the per-layer synchronisation.
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t mBc2T = 0;                     // [64 x 16]
constexpr uint32_t mBc1T = mBc2T + 64 * 16 * 2;   // [64 x 64]
constexpr uint32_t mBc0T = mBc1T + 64 * 64 * 2;   // [16 x 64]
constexpr uint32_t mBs1T = mBc0T + 16 * 64 * 2;   // [64 x 16]
constexpr uint32_t mBs0T = mBs1T + 64 * 16 * 2;   // [16 x 64]
constexpr uint32_t mA = mBs0T + 16 * 64 * 2;      // [128 x 64]
constexpr uint32_t kSmemBytesMasks = mA + kRows * 64 * 2;
constexpr int kMaskCtasPerSm = 6;

struct BwdTcMaskParams {
    const float* xyzs;
    uint32_t M;
    float bound_add, bound_mul;
    const uint4* masks;        // [M][2] x 16 bytes: the row's four (quad-thread) entries {m1s | m1c << 8, m2c}
    const float* sigmas;
    const float* rgbs;
    const float* grad_sigmas;
    const float* grad_rgbs;
    const __half* sigma_w;
    const __half* color_w;
    float msg_grid_size;
    uint32_t mask;
    float* G;
    const int32_t* M_dev;
    float density_scale;
};

// sign bit of hidden unit u of a layer (act_mask_bits layout): entry q = (u & 7) >> 1, bit 16*(u & 1) + (u >> 3) + shift
template <int U>
__device__ __forceinline__ bool unit_active(const uint32_t (&w)[4], int shift) {
    return (w[(U & 7) >> 1] >> (16 * (U & 1) + (U >> 3) + shift)) & 1u;
}

__global__ void __launch_bounds__(kRows, kMaskCtasPerSm)
k_field_bwd_tc_masks(const BwdTcMaskParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t mbar_storage;
    __shared__ uint32_t tmem_slot;
    uint32_t M = p.M;
    if (p.M_dev) M = min(M, (uint32_t)max(*p.M_dev, 0));
    if (M == 0) return;
    const int t = threadIdx.x, warp = t >> 5;
    const __half hz = __float2half(0.0f);
    {
        const __half* sw = p.sigma_w;
        const __half* cw = p.color_w;
        stage(sm + mBc2T, 64, 16, [&](int n, int k) { return k < 3 ? cw[6144 + k * 64 + n] : hz; });
        stage(sm + mBc1T, 64, 64, [&](int n, int k) { return cw[2048 + k * 64 + n]; });
        stage(sm + mBc0T, 16, 64, [&](int n, int k) { return n < 15 ? cw[k * 32 + 16 + n] : hz; });
        stage(sm + mBs1T, 64, 16, [&](int n, int k) { return sw[2048 + ((k + 1) & 15) * 64 + n]; });
        stage(sm + mBs0T, 16, 64, [&](int n, int k) { return sw[k * 32 + 16 + n]; });
    }
    const uint32_t mbar = smem_u32(&mbar_storage);
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    const uint32_t my_tmem = tmem_d + ((uint32_t)(warp * 32) << 16);
    const uint32_t sbase = smem_u32(sm);
    uint8_t* sA = sm + mA;
    uint32_t parity = 0;
    auto run_layer = [&](auto issue) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue();
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    const float inv_ds = 1.0f / p.density_scale;
    const uint32_t n_tiles = div_up(M, (uint32_t)kRows);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t r = tile * kRows + t;
        const bool live = r < M;
        float gs = 0.f, gc[3] = {0.f, 0.f, 0.f}, sig = 1.f, rgb[3] = {0.f, 0.f, 0.f}, sx[3] = {0.f, 0.f, 0.f};
        uint32_t w1[4] = {0u, 0u, 0u, 0u}, w2[4] = {0u, 0u, 0u, 0u};   // w1 = m1s | m1c << 8, w2 = m2c, per quad entry
        if (live) {
            gs = __ldg(p.grad_sigmas + r);
            sig = __ldg(p.sigmas + r);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                gc[a] = __ldg(p.grad_rgbs + (size_t)r * 3 + a);
                rgb[a] = __ldg(p.rgbs + (size_t)r * 3 + a);
                sx[a] = __ldg(p.xyzs + (size_t)r * 3 + a);
            }
            const uint4 m0 = __ldg(p.masks + (size_t)r * 2), m1 = __ldg(p.masks + (size_t)r * 2 + 1);
            w1[0] = m0.x; w2[0] = m0.y; w1[1] = m0.z; w2[1] = m0.w;
            w1[2] = m1.x; w2[2] = m1.y; w1[3] = m1.z; w2[3] = m1.w;
        }
        const bool has_grad = (gs != 0.f) | (gc[0] != 0.f) | (gc[1] != 0.f) | (gc[2] != 0.f);
        if (!__syncthreads_or(has_grad)) continue;

        float v[32];
        // gradient accumulators x saved ReLU mask -> fp16 A operand (K = 64), two halves of 32 columns
        auto mask_store = [&](const uint32_t (&w)[4], int shift) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                NSIG_TMEM_LD16(my_tmem + hf * 32, v);
                NSIG_TMEM_LD16(my_tmem + hf * 32 + 16, (&v[16]));
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int u = hf * 32 + c * 8 + 2 * j;     // hidden units u, u + 1: same quad entry, bits nt and 16 + nt
                        const uint32_t sel = (w[(u & 7) >> 1] >> ((u >> 3) + shift)) & 0x00010001u;
                        o[j] = pack_h2(v[c * 8 + 2 * j], v[c * 8 + 2 * j + 1]) & (sel * 0xffffu);
                    }
                    store_a(sA, t, hf * 4 + c, 64, make_uint4(o[0], o[1], o[2], o[3]));
                }
            }
        };
        // ---- output-activation gradients from the forward's outputs, normalised per row by a power of two ----
        float d_rgb[3], d_logit, inv_scale;
        {
#pragma unroll
            for (int a = 0; a < 3; ++a) d_rgb[a] = gc[a] * rgb[a] * (1.0f - rgb[a]);   // sigmoid'(z) = s (1 - s)
            // trunc_exp backward: g * exp(clamp(x, -15, 15)) with exp(x) = sigma / density_scale
            d_logit = gs * p.density_scale * fminf(fmaxf(sig * inv_ds, 3.0590232050182579e-7f), 3269017.3724721107f);
            const float vmax = fmaxf(fmaxf(fabsf(d_rgb[0]), fabsf(d_rgb[1])), fmaxf(fabsf(d_rgb[2]), fabsf(d_logit)));
            float sc = 1.0f;
            if (vmax > 0.0f && isfinite(vmax)) { int e; frexpf(vmax, &e); sc = scalbnf(1.0f, -max(-100, min(100, e))); }
            inv_scale = 1.0f / sc;
            d_rgb[0] *= sc; d_rgb[1] *= sc; d_rgb[2] *= sc; d_logit *= sc;
        }
        // B5: d h2 = d out x Wc2
        store_a(sA, t, 0, 16, make_uint4(pack_h2(d_rgb[0], d_rgb[1]), pack_h2(d_rgb[2], 0.f), 0u, 0u));
        store_a(sA, t, 1, 16, make_uint4(0u, 0u, 0u, 0u));
        run_layer([&] { issue_layer<64, 16>(sbase + mA, sbase + mBc2T, tmem_d, mbar); });
        // B4: d h1 = (d h2 . relu') x Wc1
        mask_store(w2, 0);
        run_layer([&] { issue_layer<64, 64>(sbase + mA, sbase + mBc1T, tmem_d, mbar); });
        // B3: d geo = ((d h1 . relu') x Wc0)[:, 16:31]
        mask_store(w1, 8);
        run_layer([&] { issue_layer<16, 64>(sbase + mA, sbase + mBc0T, tmem_d, mbar); });
        // B2: d h1s = [d geo0..14, d logit] x Ws1'
        NSIG_TMEM_LD16(my_tmem, v);
        tmem_wait_ld();
        store_a(sA, t, 0, 16, make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7])));
        store_a(sA, t, 1, 16, make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], d_logit)));
        run_layer([&] { issue_layer<64, 16>(sbase + mA, sbase + mBs1T, tmem_d, mbar); });
        // B1: d feat[16..31] = (d h1s . relu') x Ws0[:, 16:32]
        mask_store(w1, 0);
        run_layer([&] { issue_layer<16, 64>(sbase + mA, sbase + mBs0T, tmem_d, mbar); });
        NSIG_TMEM_LD16(my_tmem, v);
        tmem_wait_ld();
        const float gx = v[14] * inv_scale, gy = v[15] * inv_scale;
        if (live && (gx != 0.f || gy != 0.f)) {
            float xn[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) xn[a] = __fmul_rn(__fadd_rn(sx[a], p.bound_add), p.bound_mul);
            const Voxel vx = locate(xn[0], xn[1], xn[2], p.msg_grid_size);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                red_add_v2(p.G + (size_t)corner_slot(vx, k, p.mask) * 2, corner_grad(vx, k, gx), corner_grad(vx, k, gy));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTmemCols));
}

}  // namespace tc
}  // namespace nsig

using namespace nsig;

extern "C" int nsig_field_backward_tc(const float* xyzs, const float* dirs, uint32_t M, float bound, const void* feat,
                                      const float* grad_sigmas, const float* grad_rgbs, const void* sigma_w,
                                      const void* color_w, float density_scale, const int32_t* M_dev,
                                      float msg_resolution, uint32_t log2_T, float* G, nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!xyzs || !dirs || !feat || !grad_sigmas || !grad_rgbs || !sigma_w || !color_w || !G) return NSIG_EINVAL;
    if (((uintptr_t)feat) & 15) return NSIG_EINVAL;   // rows are read as four 16-byte chunks
    if (log2_T < 1 || log2_T > 30 || !(bound > 0.0f) || !(msg_resolution > 0.0f)) return NSIG_EINVAL;
    tc::BwdTcParams p;
    p.xyzs = xyzs; p.dirs = dirs; p.M = M;
    p.bound_add = bound; p.bound_mul = 1.0f / (2.0f * bound);
    p.feat = reinterpret_cast<const __half*>(feat);
    p.grad_sigmas = grad_sigmas; p.grad_rgbs = grad_rgbs;
    p.sigma_w = reinterpret_cast<const __half*>(sigma_w);
    p.color_w = reinterpret_cast<const __half*>(color_w);
    p.msg_grid_size = 1.0f / msg_resolution;
    p.mask = (1u << log2_T) - 1u;
    p.G = G; p.M_dev = M_dev; p.density_scale = density_scale;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tc::k_field_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        attr_set = true;
    }
    // Resident CTAs per SM: 4 by shared memory (4 x 52 KB), registers (__launch_bounds__(128, 4)) and TMEM (4 x 64 of 512
    // columns).  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for a kernel that allocates TMEM (it cannot see the
    // column count and assumes the whole 512), which left 3/4 of every SM idle in the first measurement (ncu: grid 148,
    // 6 % warps active, 432 us); tcgen05.alloc blocks until columns are free, so over-subscription is safe.
    static const int per_sm = [] { const char* e = getenv("NSIG_TC_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 4; }();
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t tiles = div_up(M, (uint32_t)tc::kRows);
    const uint32_t cap = (uint32_t)(sms * per_sm);
    tc::k_field_bwd_tc<<<tiles < cap ? tiles : cap, tc::kRows, tc::kSmemBytes, (cudaStream_t)stream>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsig_field_backward_tc_masks(const float* xyzs, uint32_t M, float bound, const void* masks,
                                            const float* sigmas, const float* rgbs, const float* grad_sigmas,
                                            const float* grad_rgbs, const void* sigma_w, const void* color_w,
                                            float density_scale, const int32_t* M_dev, float msg_resolution,
                                            uint32_t log2_T, float* G, nsig_stream_t stream) {
    if (M == 0) return 0;
    if (!xyzs || !masks || !sigmas || !rgbs || !grad_sigmas || !grad_rgbs || !sigma_w || !color_w || !G) return NSIG_EINVAL;
    if (((uintptr_t)masks) & 15) return NSIG_EINVAL;   // a row's masks are read as two 16-byte chunks
    if (log2_T < 1 || log2_T > 30 || !(bound > 0.0f) || !(msg_resolution > 0.0f) || !(density_scale > 0.0f)) return NSIG_EINVAL;
    tc::BwdTcMaskParams p;
    p.xyzs = xyzs; p.M = M; p.bound_add = bound; p.bound_mul = 1.0f / (2.0f * bound);
    p.masks = reinterpret_cast<const uint4*>(masks);
    p.sigmas = sigmas; p.rgbs = rgbs; p.grad_sigmas = grad_sigmas; p.grad_rgbs = grad_rgbs;
    p.sigma_w = reinterpret_cast<const __half*>(sigma_w);
    p.color_w = reinterpret_cast<const __half*>(color_w);
    p.msg_grid_size = 1.0f / msg_resolution;
    p.mask = (1u << log2_T) - 1u;
    p.G = G; p.M_dev = M_dev; p.density_scale = density_scale;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(tc::k_field_bwd_tc_masks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytesMasks);
        attr_set = true;
    }
    // resident CTAs per SM chosen by hand: the occupancy API answers 1 for kernels that allocate TMEM (see above)
    static const int per_sm = [] { const char* e = getenv("NSIG_TC_CTAS_PER_SM"); const int v = e ? atoi(e) : 0;
                                   return v > 0 ? v : tc::kMaskCtasPerSm; }();
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t tiles = div_up(M, (uint32_t)tc::kRows);
    const uint32_t cap = (uint32_t)(sms * per_sm);
    tc::k_field_bwd_tc_masks<<<tiles < cap ? tiles : cap, tc::kRows, tc::kSmemBytesMasks, (cudaStream_t)stream>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}
