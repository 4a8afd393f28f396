// decoder.cu — the HiDDeN message decoder (nerf/hidden_models.py:16-35, 104-137) forward AND backward as a short
// chain of tensor-core kernels.
//
// In the reference the decoder is ~30 nn.Module calls per direction; on [message_dim, 64, 12..48, 12..48] tensors every
// cuDNN/ATen launch is a 3-8 us latency-bound kernel and the ~270 of them are 60 % of the watermark step once the field
// path is fast (profiles/r01_launches_v2.txt).  Here:
//
//   * activations live in fp16 NHWC; a layer's BatchNorm (batch statistics) + GELU are never materialised: the conv
//     kernel of layer l writes the raw conv output z_l and its per-channel sum / sum of squares (fp64 atomics), and
//     every consumer of a_l = GELU(BN(z_l)) - the next conv, the weight-gradient kernel - applies the transform while
//     staging its input tile into shared memory;
//   * the backward of BN + GELU is folded the same way: dz_l = gamma*rstd*(dy - mean(dy) - yhat*mean(dy*yhat)) with
//     dy = da_l * GELU'(y) is computed on the fly from (da_l, z_l) by the data-gradient conv (the same implicit-GEMM
//     kernel with rotated weights) and by the weight-gradient kernel; one small reduction kernel per layer provides
//     the two means (which are also dbeta and dgamma);
//   * 3x3 convolutions are implicit GEMMs on mma.sync.m16n8k16 (fp16 operands, fp32 accumulate): M = pixels of an image
//     strip (staged with its halo), N = output channels, K = 9 taps x input channels; A fragments by ldmatrix from the
//     staged tile, weights in shared memory.  Weight gradients contract over pixels with ldmatrix.trans on both operands.
//
// Rounding points mirror torch.autocast(float16): conv outputs, BN outputs, GELU outputs and all activation gradients
// are rounded to fp16, statistics and parameter gradients are fp32/fp64.  Specialised for the reference's only decoder
// shape: 3x3 convs, `channels` = 64, num_bits*redundancy <= 8, BatchNorm eps 1e-3, exact (erf) GELU.
#include "nsig_common.cuh"

#include <cstdlib>
#include <mutex>

namespace nsig {

#ifndef NSIG_DEC_THREADS
#define NSIG_DEC_THREADS 256
#endif
constexpr int kDecThreads = NSIG_DEC_THREADS;
constexpr int kDecWarps = kDecThreads / 32;
constexpr float kBnEps = 1e-3f;  // hidden_models.py:24

enum InMode { IN_RAW = 0, IN_BNGELU = 1, IN_DZ = 2 };

// -DNSIG_DEC_TRACE (tools/build_variant.py, never in the shipped library): CTA 0 of every conv launch stamps %globaltimer at its
// phase boundaries; nsig_debug_dec_trace copies the stamps out (tools/dec_trace.py).
#ifdef NSIG_DEC_TRACE
__device__ unsigned long long g_dec_trace[64][12];
__device__ unsigned int g_dec_trace_n;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define DEC_TRACE_BEGIN() unsigned int tr_slot = 0; if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { tr_slot = atomicAdd(&g_dec_trace_n, 1u) & 63u; g_dec_trace[tr_slot][0] = gtime(); }
#define DEC_TRACE(k) if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_dec_trace[tr_slot][k] = gtime();
__device__ unsigned long long g_wg_trace[64][12];
__device__ unsigned int g_wg_trace_n;
#define WG_TRACE_BEGIN() unsigned int wg_slot = 0; if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { wg_slot = atomicAdd(&g_wg_trace_n, 1u) & 63u; g_wg_trace[wg_slot][0] = gtime(); g_wg_trace[wg_slot][11] = wg_slot; }
#define WG_TRACE(k) if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_wg_trace[wg_slot][k] = gtime();
// the final reduction of tap 0 runs in whichever group's CTA drew the last ticket: it stamps the most recent slot
#define WG_TRACE_LAST(k) if (blockIdx.x == 0 && threadIdx.x == 0) g_wg_trace[(g_wg_trace_n - 1u) & 63u][k] = gtime();
#else
#define DEC_TRACE_BEGIN()
#define DEC_TRACE(k)
#define WG_TRACE_BEGIN()
#define WG_TRACE(k)
#define WG_TRACE_LAST(k)
#endif

// per-channel constants of one BatchNorm layer, derived from the raw sums
struct BnCoef {
    float scale, shift;   // y = z*scale + shift  (scale = gamma*rstd, shift = beta - mean*scale)
    float mean, rstd;     // yhat = (z - mean)*rstd
    float c1, c2;         // mean(dy), mean(dy*yhat)            (IN_DZ only)
};

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}
// (Round 2, call AO: erfc by Abramowitz-Stegun 7.1.26 with one shared __expf - 14 instructions instead of erff's two divergent
// branches + expf, as close to the correctly rounded fp16 result as erff - was measured: decoder forward+backward 289.7 -> 286.9 us,
// step 0.9376 -> 0.9338 ms.  The conv kernels are not bound by these transforms; libdevice erff, i.e. torch's own arithmetic,
// stays.  nsig_decoder_gelu_probe exposes the two functions to tests/test_decoder_gpu.py.)
__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }

// a = fp16(gelu(fp16(bn(z))))   — the value torch's BatchNorm2d -> GELU chain hands to the next conv under autocast
__device__ __forceinline__ __half act_from_z(__half z, const BnCoef& c) {
    const __half y = f2h(fmaf(h2f(z), c.scale, c.shift));
    return f2h(gelu_f(h2f(y)));
}
// dz = fp16(scale*(dy - c1 - yhat*c2)),  dy = fp16(da * gelu'(y))   — GELU backward then cuDNN BN backward
__device__ __forceinline__ __half dz_from(__half da, __half z, const BnCoef& c) {
    const float zf = h2f(z);
    const __half y = f2h(fmaf(zf, c.scale, c.shift));
    const float dy = h2f(f2h(h2f(da) * gelu_grad_f(h2f(y))));
    const float yhat = (zf - c.mean) * c.rstd;
    return f2h(c.scale * (dy - c.c1 - yhat * c.c2));
}

struct BnSrc {            // where a consumer finds one layer's BatchNorm state
    const double* sums;   // [2][C]: sum z, sum z^2      (forward statistics)
    const double* bsums;  // [2][C]: sum dy, sum dy*yhat (backward statistics; IN_DZ)
    const float* gamma;
    const float* beta;
    float inv_n;          // 1 / (B*H*W)
    int stride;           // channels per row of sums / bsums (the producing layer's padded channel count)
    int valid;            // real channels: gamma / beta have this many entries; padded channels get all-zero coefficients
};

__device__ __forceinline__ BnCoef bn_coef(const BnSrc& s, int ch, bool with_bwd) {
    BnCoef c{};
    if (ch >= s.valid) return c;
    const int C = s.stride;
    const double mean = s.sums[ch] * (double)s.inv_n;
    double var = s.sums[C + ch] * (double)s.inv_n - mean * mean;   // biased variance, as BatchNorm normalises with
    var = var > 0.0 ? var : 0.0;
    c.mean = (float)mean;
    c.rstd = rsqrtf((float)var + kBnEps);
    c.scale = s.gamma[ch] * c.rstd;
    c.shift = s.beta[ch] - c.mean * c.scale;
    c.c1 = with_bwd ? (float)(s.bsums[ch] * (double)s.inv_n) : 0.f;
    c.c2 = with_bwd ? (float)(s.bsums[C + ch] * (double)s.inv_n) : 0.f;
    return c;
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const __half* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------------------
// Stage rows [r0-1, r0+R+1) x cols [-1, W+1) of image b (zero outside the image) into shared memory as
// tile[(R+2)*(W+2)][C + 8], applying the input transform.  src/src2: [B,H,W,C] fp16 (src2 = z for IN_DZ, src = da).
// ---------------------------------------------------------------------------------------------------------------
// SCH = channels per pixel in memory (<= C); tile channels [SCH, C) are zero.
// Two phases so the caller can put independent work between them: `issue` starts the global loads of up to kStagePF
// positions x chunks per thread (items base + j*blockDim + tid), `commit` applies the transform and writes the tile.
constexpr int kStagePF = 4;

// (dy, dy * yhat) of one output element for the backward BatchNorm statistics; out of line for the same reason as below
__device__ __noinline__ float2 bstat_terms(__half z, __half da, const BnCoef& co) {
    const float zf = h2f(z);
    const __half y = f2h(fmaf(zf, co.scale, co.shift));
    const float dy = h2f(f2h(h2f(da) * gelu_grad_f(h2f(y))));
    return make_float2(dy, dy * ((zf - co.mean) * co.rstd));
}

// 8 channels at a time, NOT inlined: the staging loops are unrolled for memory-level parallelism, and an inlined erff per
// element per unrolled slot turned the prologue into thousands of straight-line instructions (instruction-fetch stalls)
__device__ __noinline__ uint4 act8_from_z(uint4 v, const BnCoef* __restrict__ coef) {
    uint4 out;
    const __half* hv = reinterpret_cast<const __half*>(&v);
    __half* ho = reinterpret_cast<__half*>(&out);
#pragma unroll
    for (int k = 0; k < 8; ++k) ho[k] = act_from_z(hv[k], coef[k]);
    return out;
}
__device__ __noinline__ uint4 dz8_from(uint4 v, uint4 v2, const BnCoef* __restrict__ coef) {
    uint4 out;
    const __half* hv = reinterpret_cast<const __half*>(&v);
    const __half* hz = reinterpret_cast<const __half*>(&v2);
    __half* ho = reinterpret_cast<__half*>(&out);
#pragma unroll
    for (int k = 0; k < 8; ++k) ho[k] = dz_from(hv[k], hz[k], coef[k]);
    return out;
}

template <int C, int SCH, int MODE>
struct TileStager {
    static constexpr int STRIDE = C + 8, CHUNKS = C / 8;
    uint4 v[kStagePF], v2[kStagePF];
    uint32_t valid;

    __device__ __forceinline__ void issue(const __half* __restrict__ src, const __half* __restrict__ src2, int base, int b,
                                          int r0, int R, int H, int W) {
        const int TW = W + 2, total = (R + 2) * TW * CHUNKS;
        valid = 0u;
#pragma unroll
        for (int j = 0; j < kStagePF; ++j) {
            const int i = base + j * (int)blockDim.x + (int)threadIdx.x;
            v[j] = make_uint4(0u, 0u, 0u, 0u);
            v2[j] = make_uint4(0u, 0u, 0u, 0u);
            if (i >= total) continue;
            const int pos = i / CHUNKS, ck = i - pos * CHUNKS;
            const int tr = pos / TW, tc = pos - tr * TW;
            const int r = r0 - 1 + tr, w = tc - 1;
            if (r >= 0 && r < H && w >= 0 && w < W && ck * 8 < SCH) {
                const size_t off = (((size_t)b * H + r) * W + w) * SCH + ck * 8;
                v[j] = *reinterpret_cast<const uint4*>(src + off);
                if (MODE == IN_DZ) v2[j] = *reinterpret_cast<const uint4*>(src2 + off);
                valid |= 1u << j;
            }
        }
    }

    // element (pos, 8-channel chunk ck) lands at tile + pos * pos_stride + ck * ck_stride (halfs): row-major [pos][C + 8] by
    // default; the tcgen05 kernel passes (8, positions * 8) = the no-swizzle K-major core-matrix layout [ck][pos][8]
    __device__ __forceinline__ void commit(__half* tile, const BnCoef* __restrict__ coef, int base, int R, int W,
                                           int pos_stride = STRIDE, int ck_stride = 8) const {
        const int TW = W + 2, total = (R + 2) * TW * CHUNKS;
#pragma unroll
        for (int j = 0; j < kStagePF; ++j) {
            const int i = base + j * (int)blockDim.x + (int)threadIdx.x;
            if (i >= total) continue;
            const int pos = i / CHUNKS, ck = i - pos * CHUNKS;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);   // outside the image / padded channels: zero
            if (valid & (1u << j)) {
                if (MODE == IN_RAW) out = v[j];
                else if (MODE == IN_BNGELU) out = act8_from_z(v[j], coef + ck * 8);
                else out = dz8_from(v[j], v2[j], coef + ck * 8);
            }
            *reinterpret_cast<uint4*>(tile + (size_t)pos * pos_stride + (size_t)ck * ck_stride) = out;
        }
    }
};

struct ConvParams {
    const __half* src;     // [B,H,W,SCH]   activation (IN_RAW), z of the producing layer (IN_BNGELU) or da (IN_DZ)
    const __half* src2;    // z (IN_DZ)
    BnSrc bn;              // BatchNorm state of the input transform
    const __half* w;       // [COUT][9][CIN] fp16 (already rotated/transposed for data gradients)
    const __half* w_tc;    // the same 64x64 layer as [9 taps][64 x 64] in the UMMA canonical layout (k_dec_conv_tc), or null
    const float* bias;     // [COUT] or null
    __half* dst;           // [B,H,W,COUT]
    __half* act_out;       // optional [B,H,W,SCH]: the transformed input (a_{l-1} forward, dz_l backward) is written
                           // here once, by the CTA that owns the rows, for the weight-gradient kernel
    double* out_sums;      // [2][COUT] or null: += sum / sum of squares of the (fp16-rounded) outputs
    // data-gradient convs only: the outputs are da of the PREVIOUS layer; with z_out / bn_out of that layer the epilogue
    // also accumulates its backward statistics out_bsums[0][c] += sum dy, [1][c] += sum dy*yhat (dy = da*GELU'(BN(z))),
    // which is everything k_dec_bwd_stats would compute in a separate pass over da and z
    const __half* z_out;   // [B,H,W,COUT] or null
    BnSrc bn_out;
    double* out_bsums;     // [2][COUT] or null
    int B, H, W, R;        // R = image rows per CTA
    int cout_valid;        // outputs >= cout_valid are written as zero (channel padding)
    float* dimage;         // optional (first layer's data gradient): [B,H,W,3] fp32 = output channels 0..2 / std of the input
                           // normalisation, written by the epilogue (saves the conversion kernel behind the chain)
};

// ---------------------------------------------------------------------------------------------------------------
// 3x3 / pad 1 convolution as an implicit GEMM.  grid = (ceil(H/R), B), R*W <= 64 pixels per CTA preferred.
// A warp owns one n-tile (8 output channels) and up to 4 m-tiles (64 pixels): each weight fragment is loaded once per
// k-step straight from global memory (the 74 KB of a layer's weights stay in L1/L2; staging them through shared
// memory cost more than the math) and reused for every m-tile; A fragments come from the staged input tile.
// ---------------------------------------------------------------------------------------------------------------
// The body works on "items" = (image b, strip of R rows): item -> b = item / strips, r0 = (item % strips) * R.  The
// stand-alone kernel gives every CTA one item; the persistent forward kernel (k_dec_fwd_persist) lets a CTA loop over
// items item0, item0 + item_step, ... of a layer, loading the layer's weights and BatchNorm coefficients once.
template <int CIN, int SCH, int COUT, int MODE>
__device__ __forceinline__ void conv_cta(const ConvParams& p, unsigned char* smem_raw, int item0, int item_step, int n_items,
                                         int strips) {
    constexpr int STRIDE = CIN + 8, NT = COUT / 8, KS = CIN / 16, MB = kDecWarps >= 16 ? 2 : 4;
    if (item0 >= n_items) return;
    __half* tile = reinterpret_cast<__half*>(smem_raw);                      // [(R+2)*(W+2)][STRIDE]
    BnCoef* coef = reinterpret_cast<BnCoef*>(tile + (size_t)(p.R + 2) * (p.W + 2) * STRIDE);
    BnCoef* coef_out = coef + CIN;                                           // [COUT], data-gradient convs with out_bsums
    // the layer's weights [COUT][9*CIN], rows padded by 16 bytes (ldmatrix rows of an n-tile then sit in distinct banks)
    constexpr int WROW = 9 * CIN, WSTR = WROW + 8;
    __half* wsm = reinterpret_cast<__half*>(coef_out + COUT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const bool bstats = MODE == IN_DZ && p.out_bsums != nullptr;
    DEC_TRACE_BEGIN();
    // Asynchronous copy of the weights (up to 74 KB from L2), issued first: it completes while the BatchNorm
    // coefficients are derived and the input tile is staged and transformed.  Reading every weight fragment straight from
    // global memory inside the k-loop instead exposed an L2 round trip every few k-steps on a kernel whose math is ~1 us.
    {
        constexpr int CHUNKS = WROW / 8;                                     // 16-byte chunks per row
        for (int i = threadIdx.x; i < COUT * CHUNKS; i += blockDim.x) {
            const int row = i / CHUNKS, ck = i - row * CHUNKS;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(wsm + (size_t)row * WSTR + ck * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(p.w + (size_t)row * WROW + ck * 8) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

  for (int item = item0; item < n_items; item += item_step) {
    const int b = item / strips, r0 = (item - b * strips) * p.R, R = min(p.R, p.H - r0);
    const bool first = item == item0;
    // the first batch of input-tile loads is in flight while the BatchNorm coefficients (their own L2 round trip +
    // fp64 arithmetic) are derived: one exposed memory latency in the prologue instead of two
    TileStager<CIN, SCH, MODE> stager;
    DEC_TRACE(1);
    stager.issue(p.src, p.src2, 0, b, r0, R, p.H, p.W);
    DEC_TRACE(2);
    if (first && MODE != IN_RAW) {
        for (int ch = threadIdx.x; ch < SCH; ch += blockDim.x) coef[ch] = bn_coef(p.bn, ch, MODE == IN_DZ);
        if (bstats)
            for (int ch = threadIdx.x; ch < COUT; ch += blockDim.x) coef_out[ch] = bn_coef(p.bn_out, ch, false);
        __syncthreads();
    }
    DEC_TRACE(3);
    stager.commit(tile, coef, 0, R, p.W);
    DEC_TRACE(4);
    {
        const int total = (R + 2) * (p.W + 2) * (CIN / 8), step = kStagePF * (int)blockDim.x;
        for (int base = step; base < total; base += step) {
            stager.issue(p.src, p.src2, base, b, r0, R, p.H, p.W);
            stager.commit(tile, coef, base, R, p.W);
        }
    }
    if (first) asm volatile("cp.async.wait_group 0;" ::: "memory");
    DEC_TRACE(5);
    __syncthreads();
    DEC_TRACE(6);

    const int P = R * p.W, TW = p.W + 2, m_tiles = (P + 15) / 16, m_groups = (m_tiles + MB - 1) / MB;
    if (p.act_out) {  // materialise the centre rows of the transformed tile
        constexpr int CK = SCH / 8;
        for (int i = threadIdx.x; i < P * CK; i += blockDim.x) {
            const int pix = i / CK, ck = i - pix * CK, rr = pix / p.W, ww = pix - rr * p.W;
            *reinterpret_cast<uint4*>(p.act_out + (((size_t)b * p.H + r0 + rr) * p.W + ww) * SCH + ck * 8) =
                *reinterpret_cast<const uint4*>(tile + ((size_t)(rr + 1) * TW + (ww + 1)) * STRIDE + ck * 8);
        }
    }
    DEC_TRACE(7);
    for (int item = warp; item < m_groups * NT; item += kDecWarps) {
        const int mg = item / NT, nt = item - mg * NT;
        const __half* arow[MB];
#pragma unroll
        for (int m = 0; m < MB; ++m) {  // this lane's A row of m-tile mg*MB+m: pixel clamped to the strip
            const int pa = min((mg * MB + m) * 16 + (lane & 15), P - 1);
            const int ra = pa / p.W, wa = pa - ra * p.W;
            arow[m] = tile + ((size_t)(ra + 1) * TW + (wa + 1)) * STRIDE + (lane >> 4) * 8;
        }
        const int mcount = min(MB, m_tiles - mg * MB);
        // B fragments by ldmatrix: lane l addresses row (l & 7) of the n-tile, 8 halfs at k offset 8 * (l >> 3);
        // an x4 covers two consecutive k-steps (k is flat over taps x input channels: k-step kk = tap * KS + ks)
        const __half* wlane = wsm + (size_t)(nt * 8 + (lane & 7)) * WSTR + (lane >> 3) * 8;
        float c[MB][4];
#pragma unroll
        for (int m = 0; m < MB; ++m) { c[m][0] = c[m][1] = c[m][2] = c[m][3] = 0.f; }
        constexpr int KK = 9 * KS;
        if constexpr (KS % 2 == 0) {
            // a real loop over the 9 taps (only the KS k-steps of one tap are unrolled): the fully unrolled 36-k-step body
            // made the kernel ~75 KB of straight-line code that every SM executes exactly once, and a fifth of the warp
            // stall samples were instruction-fetch misses (stall_no_inst, profiles/r02_ncu_decoder.txt)
#pragma unroll 1
            for (int t = 0; t < 9; ++t) {
                const int off = ((t / 3 - 1) * TW + (t % 3 - 1)) * STRIDE;
#pragma unroll
                for (int kp = 0; kp < KS / 2; ++kp) {
                    uint32_t bf[4];
                    ldsm_x4(bf, wlane + (t * (KS / 2) + kp) * 32);
#pragma unroll
                    for (int half_ = 0; half_ < 2; ++half_) {
                        const int ks = 2 * kp + half_;
#pragma unroll
                        for (int m = 0; m < MB; ++m) {
                            if (m < mcount) {
                                uint32_t a[4];
                                ldsm_x4(a, arow[m] + off + ks * 16);
                                mma_16816(c[m], a, bf[2 * half_], bf[2 * half_ + 1]);
                            }
                        }
                    }
                }
            }
        } else {
#pragma unroll
        for (int kp = 0; kp < (KK + 1) / 2; ++kp) {
            uint32_t bf[4];
            if (2 * kp + 1 < KK) {
                ldsm_x4(bf, wlane + kp * 32);
            } else {   // odd tail (CIN == 16: 9 k-steps): lanes 16..31 re-address the first two matrices
                uint32_t b2[2];
                const uint32_t a = (uint32_t)__cvta_generic_to_shared(wlane - ((lane >> 4) * 16) + kp * 32);
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(b2[0]), "=r"(b2[1]) : "r"(a));
                bf[0] = b2[0]; bf[1] = b2[1]; bf[2] = 0u; bf[3] = 0u;
            }
#pragma unroll
            for (int half_ = 0; half_ < 2; ++half_) {
                const int kk = 2 * kp + half_;
                if (kk >= KK) break;
                const int t = kk / KS, ks = kk - t * KS;
                const int off = ((t / 3 - 1) * TW + (t % 3 - 1)) * STRIDE;
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    if (m < mcount) {
                        uint32_t a[4];
                        ldsm_x4(a, arow[m] + off + ks * 16);
                        mma_16816(c[m], a, bf[2 * half_], bf[2 * half_ + 1]);
                    }
                }
            }
        }
        }
        DEC_TRACE(8);
        // epilogue: rows g, g+8 of each m-tile; columns nt*8 + 2tig, +1
        const int col = nt * 8 + 2 * tig;
        const float bias0 = (p.bias && col < p.cout_valid) ? p.bias[col] : 0.f;
        const float bias1 = (p.bias && col + 1 < p.cout_valid) ? p.bias[col + 1] : 0.f;
        float l1[2] = {0.f, 0.f}, l2[2] = {0.f, 0.f};
#pragma unroll
        for (int m = 0; m < MB; ++m) {
            if (m >= mcount) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pp = (mg * MB + m) * 16 + h * 8 + g;
                if (pp < P) {
                    const int rr = pp / p.W, ww = pp - rr * p.W;
                    __half o0 = f2h(c[m][2 * h] + bias0), o1 = f2h(c[m][2 * h + 1] + bias1);
                    if (col >= p.cout_valid) o0 = f2h(0.f);
                    if (col + 1 >= p.cout_valid) o1 = f2h(0.f);
                    const size_t oidx = (((size_t)b * p.H + r0 + rr) * p.W + ww) * COUT + col;
                    *reinterpret_cast<__half2*>(p.dst + oidx) = __halves2half2(o0, o1);
                    if (MODE == IN_DZ && p.dimage != nullptr && col < 3) {   // d(image) = d(x0) / std (normalize_img backward: hidden_models.py)
                        const float stdv[4] = {0.229f, 0.224f, 0.225f, 1.0f};
                        float* di = p.dimage + ((((size_t)b * p.H + r0 + rr) * p.W + ww)) * 3;
                        di[col] = __fdiv_rn(h2f(o0), stdv[col]);
                        if (col + 1 < 3) di[col + 1] = __fdiv_rn(h2f(o1), stdv[col + 1]);
                    }
                    if (bstats) {   // backward statistics of the layer whose da this is (same arithmetic as k_dec_bwd_stats)
                        const __half2 zz = *reinterpret_cast<const __half2*>(p.z_out + oidx);
                        const __half zh[2] = {__low2half(zz), __high2half(zz)}, dh[2] = {o0, o1};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float2 t = bstat_terms(zh[e], dh[e], coef_out[col + e]);
                            l1[e] += t.x;
                            l2[e] += t.y;
                        }
                    } else {
                        const float f0 = h2f(o0), f1 = h2f(o1);
                        l1[0] += f0; l1[1] += f1; l2[0] += f0 * f0; l2[1] += f1 * f1;
                    }
                }
            }
        }
        double* const sums_dst = bstats ? p.out_bsums : p.out_sums;
        if (sums_dst) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a1 = l1[e], a2 = l2[e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a1 += __shfl_xor_sync(NSIG_FULL_MASK, a1, o); a2 += __shfl_xor_sync(NSIG_FULL_MASK, a2, o); }
                if (g == 0 && col + e < p.cout_valid) {
                    atomicAdd(sums_dst + col + e, (double)a1);
                    atomicAdd(sums_dst + COUT + col + e, (double)a2);
                }
            }
        }
    }
    DEC_TRACE(9);
    __syncthreads();   // the tile is restaged for the CTA's next item
    DEC_TRACE(10);
  }
}

template <int CIN, int SCH, int COUT, int MODE>
__global__ void __launch_bounds__(kDecThreads)
k_dec_conv(const ConvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    conv_cta<CIN, SCH, COUT, MODE>(p, smem_raw, (int)(blockIdx.y * gridDim.x + blockIdx.x), 1 << 30,
                                   (int)(gridDim.x * gridDim.y), (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// The 64 -> 64 convolution on tcgen05 + TMEM (default; NSIG_DEC_TC=0 selects k_dec_conv).  Same contract as k_dec_conv<64,64,64,MODE> (staging, transforms,
// act_out, outputs, statistics); what changes is the contraction.  The phase trace of the mma.sync kernel
// (profiles/r02_decoder_phase_trace.txt) shows its k-loop as the largest phase (3.7 us forward / 4.9 us data gradient of
// 8.3 / 13.4 us): 864 m16n8k16 MMAs and 1008 ldmatrix.x4 per CTA, every warp re-reading the whole activation tile from
// shared memory for its 8 output channels.  Here the tile is written ONCE in the UMMA K-major core-matrix layout
// [8-channel chunk][position][8 halfs] (positions = the strip with its halo, row-major over the PADDED width W + 2) and the
// tensor core reads it: output row m = r * (W + 2) + c needs, for tap (dy, dx), input position m + (1 + dy) * (W + 2) + 1 + dx -
// a shift of the operand's START ADDRESS by 16 bytes per position (rows are 16 bytes apart throughout this layout, so any
// shift keeps every 8-row core matrix contiguous).  36 MMAs (M = 128, N = 64, K = 16) issued by one thread accumulate all
// 9 taps in 64 TMEM columns; rows of the padding columns (c >= W) and past the strip are computed and dropped.
// Weights: [tap][64 x 64] in the canonical layout of field_tc.cu (leading byte offset 128, stride byte offset 1024).
// ---------------------------------------------------------------------------------------------------------------
namespace tcv {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor (no swizzle, version 1): start address, leading / stride byte offsets in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();   // a lost completion must fail loudly, never hang the GPU
    }
}
#define NSIG_DEC_TMEM_LD16(taddr, v)                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"  \
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),         \
                   "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])    \
                 : "r"(taddr))
constexpr uint32_t kCols = 64;
constexpr int kWBytes = 9 * 64 * 64 * 2;
// positions of the A operand that must be addressable: the largest tap offset 2 * (W + 2) + 2 plus 128 rows
// (= 1 mod 8: the 8 channel chunks of a position then start 16 bytes apart modulo 128 - conflict-free tile stores)
__host__ __device__ inline int npos_alloc(int W) { return (2 * (W + 2) + 2 + 128 + 7) / 8 * 8 + 1; }
__host__ __device__ inline size_t smem_bytes(int R, int W) {
    return (size_t)kWBytes + (((size_t)8 * npos_alloc(W) * 16 + 127) / 128 * 128) + (size_t)R * W * 64 * 2 + 2 * 64 * sizeof(BnCoef) + 4 * 64 * 2 * sizeof(float) + 64;
}
}  // namespace tcv

template <int MODE>
__global__ void __launch_bounds__(kDecThreads)
k_dec_conv_tc(const ConvParams p) {
    constexpr int C = 64;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, r0 = blockIdx.x * p.R, R = min(p.R, p.H - r0), TW = p.W + 2, P = R * p.W, Mrows = R * TW;
    const int NposA = tcv::npos_alloc(p.W);
    __half* wsm = reinterpret_cast<__half*>(smem_raw);
    __half* atile = reinterpret_cast<__half*>(smem_raw + tcv::kWBytes);                  // [8][NposA][8]
    __half* otile = atile + ((size_t)8 * NposA * 8 + 63) / 64 * 64;                      // [P][64] fp16 outputs of the strip
    BnCoef* coef = reinterpret_cast<BnCoef*>(otile + (size_t)p.R * p.W * C);
    BnCoef* coef_out = coef + C;
    float* red = reinterpret_cast<float*>(coef_out + C);                                 // [4][64][2] statistics partials
    uint64_t* mbar_store = reinterpret_cast<uint64_t*>(red + 4 * 64 * 2);      // [0] accumulator complete, [1] weights arrived
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar_store + 2);
    const uint32_t mbar = tcv::smem_u32(mbar_store), mbar_w = tcv::smem_u32(mbar_store + 1);
    const bool bstats = MODE == IN_DZ && p.out_bsums != nullptr;
    DEC_TRACE_BEGIN();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel of the chain may start its prologue

    // weights: 72 KB already in the operand layout (k_dec_prep_weights) -> 9 bulk asynchronous copies issued by one thread
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_w));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_w), "r"((uint32_t)tcv::kWBytes) : "memory");
#pragma unroll
        for (int t = 0; t < 9; ++t)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(tcv::smem_u32(wsm) + t * 8192), "l"(p.w_tc + (size_t)t * 4096), "r"(8192u), "r"(mbar_w) : "memory");
    }
    DEC_TRACE(11);
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tcv::smem_u32(tmem_slot)), "n"(tcv::kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }

    DEC_TRACE(1);
    // Programmatic dependent launch: everything above (barriers, the 72 KB weight fetch, the TMEM allocation) touches nothing
    // the preceding kernel of the chain writes, so this grid may start while that one is still running; from here on its
    // outputs (activations, statistics) are needed.  No-op when the kernel was launched without the attribute.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    TileStager<C, C, MODE> stager;
    stager.issue(p.src, p.src2, 0, b, r0, R, p.H, p.W);
    DEC_TRACE(2);
    for (int ch = tid; ch < C; ch += blockDim.x) coef[ch] = bn_coef(p.bn, ch, MODE == IN_DZ);
    if (bstats)
        for (int ch = tid; ch < C; ch += blockDim.x) coef_out[ch] = bn_coef(p.bn_out, ch, false);
    __syncthreads();
    DEC_TRACE(3);
    stager.commit(atile, coef, 0, R, p.W, 8, NposA * 8);
    {
        const int total = (R + 2) * TW * (C / 8), step = kStagePF * (int)blockDim.x;
        for (int base = step; base < total; base += step) {
            stager.issue(p.src, p.src2, base, b, r0, R, p.H, p.W);
            stager.commit(atile, coef, base, R, p.W, 8, NposA * 8);
        }
    }
    DEC_TRACE(4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes (the tile) -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    DEC_TRACE(5);

    if (tid == 0) {
        // instruction descriptor: D = fp32 (bit 4), A = B = fp16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t chunk_bytes = (uint32_t)NposA * 16;
        // descriptors differ in their start-address field only (16-byte units, no carry out of its 14 bits)
        const uint64_t da0 = tcv::make_desc(tcv::smem_u32(atile), chunk_bytes, 128);
        const uint64_t db0 = tcv::make_desc(tcv::smem_u32(wsm), 128, 1024);
        const uint32_t kstep = (2 * chunk_bytes) >> 4, trow = (uint32_t)TW;
        tcv::mbar_wait(mbar_w, 0);     // weights in shared memory (async proxy writes: no proxy fence needed)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                tcv::mma_f16_ss(tmem_d, da0 + (uint64_t)((t / 3) * trow + (t % 3) + ks * kstep),
                                db0 + (uint64_t)(t * 512 + ks * 16), idesc, (t | ks) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
    }
    DEC_TRACE(6);
    if (p.act_out) {  // materialise the centre rows of the transformed tile (while the tensor core works)
        for (int i = tid; i < P * 8; i += blockDim.x) {
            const int pix = i >> 3, ck = i & 7, rr = pix / p.W, ww = pix - rr * p.W;
            *reinterpret_cast<uint4*>(p.act_out + (((size_t)b * p.H + r0 + rr) * p.W + ww) * C + ck * 8) =
                *reinterpret_cast<const uint4*>(atile + (size_t)ck * NposA * 8 + (size_t)((rr + 1) * TW + (ww + 1)) * 8);
        }
    }
    DEC_TRACE(7);
    tcv::mbar_wait(mbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DEC_TRACE(8);

    // accumulators -> fp16 outputs in shared memory: warp w reads TMEM lanes 32 * (w % 4).., columns 32 * (w / 4)..
    {
        const int q = warp & 3, hf = warp >> 2;
        if (32 * q < Mrows) {
            float v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * hf);
            NSIG_DEC_TMEM_LD16(taddr, v);
            NSIG_DEC_TMEM_LD16(taddr + 16, (&v[16]));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int m = 32 * q + lane, rr = m / TW, cc = m - rr * TW;
            if (m < Mrows && cc < p.W) {
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = 32 * hf + 2 * j;
                    const float b0 = p.bias ? p.bias[col] : 0.f, b1 = p.bias ? p.bias[col + 1] : 0.f;
                    const __half2 h = __halves2half2(f2h(v[2 * j] + b0), f2h(v[2 * j + 1] + b1));
                    w[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                uint4* dst = reinterpret_cast<uint4*>(otile + (size_t)(rr * p.W + cc) * C + 32 * hf);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(tcv::kCols));
    DEC_TRACE(9);

    // the strip's outputs are contiguous in the NHWC tensor
    {
        uint4* gdst = reinterpret_cast<uint4*>(p.dst + ((size_t)b * p.H + r0) * p.W * C);
        const uint4* osrc = reinterpret_cast<const uint4*>(otile);
        for (int i = tid; i < P * 8; i += blockDim.x) gdst[i] = osrc[i];
    }
    // statistics of the (fp16-rounded) outputs: thread (channel, part) walks every 4th pixel
    double* const sums_dst = bstats ? p.out_bsums : p.out_sums;
    if (sums_dst) {
        const int ch = tid & 63, part = tid >> 6;
        float s1 = 0.f, s2 = 0.f;
        if (bstats) {
            const __half* zsrc = p.z_out + ((size_t)b * p.H + r0) * p.W * C + ch;
            constexpr int PF = 8, STEP = kDecThreads / 64;   // z of 8 pixels requested together (one L2 round trip, not eight)
            for (int pix0 = part; pix0 < P; pix0 += PF * STEP) {
                __half zv[PF];
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    const int pix = pix0 + j * STEP;
                    zv[j] = pix < P ? zsrc[(size_t)pix * C] : f2h(0.f);
                }
#pragma unroll
                for (int j = 0; j < PF; ++j) {
                    const int pix = pix0 + j * STEP;
                    if (pix < P) {
                        const float2 t = bstat_terms(zv[j], otile[(size_t)pix * C + ch], coef_out[ch]);
                        s1 += t.x;
                        s2 += t.y;
                    }
                }
            }
        } else {
            for (int pix = part; pix < P; pix += kDecThreads / 64) {
                const float f = h2f(otile[(size_t)pix * C + ch]);
                s1 += f;
                s2 += f * f;
            }
        }
        red[(part * 64 + ch) * 2] = s1;
        red[(part * 64 + ch) * 2 + 1] = s2;
        __syncthreads();
        if (tid < 128) {
            const int c2 = tid & 63, which = tid >> 6;
            float acc = 0.f;
#pragma unroll
            for (int pt = 0; pt < kDecThreads / 64; ++pt) acc += red[(pt * 64 + c2) * 2 + which];
            atomicAdd(sums_dst + which * C + c2, (double)acc);
        }
    }
    DEC_TRACE(10);
}

// ---------------------------------------------------------------------------------------------------------------
// backward statistics of one layer: bsums[0][c] += sum dy, bsums[1][c] += sum dy*yhat   (dy = da * GELU'(BN(z)))
// grid = any, grid-stride over pixels; C channels (64 or 8)
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kDecThreads)
k_dec_bwd_stats(const __half* __restrict__ da, const __half* __restrict__ z, BnSrc bn, double* __restrict__ bsums, int n_pix) {
    __shared__ BnCoef coef[C];
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) coef[ch] = bn_coef(bn, ch, false);
    __syncthreads();
    constexpr int CK = C / 8;                       // 16-byte chunks per pixel
    const int slot = threadIdx.x % CK;               // this thread always handles the same 8 channels
    float a1[8], a2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a1[k] = 0.f; a2[k] = 0.f; }
    const int per_blk = kDecThreads / CK;
    for (int pix = blockIdx.x * per_blk + threadIdx.x / CK; pix < n_pix; pix += gridDim.x * per_blk) {
        const uint4 v = *reinterpret_cast<const uint4*>(da + (size_t)pix * C + slot * 8);
        const uint4 v2 = *reinterpret_cast<const uint4*>(z + (size_t)pix * C + slot * 8);
        const __half* hd = reinterpret_cast<const __half*>(&v);
        const __half* hz = reinterpret_cast<const __half*>(&v2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const BnCoef& c = coef[slot * 8 + k];
            const float zf = h2f(hz[k]);
            const __half y = f2h(fmaf(zf, c.scale, c.shift));
            const float dy = h2f(f2h(h2f(hd[k]) * gelu_grad_f(h2f(y))));
            a1[k] += dy;
            a2[k] += dy * ((zf - c.mean) * c.rstd);
        }
    }
    // lanes l, l+CK, l+2CK, ... of a warp hold the same channels: fold them before touching shared memory
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int o = 16; o >= CK; o >>= 1) { a1[k] += __shfl_xor_sync(NSIG_FULL_MASK, a1[k], o); a2[k] += __shfl_xor_sync(NSIG_FULL_MASK, a2[k], o); }
    }
    // per-warp partials, then a fixed-order sum over the warps (no shared-memory float atomics: results are reproducible)
    __shared__ float part[2][kDecWarps][C];
    if ((threadIdx.x & 31) < CK) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { part[0][threadIdx.x >> 5][slot * 8 + k] = a1[k]; part[1][threadIdx.x >> 5][slot * 8 + k] = a2[k]; }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < kDecWarps; ++w) { s1 += part[0][w][ch]; s2 += part[1][w][ch]; }
        atomicAdd(bsums + ch, (double)s1);
        atomicAdd(bsums + C + ch, (double)s2);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient of one conv layer: dW[co][ci][tap] += sum_pixels dz[p][co] * a[p + tap][ci];  db[co] += sum dz.
// grid = (9 taps, G image groups).  A CTA owns ONE tap and keeps its [COUT x CIN] partial in registers while it walks
// over its images strip by strip (a and dz are the tensors materialised by the forward / data-gradient convs, so
// staging is a plain shifted copy), then adds it to dW with one atomic per element: 9*G*COUT*CIN atomics per layer
// instead of one full dW per strip.  Both MMA operands are transposed 8x8 blocks (ldmatrix.trans): the contraction
// runs over pixels, which is the row index of both tiles.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
    const __half* a;       // input activation of the layer [B,H,W,CIN] (x0 or the materialised a_{l-1})
    const __half* dz;      // gradient wrt the conv output  [B,H,W,DCH] (materialised by the data-gradient conv)
    float* dW;             // [cout_real][cin_real][3][3]
    float* db;             // [cout_real]
    // two-phase reduction (null: one atomicAdd per element and CTA instead): every CTA stores its [COUT x CIN] partial to
    // partial[(group * 9 + tap)] with plain coalesced stores and takes a ticket of its tap; the CTA that draws the last
    // ticket sums the groups IN GROUP ORDER and adds the result to dW - deterministic, and 9*G*COUT*CIN scattered L2
    // atomics (stride 36 B) become 8-byte stores plus one read-modify-write of dW
    float* partial;        // [G][9][COUT][CIN] fp32
    unsigned int* tickets; // [9] (+1 for the bias), zero before the launch
    int B, H, W, R, cin_real, cout_real;
};

constexpr int kWgradBatchMax = 8;
struct WgradBatch { WgradParams p[kWgradBatchMax]; };   // blockIdx.z selects the layer (layers of one shape in ONE launch)

template <int CIN, int COUT, int DCH>
__global__ void __launch_bounds__(kDecThreads)
k_dec_wgrad(const WgradBatch q) {
    const WgradParams& p = q.p[blockIdx.z];
    constexpr int ASTR = CIN + 8, DSTR = COUT + 8, MT = COUT / 16, NT = CIN / 8, ITEMS = MT * NT;
    constexpr int PER_WARP = (ITEMS + kDecWarps - 1) / kDecWarps;          // 4 (64x64), 1 (64x16 or 16x64)
    constexpr int NT_PER = PER_WARP;                                        // a warp's items share one m-tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Pmax = (p.R * p.W + 15) / 16 * 16;
    __half* atile = reinterpret_cast<__half*>(smem_raw);   // 2 x { [Pmax][ASTR] a shifted by this CTA's tap, [Pmax][DSTR] dz }
    const int t = blockIdx.x, dy = t / 3 - 1, dx = t % 3 - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    const int item0 = warp * PER_WARP;                                      // items [item0, item0 + PER_WARP): (mt, nt..)
    const int mt = item0 / NT, nt0 = item0 - mt * NT;
    const bool active = item0 < ITEMS;
    float c[PER_WARP][4];
#pragma unroll
    for (int j = 0; j < PER_WARP; ++j) { c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f; }
    float dbias = 0.f;   // centre-tap CTAs: thread (co = tid % 64, part = tid / 64) sums every 4th pixel of column co of dz
    const int bco = threadIdx.x & 63, bpart = threadIdx.x >> 6;

    // Work items = (image of this group, strip of R rows), double-buffered: item k+1's tiles travel global -> shared memory
    // as 16-byte asynchronous copies (zero-filled outside the image / past the real channels) while item k is contracted.
    // The register-staged version paid a dependent L2 round trip per copy iteration (~9 per tile) and nothing overlapped:
    // 25 us per layer for 36 MMAs per warp.
    const int spi = (p.H + p.R - 1) / p.R;                                  // strips per image
    const int n_img = ((int)p.B - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
    const int n_items = n_img * spi;
    const size_t buf_halfs = (size_t)Pmax * (ASTR + DSTR);
    // pix / W by multiplication (exact for pix, W < 2^16): the two runtime divisions per 16-byte copy made ISSUING one item's
    // copies take 2.1 us (phase trace of this kernel)
    const uint32_t w_magic = 0xFFFFFFFFu / (uint32_t)p.W + 1u;
    auto stage = [&](int k, int which) {
        const int b = blockIdx.y + (k / spi) * gridDim.y, r0 = (k % spi) * p.R;
        const int R = min(p.R, p.H - r0), P = R * p.W, Ppad = (P + 15) / 16 * 16;
        __half* at = atile + which * buf_halfs;
        __half* dt = at + (size_t)Pmax * ASTR;
        const __half* a_img = p.a + (size_t)b * p.H * p.W * CIN;
        const __half* d_img = p.dz + ((size_t)b * p.H + r0) * p.W * DCH;   // the strip's dz rows are contiguous
        for (int i = threadIdx.x; i < Ppad * (CIN / 8); i += blockDim.x) {
            const int pix = i / (CIN / 8), ck = i - pix * (CIN / 8);
            const __half* src = p.a;
            uint32_t bytes = 0;
            if (pix < P) {
                const int prow = (int)__umulhi((uint32_t)pix, w_magic);
                const int rr = r0 + prow + dy, ww = pix - prow * p.W + dx;
                if (rr >= 0 && rr < p.H && ww >= 0 && ww < p.W) {
                    src = a_img + (size_t)(rr * p.W + ww) * CIN + ck * 8;
                    bytes = 16;
                }
            }
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(at + (size_t)pix * ASTR + ck * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        for (int i = threadIdx.x; i < Ppad * (COUT / 8); i += blockDim.x) {
            const int pix = i / (COUT / 8), ck = i - pix * (COUT / 8);
            const __half* src = p.dz;
            uint32_t bytes = 0;
            if (pix < P && ck * 8 < DCH) {
                src = d_img + (size_t)pix * DCH + ck * 8;
                bytes = 16;
            }
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dt + (size_t)pix * DSTR + ck * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    WG_TRACE_BEGIN();
    if (n_items > 0) stage(0, 0);
    WG_TRACE(1);
    for (int k = 0; k < n_items; ++k) {
        if (k + 1 < n_items) {
            stage(k + 1, (k + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        {
            const int r0 = (k % spi) * p.R;
            const int R = min(p.R, p.H - r0), P = R * p.W, Ppad = (P + 15) / 16 * 16;
            const __half* at = atile + (k & 1) * buf_halfs;
            const __half* dtile = at + (size_t)Pmax * ASTR;
            if (t == 4 && bco < p.cout_real) {  // bias gradient: column sums of dz (centre-tap CTAs only)
                for (int pix = bpart; pix < P; pix += kDecThreads / 64) dbias += h2f(dtile[(size_t)pix * DSTR + bco]);
            }
            if (active) {
                for (int k0 = 0; k0 < Ppad; k0 += 16) {
                    uint32_t a[4];
                    ldsm_x4_trans(a, dtile + (size_t)(k0 + (lane & 7) + ((lane >> 4) << 3)) * DSTR + mt * 16 + (((lane >> 3) & 1) << 3));
#pragma unroll
                    for (int j = 0; j < NT_PER; ++j) {
                        uint32_t bb[2];
                        ldsm_x2_trans(bb, at + (size_t)(k0 + (lane & 15)) * ASTR + (nt0 + j) * 8);
                        mma_16816(c[j], a, bb[0], bb[1]);
                    }
                }
            }
        }
        __syncthreads();   // everybody is done with buffer k&1 before item k+2 is copied into it
        WG_TRACE(2 + (k < 2 ? k : 1));
    }
    if (p.partial == nullptr) {
        if (t == 4 && bco < p.cout_real) atomicAdd(p.db + bco, dbias);
        if (active) {
#pragma unroll
            for (int j = 0; j < NT_PER; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int co = mt * 16 + h * 8 + g;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int ci = (nt0 + j) * 8 + 2 * tig + e;
                        if (co < p.cout_real && ci < p.cin_real) atomicAdd(p.dW + ((size_t)co * p.cin_real + ci) * 9 + t, c[j][2 * h + e]);
                    }
                }
        }
        return;
    }
    // slot (group, tap) of the partials; slot (group, 9) holds the group's bias partial (written by its centre-tap CTA)
    float* mine = p.partial + ((size_t)blockIdx.y * 10 + t) * (COUT * CIN);
    __shared__ float s_db[kDecThreads];
    if (t == 4) {
        s_db[threadIdx.x] = dbias;
        __syncthreads();
        if (threadIdx.x < 64 && threadIdx.x < COUT) {
            float acc = 0.f;
            for (int part = 0; part < kDecThreads / 64; ++part) acc += s_db[part * 64 + threadIdx.x];
            p.partial[((size_t)blockIdx.y * 10 + 9) * (COUT * CIN) + threadIdx.x] = acc;
        }
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < NT_PER; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = mt * 16 + h * 8 + g, ci = (nt0 + j) * 8 + 2 * tig;
                *reinterpret_cast<float2*>(mine + co * CIN + ci) = make_float2(c[j][2 * h], c[j][2 * h + 1]);
            }
    }
    if (p.tickets == nullptr) return;   // the groups are summed by k_dec_wgrad_reduce (one launch for all layers, after the join)
    __shared__ unsigned int s_ticket;
    WG_TRACE(4);
    __threadfence();
    __syncthreads();
    WG_TRACE(5);
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.tickets + t, 1u);
    __syncthreads();
    WG_TRACE(6);
    if (s_ticket != gridDim.y - 1) return;
    WG_TRACE_LAST(7);
    __threadfence();   // acquire: the other groups' partials were fenced before their tickets
    const float* base = p.partial + (size_t)t * (COUT * CIN);
    // thread owns NQ groups of 4 consecutive elements (16-byte loads); the groups are summed in group order (deterministic)
    constexpr int NQ = COUT * CIN / (4 * kDecThreads);            // 4 (64x64), 1 (64x16 / 16x64)
    constexpr int GU = NQ >= 4 ? 4 : 8;                            // groups whose loads are in flight together
    float4 acc[NQ];
#pragma unroll
    for (int e = 0; e < NQ; ++e) acc[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned int grp = 0; grp < gridDim.y; grp += GU) {
        float4 v[GU][NQ];
#pragma unroll
        for (int u = 0; u < GU; ++u)
#pragma unroll
            for (int e = 0; e < NQ; ++e)
                v[u][e] = (grp + u < gridDim.y)
                              ? __ldcg(reinterpret_cast<const float4*>(base + (size_t)(grp + u) * 10 * (COUT * CIN)) + e * kDecThreads + threadIdx.x)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < GU; ++u)
            if (grp + u < gridDim.y) {
#pragma unroll
                for (int e = 0; e < NQ; ++e) {
                    acc[e].x += v[u][e].x; acc[e].y += v[u][e].y; acc[e].z += v[u][e].z; acc[e].w += v[u][e].w;
                }
            }
    }
    // dW += acc, this CTA being the sole writer of tap t: all the old values are requested first - as "load, add, store"
    // per element the compiler must keep the (possibly aliasing) accesses in order, i.e. dependent L2 round trips
    WG_TRACE_LAST(8);
    float oldw[NQ][4];
#pragma unroll
    for (int e = 0; e < NQ; ++e)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = (e * kDecThreads + (int)threadIdx.x) * 4 + c, co = i / CIN, ci = i - co * CIN;
            oldw[e][c] = (co < p.cout_real && ci < p.cin_real) ? __ldcg(p.dW + ((size_t)co * p.cin_real + ci) * 9 + t) : 0.f;
        }
#pragma unroll
    for (int e = 0; e < NQ; ++e) {
        const float a4[4] = {acc[e].x, acc[e].y, acc[e].z, acc[e].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = (e * kDecThreads + (int)threadIdx.x) * 4 + c, co = i / CIN, ci = i - co * CIN;
            if (co < p.cout_real && ci < p.cin_real) p.dW[((size_t)co * p.cin_real + ci) * 9 + t] = oldw[e][c] + a4[c];
        }
    }
    WG_TRACE_LAST(9);
    if (t == 4 && (int)threadIdx.x < p.cout_real) {
        float acc = 0.f;
        for (unsigned int grp = 0; grp < gridDim.y; ++grp)
            acc += __ldcg(p.partial + ((size_t)grp * 10 + 9) * (COUT * CIN) + threadIdx.x);
        p.db[threadIdx.x] += acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient of the 64 -> 64 layers with ALL NINE TAPS in one CTA (k_dec_wgrad loads a and dz once PER TAP: 9 x the
// traffic, 9 x the CTAs; behind the data-gradient chain, where all layers' weight gradients run at once, the batched launch of
// 1008 such CTAs took 40-70 us).  grid = (G image groups, layers); 16 warps: warp (mt = w / 4, pair = w % 4) keeps the
// [16 cout x 16 cin] block of all 9 taps in registers (72 accumulators) while the CTA walks its images: the image's a tile is
// staged WITH its halo (padded width W + 2, zero outside) and dz next to it, double-buffered by 16-byte asynchronous copies;
// per 16-pixel k-step the dz fragment is loaded once and reused for the 9 taps, whose a fragments are the same ldmatrix rows
// shifted by dy * (W + 2) + dx.  Partials go to the same [group][tap] slots k_dec_wgrad fills, in the same order of
// accumulation over pixels and images (same bits); k_dec_wgrad_reduce sums the groups.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWgAllThreads = 512;
__global__ void __launch_bounds__(kWgAllThreads, 1)
k_dec_wgrad_all(const WgradBatch q) {
    constexpr int C = 64, ASTR = C + 8, DSTR = C + 8;
    const WgradParams& p = q.p[blockIdx.y];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int TW = p.W + 2, NPAD = (p.H + 2) * TW, P = p.H * p.W, Ppad = (P + 15) / 16 * 16;
    const size_t buf_halfs = (size_t)NPAD * ASTR + (size_t)Ppad * DSTR;
    __half* buf0 = reinterpret_cast<__half*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
    const int mt = warp >> 2, nt0 = (warp & 3) * 2;
    const uint32_t w_magic = 0xFFFFFFFFu / (uint32_t)p.W + 1u, tw_magic = 0xFFFFFFFFu / (uint32_t)TW + 1u;
    float c[9][2][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j) { c[t][j][0] = c[t][j][1] = c[t][j][2] = c[t][j][3] = 0.f; }
    float dbias = 0.f;
    const int bco = tid & 63, bpart = tid >> 6;
    const int n_img = ((int)p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    auto stage = [&](int k, int which) {
        const int b = blockIdx.x + k * gridDim.x;
        __half* at = buf0 + which * buf_halfs;
        __half* dt = at + (size_t)NPAD * ASTR;
        const __half* a_img = p.a + (size_t)b * P * C;
        const __half* d_img = p.dz + (size_t)b * P * C;
        for (int i = tid; i < NPAD * 8; i += kWgAllThreads) {
            const int pos = i >> 3, ck = i & 7;
            const int tr = (int)__umulhi((uint32_t)pos, tw_magic), tc = pos - tr * TW;
            const int rr = tr - 1, ww = tc - 1;
            const __half* src = p.a;
            uint32_t bytes = 0;
            if (rr >= 0 && rr < p.H && ww >= 0 && ww < p.W) { src = a_img + (size_t)(rr * p.W + ww) * C + ck * 8; bytes = 16; }
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(at + (size_t)pos * ASTR + ck * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        for (int i = tid; i < Ppad * 8; i += kWgAllThreads) {
            const int pix = i >> 3, ck = i & 7;
            const __half* src = p.dz;
            uint32_t bytes = 0;
            if (pix < P) { src = d_img + (size_t)pix * C + ck * 8; bytes = 16; }
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dt + (size_t)pix * DSTR + ck * 8);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_img > 0) stage(0, 0);
    for (int k = 0; k < n_img; ++k) {
        if (k + 1 < n_img) {
            stage(k + 1, (k + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const __half* at = buf0 + (k & 1) * buf_halfs;
        const __half* dt = at + (size_t)NPAD * ASTR;
        if (bco < p.cout_real)
            for (int pix = bpart; pix < P; pix += kWgAllThreads / 64) dbias += h2f(dt[(size_t)pix * DSTR + bco]);
        for (int k0 = 0; k0 < Ppad; k0 += 16) {
            uint32_t a[4];
            ldsm_x4_trans(a, dt + (size_t)(k0 + (lane & 7) + ((lane >> 4) << 3)) * DSTR + mt * 16 + (((lane >> 3) & 1) << 3));
            // this lane's B row of the k-step: pixel k0 + (lane & 15) at the centre tap, in padded-tile coordinates
            const int pix = min(k0 + (lane & 15), P - 1);
            const int prow = (int)__umulhi((uint32_t)pix, w_magic);
            const __half* brow = at + (size_t)((prow + 1) * TW + (pix - prow * p.W) + 1) * ASTR + nt0 * 8;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int off = ((t / 3 - 1) * TW + (t % 3 - 1)) * ASTR;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t bb[2];
                    ldsm_x2_trans(bb, brow + off + j * 8);
                    mma_16816(c[t][j], a, bb[0], bb[1]);
                }
            }
        }
        __syncthreads();   // everybody is done with buffer k&1 before image k+2 is copied into it
    }
    // partials: slot (group, tap) as [cout][cin]; slot (group, 9) = the group's bias partial
    __shared__ float s_db[kWgAllThreads];
    s_db[tid] = dbias;
    __syncthreads();
    if (tid < 64) {
        float acc = 0.f;
        for (int part = 0; part < kWgAllThreads / 64; ++part) acc += s_db[part * 64 + tid];
        p.partial[((size_t)blockIdx.x * 10 + 9) * (C * C) + tid] = acc;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        float* mine = p.partial + ((size_t)blockIdx.x * 10 + t) * (C * C);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int co = mt * 16 + h * 8 + g, ci = (nt0 + j) * 8 + 2 * tig;
                *reinterpret_cast<float2*>(mine + co * C + ci) = make_float2(c[t][j][2 * h], c[t][j][2 * h + 1]);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Second phase of the weight gradients for ALL layers in one launch: dW[co][ci][tap] += sum over the image groups (in group
// order: deterministic, the same additions the ticket scheme's last CTA performed) of the partials k_dec_wgrad stored;
// db likewise.  The ticket scheme ended every weight-gradient kernel with 9 CTAs summing 16 groups x 16 KB and writing dW with
// stride-36-byte read-modify-writes while the other 135 CTAs had left: 7 of the kernel's 12.5 us (phase trace), nine times
// per backward, next to the data-gradient chain whose stragglers it slowed.  grid = (17, layers).
// ---------------------------------------------------------------------------------------------------------------
struct WgradReduceParams {
    const float* partial[17];
    float* dW[17];
    float* db[17];
    int n_elem[17];                 // COUT * CIN of the layer's (padded) partial
    int cin_pad[17], cin_real[17], cout_real[17];
    int groups;
};
__global__ void __launch_bounds__(256)
k_dec_wgrad_reduce(const WgradReduceParams q) {
    // block (x, layer): x < 16 = 256 consecutive elements (co, ci) of the [COUT x CIN] partial, ALL nine taps - a thread's nine
    // results are 36 contiguous bytes of dW[co][ci][3][3] and a warp's a contiguous 1152: coalesced read-modify-write (one block
    // per tap wrote dW with stride-36-byte accesses: 21 us under ncu for 16 MB of partials); x == 16 = the bias
    const int l = blockIdx.y, N = q.n_elem[l], G = q.groups;
    const float* __restrict__ part = q.partial[l];
    if (blockIdx.x == 16) {
        if ((int)threadIdx.x < q.cout_real[l]) {
            float acc = 0.f;
            for (int g = 0; g < G; ++g) acc += __ldcg(part + ((size_t)g * 10 + 9) * N + threadIdx.x);
            q.db[l][threadIdx.x] += acc;
        }
        return;
    }
    const int e = blockIdx.x * 256 + (int)threadIdx.x;
    if (e >= N) return;
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (int g0 = 0; g0 < G; g0 += 4) {      // groups in order (deterministic): 4 groups x 9 taps of loads in flight
        float v[4][9];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int t = 0; t < 9; ++t)
                v[u][t] = g0 + u < G ? __ldcg(part + ((size_t)(g0 + u) * 10 + t) * N + e) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (g0 + u < G) {
#pragma unroll
                for (int t = 0; t < 9; ++t) acc[t] += v[u][t];
            }
    }
    const int cin_pad = q.cin_pad[l], co = e / cin_pad, ci = e - co * cin_pad;
    if (co >= q.cout_real[l] || ci >= q.cin_real[l]) return;
    float* __restrict__ dst = q.dW[l] + ((size_t)co * q.cin_real[l] + ci) * 9;
    float oldw[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) oldw[t] = __ldcg(dst + t);
#pragma unroll
    for (int t = 0; t < 9; ++t) dst[t] = oldw[t] + acc[t];
}

// ---------------------------------------------------------------------------------------------------------------
// parameter preparation: fp16 conv weights in [COUT_PAD][9][CIN_PAD] (forward) and the rotated / transposed copy
// [CIN_PAD][9][COUT_PAD] with tap 8-t (data gradient), from torch's fp32 [cout][cin][3][3].
// ---------------------------------------------------------------------------------------------------------------
struct PrepParams {
    const float* w[17]; __half* wf[17]; __half* wr[17];
    __half* wf_tc[17]; __half* wr_tc[17];   // 64 -> 64 layers: [tap][64 x 64] canonical (rows = conv outputs, k = conv inputs)
    int cout[17], cin[17], cout_pad[17], cin_pad[17];
};
__global__ void __launch_bounds__(256)
k_dec_prep_weights(const PrepParams q) {
    const int l = blockIdx.y;
    const float* __restrict__ w = q.w[l];
    const int cout = q.cout[l], cin = q.cin[l], cout_pad = q.cout_pad[l], cin_pad = q.cin_pad[l];
    __half* __restrict__ wf = q.wf[l];
    __half* __restrict__ wr = q.wr[l];
    const int n = cout_pad * 9 * cin_pad;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        {   // forward layout: i = (co, t, ci)
            const int co = i / (9 * cin_pad), rem = i - co * 9 * cin_pad, t = rem / cin_pad, ci = rem - t * cin_pad;
            wf[i] = (co < cout && ci < cin) ? f2h(w[((size_t)co * cin + ci) * 9 + t]) : f2h(0.f);
        }
        {   // data-gradient layout: i = (ci, t, co) holds w[co][ci][8 - t]
            const int ci = i / (9 * cout_pad), rem = i - ci * 9 * cout_pad, t = rem / cout_pad, co = rem - t * cout_pad;
            wr[i] = (co < cout && ci < cin) ? f2h(w[((size_t)co * cin + ci) * 9 + (8 - t)]) : f2h(0.f);
        }
    }
    if (q.wf_tc[l] != nullptr) {   // element (row r, tap t, k) of either orientation -> t * 4096 + canonical(r, k) halfs
        __half* __restrict__ wf_tc = q.wf_tc[l];
        __half* __restrict__ wr_tc = q.wr_tc[l];
        for (int i = blockIdx.x * 256 + threadIdx.x; i < 64 * 9 * 64; i += gridDim.x * 256) {
            const int r = i / 576, rem = i - r * 576, t = rem >> 6, k = rem & 63;
            const int dst = t * 4096 + (r >> 3) * 512 + (k >> 3) * 64 + (r & 7) * 8 + (k & 7);
            wf_tc[dst] = f2h(w[((size_t)r * 64 + k) * 9 + t]);            // forward: r = co, k = ci
            wr_tc[dst] = f2h(w[((size_t)k * 64 + r) * 9 + (8 - t)]);      // data gradient: r = ci, k = co, taps mirrored
        }
    }
}

// image [B,H,W,3] fp32 -> x0 [B,H,W,16] fp16 = ((x - mean)/std, 0...)   (hidden_models.normalize_img + autocast cast)
__global__ void __launch_bounds__(256)
k_dec_prep_input(const float* __restrict__ img, int n_pix, __half* __restrict__ x0) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_pix) return;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, std[3] = {0.229f, 0.224f, 0.225f};
    __align__(16) __half o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = f2h(0.f);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = f2h(__fdiv_rn(__fsub_rn(img[(size_t)i * 3 + k], mean[k]), std[k]));
    *reinterpret_cast<uint4*>(x0 + (size_t)i * 16) = *reinterpret_cast<const uint4*>(o);
    *reinterpret_cast<uint4*>(x0 + (size_t)i * 16 + 8) = *reinterpret_cast<const uint4*>(o + 8);
}
// ---------------------------------------------------------------------------------------------------------------
// head: a9 = GELU(BN(z9)) [B,H,W,8 (nb real)] -> AdaptiveAvgPool2d(1) -> Linear(nb, nb) -> sum over redundancy.
// One CTA per image.  Backward: dlogits[B,num_bits] -> dlin_w, dlin_b (+=), da9 [B,H,W,8] fp16.
// ---------------------------------------------------------------------------------------------------------------
struct HeadParams {
    const __half* z9; BnSrc bn; const float* lin_w; const float* lin_b;
    int B, HW, nb, num_bits, redundancy;
    float* logits;          // fwd out [B, num_bits]
    __half* pooled;         // [B, 8] fp16 (saved for backward)
    const float* dlogits;   // bwd in [B, num_bits]
    float* dlin_w; float* dlin_b; __half* da9;
};

__global__ void __launch_bounds__(256)
k_dec_head_fwd(const HeadParams p) {
    __shared__ BnCoef coef[8];
    __shared__ float acc[8];
    if (threadIdx.x < 8) { coef[threadIdx.x] = threadIdx.x < p.nb ? bn_coef(p.bn, threadIdx.x, false) : BnCoef{}; acc[threadIdx.x] = 0.f; }
    __syncthreads();
    const int b = blockIdx.x;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int pix = threadIdx.x; pix < p.HW; pix += blockDim.x) {
        const uint4 v = *reinterpret_cast<const uint4*>(p.z9 + ((size_t)b * p.HW + pix) * 8);
        const __half* hz = reinterpret_cast<const __half*>(&v);
        for (int k = 0; k < p.nb; ++k) s[k] += h2f(act_from_z(hz[k], coef[k]));
    }
    {   // block sum in a fixed order: warp shuffles, then the per-warp partials one after the other
        __shared__ float wpart[kDecWarps][8];
        for (int k = 0; k < p.nb; ++k) {
            float v = s[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NSIG_FULL_MASK, v, o);
            if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5][k] = v;
        }
        __syncthreads();
        if ((int)threadIdx.x < p.nb) {
            float v = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += wpart[w][threadIdx.x];
            acc[threadIdx.x] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __half pooled[8];
        for (int k = 0; k < 8; ++k) { pooled[k] = f2h(k < p.nb ? acc[k] / (float)p.HW : 0.f); p.pooled[b * 8 + k] = pooled[k]; }
        // Linear under autocast: fp16 operands, fp32 accumulate, fp16 result; then sum over redundancy (hidden_models.py:130-135)
        for (int bit = 0; bit < p.num_bits; ++bit) {
            float tot = 0.f;
            for (int rdn = 0; rdn < p.redundancy; ++rdn) {
                const int o = bit * p.redundancy + rdn;
                float v = 0.f;
                for (int k = 0; k < p.nb; ++k) v = fmaf(h2f(f2h(p.lin_w[o * p.nb + k])), h2f(pooled[k]), v);
                tot += h2f(f2h(v + h2f(f2h(p.lin_b[o]))));
            }
            p.logits[b * p.num_bits + bit] = h2f(f2h(tot));
        }
    }
}

__global__ void __launch_bounds__(256)
k_dec_head_bwd(const HeadParams p) {
    __shared__ float dpool[8];
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        float dp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < p.nb; ++o) {
            const float dlo = h2f(f2h(p.dlogits[b * p.num_bits + o / p.redundancy]));   // gradient arrives in fp16 under autocast
            for (int k = 0; k < p.nb; ++k) dp[k] += dlo * h2f(f2h(p.lin_w[o * p.nb + k]));
        }
        for (int k = 0; k < 8; ++k) dpool[k] = h2f(f2h(dp[k])) / (float)p.HW;
    }
    // gradients of the linear layer: CTA 0, one thread per element, images summed in order (deterministic; nb <= 8)
    if (b == 0 && threadIdx.x >= 32 && (int)threadIdx.x - 32 < p.nb * (p.nb + 1)) {
        const int e = (int)threadIdx.x - 32, o = e / (p.nb + 1), k = e - o * (p.nb + 1);   // k == nb: the bias
        float acc = 0.f;
        for (int bb = 0; bb < p.B; ++bb) {
            const float dlo = h2f(f2h(p.dlogits[bb * p.num_bits + o / p.redundancy]));
            acc += k < p.nb ? dlo * h2f(p.pooled[bb * 8 + k]) : dlo;
        }
        if (k < p.nb) p.dlin_w[o * p.nb + k] += acc; else p.dlin_b[o] += acc;
    }
    __syncthreads();
    for (int pix = threadIdx.x; pix < p.HW; pix += blockDim.x) {
        __align__(16) __half o[8];
        for (int k = 0; k < 8; ++k) o[k] = f2h(k < p.nb ? dpool[k] : 0.f);
        *reinterpret_cast<uint4*>(p.da9 + ((size_t)b * p.HW + pix) * 8) = *reinterpret_cast<const uint4*>(o);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent forward: the whole decoder forward (weight/input preparation, L+1 convolutions, head) as ONE kernel.
// The per-layer kernels above are 9-11 us each for ~1 us of math: every launch pays a grid ramp-up, an L2 round trip for
// the BatchNorm sums and a pipeline drain, and BatchNorm's batch statistics force a global dependency between layers.
// Here the grid (one resident wave, at most one CTA per (image, strip) item) stays on the SMs and the layers are
// separated by a grid barrier (one atomic per CTA + an acquire spin on an L2 counter) instead of a kernel boundary.
// Arithmetic, rounding points and memory layout are exactly those of the per-layer kernels (same conv_cta body), which
// remain the path for the backward pass and for shapes whose items exceed what fits co-resident.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPersistMaxLayers = 9;   // conv blocks 0..L (the reference decoder: num_blocks = 8 -> 9 convs)

struct FwdPersistParams {
    ConvParams layer[kPersistMaxLayers];
    PrepParams prep;
    HeadParams head;
    const float* image;
    __half* x0;
    unsigned int* barrier;   // zero-initialised by the caller (inside the statistics memset)
    int L, n_pix, strips, n_items;
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                       // this CTA's global writes (and its warps' atomics) before the arrival
        atomicAdd(counter, 1u);
        unsigned int seen = 0, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
            if (seen < target && ++spins > (1u << 23)) __trap();   // a CTA that never arrives must not hang the GPU
        } while (seen < target);
    }
    __syncthreads();
}

__device__ __forceinline__ void prep_weights_slice(const PrepParams& q, int l, int gtid, int gthreads) {
    const float* __restrict__ w = q.w[l];
    const int cout = q.cout[l], cin = q.cin[l], cout_pad = q.cout_pad[l], cin_pad = q.cin_pad[l];
    __half* __restrict__ wf = q.wf[l];
    __half* __restrict__ wr = q.wr[l];
    const int n = cout_pad * 9 * cin_pad;
    for (int i = gtid; i < n; i += gthreads) {
        {
            const int co = i / (9 * cin_pad), rem = i - co * 9 * cin_pad, t = rem / cin_pad, ci = rem - t * cin_pad;
            wf[i] = (co < cout && ci < cin) ? f2h(w[((size_t)co * cin + ci) * 9 + t]) : f2h(0.f);
        }
        {
            const int ci = i / (9 * cout_pad), rem = i - ci * 9 * cout_pad, t = rem / cout_pad, co = rem - t * cout_pad;
            wr[i] = (co < cout && ci < cin) ? f2h(w[((size_t)co * cin + ci) * 9 + (8 - t)]) : f2h(0.f);
        }
    }
}

__device__ __forceinline__ void head_fwd_image(const HeadParams& p, int b, BnCoef* coef, float* acc) {
    __syncthreads();
    if (threadIdx.x < 8) { coef[threadIdx.x] = (int)threadIdx.x < p.nb ? bn_coef(p.bn, threadIdx.x, false) : BnCoef{}; acc[threadIdx.x] = 0.f; }
    __syncthreads();
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int pix = threadIdx.x; pix < p.HW; pix += blockDim.x) {
        const uint4 v = *reinterpret_cast<const uint4*>(p.z9 + ((size_t)b * p.HW + pix) * 8);
        const __half* hz = reinterpret_cast<const __half*>(&v);
        for (int k = 0; k < p.nb; ++k) s[k] += h2f(act_from_z(hz[k], coef[k]));
    }
    {   // block sum in a fixed order: warp shuffles, then the per-warp partials one after the other
        __shared__ float wpart[kDecWarps][8];
        for (int k = 0; k < p.nb; ++k) {
            float v = s[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NSIG_FULL_MASK, v, o);
            if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5][k] = v;
        }
        __syncthreads();
        if ((int)threadIdx.x < p.nb) {
            float v = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += wpart[w][threadIdx.x];
            acc[threadIdx.x] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __half pooled[8];
        for (int k = 0; k < 8; ++k) { pooled[k] = f2h(k < p.nb ? acc[k] / (float)p.HW : 0.f); p.pooled[b * 8 + k] = pooled[k]; }
        for (int bit = 0; bit < p.num_bits; ++bit) {
            float tot = 0.f;
            for (int rdn = 0; rdn < p.redundancy; ++rdn) {
                const int o = bit * p.redundancy + rdn;
                float v = 0.f;
                for (int k = 0; k < p.nb; ++k) v = fmaf(h2f(f2h(p.lin_w[o * p.nb + k])), h2f(pooled[k]), v);
                tot += h2f(f2h(v + h2f(f2h(p.lin_b[o]))));
            }
            p.logits[b * p.num_bits + bit] = h2f(f2h(tot));
        }
    }
}

__global__ void __launch_bounds__(kDecThreads)
k_dec_fwd_persist(const FwdPersistParams q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BnCoef head_coef[8];
    __shared__ float head_acc[8];
    const int gtid = blockIdx.x * kDecThreads + threadIdx.x, gthreads = gridDim.x * kDecThreads;
    const unsigned int n_cta = gridDim.x;
    // phase 0: fp16 weight copies (both orientations; the backward kernels read the rotated one) and the normalised input
    for (int l = 0; l <= q.L; ++l) prep_weights_slice(q.prep, l, gtid, gthreads);
    {
        const float mean[3] = {0.485f, 0.456f, 0.406f}, std[3] = {0.229f, 0.224f, 0.225f};
        for (int i = gtid; i < q.n_pix; i += gthreads) {
            __align__(16) __half o[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = f2h(0.f);
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = f2h(__fdiv_rn(__fsub_rn(q.image[(size_t)i * 3 + k], mean[k]), std[k]));
            *reinterpret_cast<uint4*>(q.x0 + (size_t)i * 16) = *reinterpret_cast<const uint4*>(o);
            *reinterpret_cast<uint4*>(q.x0 + (size_t)i * 16 + 8) = *reinterpret_cast<const uint4*>(o + 8);
        }
    }
    unsigned int phase = 1;
    grid_barrier(q.barrier, phase++ * n_cta);
    for (int l = 0; l <= q.L; ++l) {
        if (l == 0) conv_cta<16, 16, 64, IN_RAW>(q.layer[l], smem_raw, (int)blockIdx.x, (int)gridDim.x, q.n_items, q.strips);
        else if (l == q.L) conv_cta<64, 64, 8, IN_BNGELU>(q.layer[l], smem_raw, (int)blockIdx.x, (int)gridDim.x, q.n_items, q.strips);
        else conv_cta<64, 64, 64, IN_BNGELU>(q.layer[l], smem_raw, (int)blockIdx.x, (int)gridDim.x, q.n_items, q.strips);
        grid_barrier(q.barrier, phase++ * n_cta);
    }
    for (int b = blockIdx.x; b < q.head.B; b += gridDim.x) head_fwd_image(q.head, b, head_coef, head_acc);
}

struct BnGradParams { const double* bsums[17]; float* dgamma[17]; float* dbeta[17]; int c_pad[17], c_real[17]; };
__global__ void __launch_bounds__(64)
k_dec_bn_grads(const BnGradParams q) {   // dbeta = sum dy, dgamma = sum dy*yhat: the backward statistics themselves
    const int l = blockIdx.x, ch = threadIdx.x;
    if (ch < q.c_real[l]) {
        atomicAdd(q.dbeta[l] + ch, (float)q.bsums[l][ch]);
        atomicAdd(q.dgamma[l] + ch, (float)q.bsums[l][q.c_pad[l] + ch]);
    }
}

}  // namespace nsig

using namespace nsig;

// ---------------------------------------------------------------------------------------------------------------
// host side: workspace layout and the launch chains
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxLayers = 16;

struct DecLayout {
    int B, H, W, L;            // L = number of 64-channel conv blocks (num_blocks); block L is the nb-channel one
    size_t n_pix;
    size_t off_x0, off_z[kMaxLayers + 1], off_a[kMaxLayers + 1], off_da[2], off_dz[kMaxLayers + 1], off_da9, off_dx0, off_pooled;
    size_t off_wf[kMaxLayers + 1], off_wr[kMaxLayers + 1], off_sums[kMaxLayers + 1], off_bsums[kMaxLayers + 1];
    size_t off_wf_tc[kMaxLayers + 1], off_wr_tc[kMaxLayers + 1];   // UMMA-layout copies of the 64 -> 64 layers (0 = none)
    size_t off_sums_begin, sums_bytes, off_barrier, off_tickets, off_partial[kMaxLayers + 1], total;
};

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }
constexpr int kWgradGroupsMax = 32;   // upper bound of the weight-gradient grid's image groups (NSIG_DEC_WGRAD_G)

DecLayout make_layout(int B, int H, int W, int L) {
    DecLayout d{};
    d.B = B; d.H = H; d.W = W; d.L = L;
    d.n_pix = (size_t)B * H * W;
    size_t o = 0;
    d.off_x0 = o; o = align_up(o + d.n_pix * 16 * 2);
    for (int l = 0; l < L; ++l) { d.off_z[l] = o; o = align_up(o + d.n_pix * 64 * 2); }
    d.off_z[L] = o; o = align_up(o + d.n_pix * 8 * 2);
    for (int l = 0; l < L; ++l) { d.off_a[l] = o; o = align_up(o + d.n_pix * 64 * 2); }   // a_l = GELU(BN(z_l)), l < L
    for (int k = 0; k < 2; ++k) { d.off_da[k] = o; o = align_up(o + d.n_pix * 64 * 2); }
    // dz_l per layer: the weight-gradient kernels run on a side stream while the data-gradient chain goes on
    for (int l = 0; l <= L; ++l) { d.off_dz[l] = o; o = align_up(o + d.n_pix * (l == L ? 8 : 64) * 2); }
    d.off_da9 = o; o = align_up(o + d.n_pix * 8 * 2);
    d.off_dx0 = o; o = align_up(o + d.n_pix * 16 * 2);
    d.off_pooled = o; o = align_up(o + (size_t)B * 8 * 2);
    for (int l = 0; l <= L; ++l) {
        const int cin = l == 0 ? 16 : 64, cout = l == L ? 16 : 64;
        d.off_wf[l] = o; o = align_up(o + (size_t)cout * 9 * cin * 2);
        d.off_wr[l] = o; o = align_up(o + (size_t)cout * 9 * cin * 2);
        if (l > 0 && l < L) {
            d.off_wf_tc[l] = o; o = align_up(o + (size_t)64 * 9 * 64 * 2);
            d.off_wr_tc[l] = o; o = align_up(o + (size_t)64 * 9 * 64 * 2);
        }
    }
    d.off_sums_begin = o;
    for (int l = 0; l <= L; ++l) {
        d.off_sums[l] = o; o += 2 * 64 * sizeof(double);
        d.off_bsums[l] = o; o += 2 * 64 * sizeof(double);
    }
    d.off_barrier = o; o += 256;            // grid-barrier counter of the persistent forward, cleared with the statistics
    d.off_tickets = o; o += (size_t)(kMaxLayers + 1) * 16 * sizeof(unsigned int);   // weight-gradient tickets, per layer x tap
    d.sums_bytes = o - d.off_sums_begin;
    o = align_up(o);
    for (int l = 0; l <= L; ++l) {   // per-layer partials of the weight-gradient reduction (layers run concurrently)
        const int cin = l == 0 ? 16 : 64, cout = l == L ? 16 : 64;
        d.off_partial[l] = o; o = align_up(o + (size_t)kWgradGroupsMax * 10 * cout * cin * sizeof(float));
    }
    d.total = align_up(o);
    return d;
}

// fp16 weights of every layer in both orientations, as one buffer (nsig_decoder_prepare_weights): offsets of layer l
struct WeightLayout { size_t off_wf[kMaxLayers + 1], off_wr[kMaxLayers + 1], off_wf_tc[kMaxLayers + 1], off_wr_tc[kMaxLayers + 1], total; };
WeightLayout make_weight_layout(int L) {
    WeightLayout w{};
    size_t o = 0;
    for (int l = 0; l <= L; ++l) {
        const int cin = l == 0 ? 16 : 64, cout = l == L ? 16 : 64;
        w.off_wf[l] = o; o = align_up(o + (size_t)cout * 9 * cin * 2);
        w.off_wr[l] = o; o = align_up(o + (size_t)cout * 9 * cin * 2);
        if (l > 0 && l < L) {
            w.off_wf_tc[l] = o; o = align_up(o + (size_t)64 * 9 * 64 * 2);
            w.off_wr_tc[l] = o; o = align_up(o + (size_t)64 * 9 * 64 * 2);
        }
    }
    w.total = o;
    return w;
}
PrepParams make_prep(const float* const* params, int L, int nb, unsigned char* base, const size_t* off_wf, const size_t* off_wr,
                     const size_t* off_wf_tc, const size_t* off_wr_tc) {
    PrepParams q{};
    for (int l = 0; l <= L; ++l) {
        q.w[l] = params[4 * l];
        q.wf[l] = reinterpret_cast<__half*>(base + off_wf[l]); q.wr[l] = reinterpret_cast<__half*>(base + off_wr[l]);
        if (l > 0 && l < L) {
            q.wf_tc[l] = reinterpret_cast<__half*>(base + off_wf_tc[l]); q.wr_tc[l] = reinterpret_cast<__half*>(base + off_wr_tc[l]);
        }
        q.cin[l] = l == 0 ? 3 : 64; q.cout[l] = l == L ? nb : 64; q.cin_pad[l] = l == 0 ? 16 : 64; q.cout_pad[l] = l == L ? 16 : 64;
    }
    return q;
}

// image rows per conv CTA: at most 64 pixels (4 m-tiles per warp pass), at least ~148 CTAs when the batch allows it
int conv_rows(int B, int H, int W) {
    int R = 64 / W > 0 ? 64 / W : 1;
    const int strips = (148 + B - 1) / B;
    const int want = (H + strips - 1) / strips > 0 ? (H + strips - 1) / strips : 1;
    if (want < R) R = want;
    return R < H ? R : H;
}
// the tcgen05 conv needs the strip's rows over the padded width as one M = 128 operand
bool conv_tc_enabled() {
    static const bool on = [] { const char* e = getenv("NSIG_DEC_TC"); return !(e && e[0] == '0'); }();   // default on; NSIG_DEC_TC=0: mma.sync kernels
    return on;
}
template <int MODE>
int launch_conv_tc(ConvParams p, cudaStream_t st) {
    p.R = conv_rows(p.B, p.H, p.W);
    const size_t smem = tcv::smem_bytes(p.R, p.W);
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_dec_conv_tc<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    // NSIG_DEC_PDL=1: launched with programmatic stream serialization - the grid may begin once every CTA of the preceding
    // kernel has executed griddepcontrol.launch_dependents (k_dec_conv_tc does at entry; any other kernel implicitly when it
    // completes) and blocks in griddepcontrol.wait until that kernel has completed and flushed.  Measured (round 2, call BI):
    // the kernels do start up to 22 us early and sit in the wait with their weights loaded, but forward+backward takes 271 us
    // instead of 265 us - what a kernel boundary costs here is the completion + flush + wake-up, not the launch.  Off by default.
    static const bool pdl = [] { const char* e = getenv("NSIG_DEC_PDL"); return e && e[0] == '1'; }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((p.H + p.R - 1) / p.R, p.B);
    cfg.blockDim = dim3(kDecThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_dec_conv_tc<MODE>, p);
    if (e != cudaSuccess) return (int)e;
    NSIG_LAUNCH_CHECK();
    return 0;
}
bool conv_tc_fits(const ConvParams& p) {
    const int R = conv_rows(p.B, p.H, p.W);
    return conv_tc_enabled() && kDecThreads == 256 && R * (p.W + 2) <= 128 && tcv::smem_bytes(R, p.W) <= 200 * 1024 &&
           p.cout_valid == 64 && p.w_tc != nullptr && (((uintptr_t)p.w_tc) & 15) == 0;
}

// image rows per weight-gradient step: as many pixels as fit next to each other in ~160 KB of shared memory
int wgrad_rows(int H, int W, int cin, int cout) {
    const size_t per_pix = (size_t)(cin + 8 + cout + 8) * 2;
    int R = (int)((80 * 1024) / (per_pix * (size_t)W));   // two buffers (double-buffered staging) within 160 KB
    if (R < 1) return 0;
    return R < H ? R : H;
}

template <int CIN, int SCH, int COUT, int MODE>
int launch_conv(ConvParams p, cudaStream_t st) {
    p.R = conv_rows(p.B, p.H, p.W);
    const size_t smem = (size_t)(p.R + 2) * (p.W + 2) * (CIN + 8) * 2 + (size_t)(CIN + COUT) * sizeof(BnCoef) +
                        (size_t)COUT * (9 * CIN + 8) * 2;
    if (smem > 200 * 1024) return NSIG_EINVAL;
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_dec_conv<CIN, SCH, COUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    k_dec_conv<CIN, SCH, COUT, MODE><<<dim3((p.H + p.R - 1) / p.R, p.B), kDecThreads, smem, st>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int wgrad_groups(int B) {
    static const int g_max = [] {
        const char* e = getenv("NSIG_DEC_WGRAD_G");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? (v < kWgradGroupsMax ? v : kWgradGroupsMax) : 16;
    }();
    return B < g_max ? B : g_max;
}

template <int CIN, int COUT, int DCH>
int launch_wgrad_batch(const WgradParams* layers, int n, cudaStream_t st) {
    if (n < 1 || n > kWgradBatchMax) return NSIG_EINVAL;
    WgradBatch q{};
    const int R = wgrad_rows(layers[0].H, layers[0].W, CIN, COUT);
    if (R <= 0) return NSIG_EINVAL;
    for (int i = 0; i < n; ++i) { q.p[i] = layers[i]; q.p[i].R = R; }
    const size_t smem = 2 * (size_t)((R * layers[0].W + 15) / 16 * 16) * (CIN + 8 + COUT + 8) * 2;
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_dec_wgrad<CIN, COUT, DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    const int G = wgrad_groups(layers[0].B);   // image groups: 9*G CTAs per layer, each stores (or atomically adds) its [COUT x CIN] partial
    k_dec_wgrad<CIN, COUT, DCH><<<dim3(9, G, n), kDecThreads, smem, st>>>(q);
    NSIG_LAUNCH_CHECK();
    return 0;
}
template <int CIN, int COUT, int DCH>
int launch_wgrad(WgradParams p, cudaStream_t st) { return launch_wgrad_batch<CIN, COUT, DCH>(&p, 1, st); }

// the all-taps kernel needs two whole images (a with halo + dz) in shared memory and the two-phase reduction
size_t wgrad_all_smem(int H, int W) {
    const size_t P = (size_t)H * W, Ppad = (P + 15) / 16 * 16;
    return 2 * ((size_t)(H + 2) * (W + 2) + Ppad) * 72 * 2;
}
bool wgrad_all_fits(const WgradParams& p) {
    static const bool on = [] { const char* e = getenv("NSIG_DEC_WGRAD_PER_TAP"); return !(e && e[0] == '1'); }();
    return on && p.partial != nullptr && p.tickets == nullptr && wgrad_all_smem(p.H, p.W) <= 200 * 1024 && p.H * p.W < 65536;
}
int launch_wgrad_all(const WgradParams* layers, int n, cudaStream_t st) {
    if (n < 1 || n > kWgradBatchMax) return NSIG_EINVAL;
    WgradBatch q{};
    for (int i = 0; i < n; ++i) q.p[i] = layers[i];
    const size_t smem = wgrad_all_smem(layers[0].H, layers[0].W);
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_dec_wgrad_all, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    k_dec_wgrad_all<<<dim3(wgrad_groups(layers[0].B), n), kWgAllThreads, smem, st>>>(q);
    NSIG_LAUNCH_CHECK();
    return 0;
}

// Side stream of the backward chain.  The data-gradient convs form a dependent chain (layer l needs the statistics of
// da_l); the weight gradient of layer l only needs a_{l-1} (forward) and dz_l (written by the data-gradient conv of
// layer l), and nothing downstream of it but the optimizer.  It is launched on a second stream that forks from the
// caller's stream after that conv and joins it at the end of the call: in a stream capture this becomes a parallel
// branch of the graph, eagerly the kernels simply overlap (neither chain fills the 148 SMs).
constexpr int kMaxSideStreams = 4;
struct SideStream {
    cudaStream_t stream[kMaxSideStreams] = {};   // layer l's weight gradient goes to stream[l % n]
    cudaEvent_t fork[kMaxLayers + 1] = {};
    cudaEvent_t join[kMaxSideStreams] = {};
    int n = 0;
    bool ok = false;
};
std::mutex g_side_mutex;   // one backward launch chain at a time: the fork/join events are shared per device

SideStream* side_stream_for_current_device() {
    static SideStream side[64];
    static const bool disabled = [] { const char* e = getenv("NSIG_DEC_NO_SIDE"); return e && e[0] == '1'; }();
    int dev = 0;
    if (disabled || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    // one weight-gradient kernel is 9 x 16 CTAs of latency-bound work (two staged images, 36 MMAs per warp, 4096 atomics
    // per CTA): consecutive layers' kernels do not compete for anything, so they alternate between n streams and the
    // weight-gradient chain stops being longer than the data-gradient chain it hangs off
    static const int n_streams = [] {
        const char* e = getenv("NSIG_DEC_SIDE_STREAMS");
        const int v = e ? atoi(e) : 3;
        return v < 1 ? 1 : (v > kMaxSideStreams ? kMaxSideStreams : v);
    }();
    SideStream& s = side[dev];
    if (!s.ok) {
        for (int k = 0; k < n_streams; ++k) {
            if (cudaStreamCreateWithFlags(&s.stream[k], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&s.join[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        for (int l = 0; l <= kMaxLayers; ++l)
            if (cudaEventCreateWithFlags(&s.fork[l], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.n = n_streams;
        s.ok = true;
    }
    return &s;
}

}  // namespace

namespace {
// Deferred tail of nsig_decoder_backward (nsig_decoder_defer_weight_grads): the weight-gradient kernels are left running on the
// side streams and the caller's stream goes on with what needs the INPUT gradient only; nsig_decoder_finish_backward joins them
// and launches the group reduction.  One pending backward per device.
struct DeferredTail {
    bool pending = false;
    bool reduce = false;
    nsig::WgradReduceParams r{};
    int layers = 0;
};
DeferredTail g_deferred[64];
bool g_defer_on = false;

int finish_deferred_locked(int dev, cudaStream_t st) {
    DeferredTail& t = g_deferred[dev];
    if (!t.pending) return 0;
    t.pending = false;
    SideStream* side = side_stream_for_current_device();
    if (side) {
        for (int k = 0; k < side->n; ++k) {
            cudaError_t e = cudaEventRecord(side->join[k], side->stream[k]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, side->join[k], 0);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (t.reduce) {
        nsig::k_dec_wgrad_reduce<<<dim3(17, t.layers), 256, 0, st>>>(t.r);
        NSIG_LAUNCH_CHECK();
    }
    return 0;
}
}  // namespace

namespace nsig {
// parity probe: the decoder kernels' own GELU / GELU' device functions applied to n fp16 values
__global__ void __launch_bounds__(256)
k_dec_gelu_probe(const __half* __restrict__ y, uint32_t n, __half* __restrict__ gelu, __half* __restrict__ gelu_grad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = h2f(y[i]);
    gelu[i] = f2h(gelu_f(x));
    gelu_grad[i] = f2h(gelu_grad_f(x));
}
}  // namespace nsig

extern "C" {

size_t nsig_decoder_workspace_bytes(uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks) {
    if (num_blocks == 0 || num_blocks > (uint32_t)kMaxLayers) return 0;
    return make_layout((int)B, (int)H, (int)W, (int)num_blocks).total;
}

size_t nsig_decoder_weights_bytes(uint32_t num_blocks) {
    if (num_blocks == 0 || num_blocks > (uint32_t)kMaxLayers) return 0;
    return make_weight_layout((int)num_blocks).total;
}

// The fp16 copies of the conv weights (forward layout + rotated/transposed data-gradient layout) only change when the
// optimizer changes the weights: a caller that knows when that happens converts them once, off the critical path (e.g.
// right behind the optimizer kernel on its side stream), and hands the buffer to forward/backward as `prepared_weights`.
int nsig_decoder_prepare_weights(const float* const* params, uint32_t num_blocks, uint32_t num_bits, uint32_t redundancy,
                                 void* weights, nsig_stream_t stream) {
    const int L = (int)num_blocks, nb = (int)(num_bits * redundancy);
    if (!params || !weights || L < 1 || L > kMaxLayers || nb < 1 || nb > 8) return NSIG_EINVAL;
    const WeightLayout wl = make_weight_layout(L);
    const PrepParams q = make_prep(params, L, nb, reinterpret_cast<unsigned char*>(weights), wl.off_wf, wl.off_wr, wl.off_wf_tc, wl.off_wr_tc);
    k_dec_prep_weights<<<dim3(16, L + 1), 256, 0, (cudaStream_t)stream>>>(q);
    NSIG_LAUNCH_CHECK();
    return 0;
}

#ifdef NSIG_DEC_TRACE
int nsig_debug_dec_trace(unsigned long long* host_out /* [64][12] */, unsigned int* host_n, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, g_dec_trace, sizeof(unsigned long long) * 64 * 12);
    cudaMemcpyFromSymbol(host_n, g_dec_trace_n, sizeof(unsigned int));
    if (reset == 2) {   // the weight-gradient kernels' stamps instead
        cudaMemcpyFromSymbol(host_out, g_wg_trace, sizeof(unsigned long long) * 64 * 12);
        cudaMemcpyFromSymbol(host_n, g_wg_trace_n, sizeof(unsigned int));
        return 0;
    }
    if (reset) {
        static unsigned long long z[64][12];
        unsigned int zero = 0;
        cudaMemcpyToSymbol(g_dec_trace, z, sizeof(z));
        cudaMemcpyToSymbol(g_dec_trace_n, &zero, sizeof(zero));
        cudaMemcpyToSymbol(g_wg_trace, z, sizeof(z));
        cudaMemcpyToSymbol(g_wg_trace_n, &zero, sizeof(zero));
    }
    return 0;
}
#endif

int nsig_decoder_gelu_probe(const void* y, uint32_t n, void* gelu, void* gelu_grad, nsig_stream_t stream) {
    if (n == 0) return 0;
    if (!y || !gelu || !gelu_grad) return NSIG_EINVAL;
    k_dec_gelu_probe<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(y), n,
                                                                          reinterpret_cast<__half*>(gelu),
                                                                          reinterpret_cast<__half*>(gelu_grad));
    NSIG_LAUNCH_CHECK();
    return 0;
}

// params (HOST array of device pointers, fp32), per conv block l = 0..num_blocks (the last one has nb outputs):
//   params[4l+0] conv weight [cout,cin,3,3], [4l+1] conv bias, [4l+2] BN weight, [4l+3] BN bias;
//   then linear weight [nb,nb], linear bias [nb].
int nsig_decoder_forward(const float* image, uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks, uint32_t num_bits,
                         uint32_t redundancy, const float* const* params, void* workspace, float* logits,
                         const void* prepared_weights, nsig_stream_t stream) {
    if (B == 0) return 0;
    const int L = (int)num_blocks, nb = (int)(num_bits * redundancy);
    if (!image || !params || !workspace || !logits || L < 1 || L > kMaxLayers || nb < 1 || nb > 8 || H == 0 || W == 0) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const DecLayout d = make_layout((int)B, (int)H, (int)W, L);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    auto H16 = [&](size_t off) { return reinterpret_cast<__half*>(ws + off); };
    auto D64 = [&](size_t off) { return reinterpret_cast<double*>(ws + off); };
    cudaError_t e = cudaMemsetAsync(ws + d.off_sums_begin, 0, d.sums_bytes, st);
    if (e != cudaSuccess) return (int)e;
    const PrepParams q = make_prep(params, L, nb, ws, d.off_wf, d.off_wr, d.off_wf_tc, d.off_wr_tc);
    const WeightLayout wl = make_weight_layout(L);
    const unsigned char* pw = reinterpret_cast<const unsigned char*>(prepared_weights);
    auto WF = [&](int l) { return pw ? reinterpret_cast<const __half*>(pw + wl.off_wf[l]) : H16(d.off_wf[l]); };
    const float inv_n = 1.0f / (float)d.n_pix;
    auto conv_params = [&](int l) {
        ConvParams p{};
        p.B = (int)B; p.H = (int)H; p.W = (int)W;
        p.w = WF(l); p.bias = params[4 * l + 1]; p.dst = H16(d.off_z[l]); p.out_sums = D64(d.off_sums[l]);
        if (l > 0 && l < L) p.w_tc = pw ? reinterpret_cast<const __half*>(pw + wl.off_wf_tc[l]) : H16(d.off_wf_tc[l]);
        p.cout_valid = l == L ? nb : 64;
        if (l == 0) {
            p.src = H16(d.off_x0);
        } else {
            p.src = H16(d.off_z[l - 1]);
            p.act_out = H16(d.off_a[l - 1]);   // a_{l-1}, kept for the weight gradient of this layer
            p.bn = BnSrc{D64(d.off_sums[l - 1]), nullptr, params[4 * (l - 1) + 2], params[4 * (l - 1) + 3], inv_n, 64, 64};
        }
        return p;
    };
    HeadParams h{};
    h.z9 = H16(d.off_z[L]); h.bn = BnSrc{D64(d.off_sums[L]), nullptr, params[4 * L + 2], params[4 * L + 3], inv_n, 8, nb};
    h.lin_w = params[4 * (L + 1)]; h.lin_b = params[4 * (L + 1) + 1];
    h.B = (int)B; h.HW = (int)(H * W); h.nb = nb; h.num_bits = (int)num_bits; h.redundancy = (int)redundancy;
    h.logits = logits; h.pooled = H16(d.off_pooled);

    // ---- persistent path: one kernel, grid barriers between the layers ----
    // Measured (round 2, B=32 blocks of 12x12, graph replay): persistent 118.8 us vs per-layer kernels 110.8 us - ten grid
    // barriers (an L2 atomic + acquire spin each) cost more than ten kernel boundaries inside a CUDA graph, so the
    // per-layer chain stays the default; NSIG_DEC_PERSIST=1 selects the persistent kernel (tools/bench_decoder.py).
    static const bool persist = [] { const char* e = getenv("NSIG_DEC_PERSIST"); return e && e[0] == '1'; }();
    if (persist && !pw && L + 1 <= kPersistMaxLayers) {
        const int R = conv_rows((int)B, (int)H, (int)W);
        const int strips = ((int)H + R - 1) / R;
        const size_t smem = (size_t)(R + 2) * (W + 2) * (64 + 8) * 2 + (size_t)(64 + 64) * sizeof(BnCoef) + (size_t)64 * (9 * 64 + 8) * 2;
        if (smem <= 200 * 1024) {
            static bool set = false;
            if (!set) { cudaFuncSetAttribute(k_dec_fwd_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
            int per_sm = 0, dev = 0, sms = 148;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dec_fwd_persist, kDecThreads, smem) == cudaSuccess && per_sm >= 1) {
                if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                FwdPersistParams fp{};
                for (int l = 0; l <= L; ++l) { fp.layer[l] = conv_params(l); fp.layer[l].R = R; }
                fp.prep = q; fp.head = h; fp.image = image; fp.x0 = H16(d.off_x0);
                fp.barrier = reinterpret_cast<unsigned int*>(ws + d.off_barrier);
                fp.L = L; fp.n_pix = (int)d.n_pix; fp.strips = strips; fp.n_items = strips * (int)B;
                // one resident wave: every CTA must be on an SM for the grid barrier to complete.  One CTA per SM is
                // enough (the layers are latency-bound) and leaves room for kernels of parallel graph branches.
                const int cap = sms * 1;
                const int grid = fp.n_items < cap ? fp.n_items : cap;
                k_dec_fwd_persist<<<grid, kDecThreads, smem, st>>>(fp);
                NSIG_LAUNCH_CHECK();
                return 0;
            }
        }
    }

    if (!pw) {
        k_dec_prep_weights<<<dim3(16, L + 1), 256, 0, st>>>(q);
        NSIG_LAUNCH_CHECK();
    }
    k_dec_prep_input<<<(unsigned)((d.n_pix + 255) / 256), 256, 0, st>>>(image, (int)d.n_pix, H16(d.off_x0));
    NSIG_LAUNCH_CHECK();
    for (int l = 0; l <= L; ++l) {
        ConvParams p = conv_params(l);
        int rc;
        if (l == 0) rc = launch_conv<16, 16, 64, IN_RAW>(p, st);
        else if (l < L && conv_tc_fits(p)) rc = launch_conv_tc<IN_BNGELU>(p, st);
        else rc = l == L ? launch_conv<64, 64, 8, IN_BNGELU>(p, st) : launch_conv<64, 64, 64, IN_BNGELU>(p, st);
        if (rc) return rc;
    }
    k_dec_head_fwd<<<B, 256, 0, st>>>(h);
    NSIG_LAUNCH_CHECK();
    return 0;
}

// grads: HOST array of device pointers (fp32, ACCUMULATED into) in the order of `params`.
// dimage (optional) [B,H,W,3] fp32: gradient wrt the (un-normalised) input image.
int nsig_decoder_backward(const float* dlogits, uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks, uint32_t num_bits,
                          uint32_t redundancy, const float* const* params, float* const* grads, void* workspace,
                          float* dimage, const void* prepared_weights, nsig_stream_t stream) {
    if (B == 0) return 0;
    const int L = (int)num_blocks, nb = (int)(num_bits * redundancy);
    if (!dlogits || !params || !grads || !workspace || L < 1 || L > kMaxLayers || nb < 1 || nb > 8) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const DecLayout d = make_layout((int)B, (int)H, (int)W, L);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    auto H16 = [&](size_t off) { return reinterpret_cast<__half*>(ws + off); };
    auto D64 = [&](size_t off) { return reinterpret_cast<double*>(ws + off); };
    const float inv_n = 1.0f / (float)d.n_pix;
    const int n_pix = (int)d.n_pix;
    const WeightLayout wl = make_weight_layout(L);
    const unsigned char* pw = reinterpret_cast<const unsigned char*>(prepared_weights);   // must be what the forward used
    auto WR = [&](int l) { return pw ? reinterpret_cast<const __half*>(pw + wl.off_wr[l]) : H16(d.off_wr[l]); };

    HeadParams h{};
    h.lin_w = params[4 * (L + 1)]; h.B = (int)B; h.HW = (int)(H * W); h.nb = nb; h.num_bits = (int)num_bits;
    h.redundancy = (int)redundancy; h.pooled = H16(d.off_pooled); h.dlogits = dlogits;
    h.dlin_w = grads[4 * (L + 1)]; h.dlin_b = grads[4 * (L + 1) + 1]; h.da9 = H16(d.off_da9);
    k_dec_head_bwd<<<B, 256, 0, st>>>(h);
    NSIG_LAUNCH_CHECK();

    std::lock_guard<std::mutex> lock(g_side_mutex);
    SideStream* side = side_stream_for_current_device();
    int cur_dev = 0;
    if (cudaGetDevice(&cur_dev) != cudaSuccess || cur_dev < 0 || cur_dev >= 64) cur_dev = 0;
    if (int rc = finish_deferred_locked(cur_dev, st)) return rc;   // an earlier deferred tail nobody finished
    // weight-gradient reduction: tickets + partials (deterministic, default) or fp32 atomics (NSIG_DEC_WGRAD_ATOMIC=1)
    static const bool two_phase = [] { const char* e = getenv("NSIG_DEC_WGRAD_ATOMIC"); return !(e && e[0] == '1'); }();
    // second phase: one reduction launch for all layers behind the join (default) or, NSIG_DEC_WGRAD_TICKETS=1, inside every
    // weight-gradient kernel by the CTA that draws the last ticket of its tap (round-2 scheme; same bits)
    static const bool ticket_reduce = [] { const char* e = getenv("NSIG_DEC_WGRAD_TICKETS"); return e && e[0] == '1'; }();
    const __half* da = H16(d.off_da9);
    const int stat_blocks = n_pix / 64 < 64 ? (n_pix / 64 > 0 ? n_pix / 64 : 1) : 64;
    // NSIG_DEC_WGRAD_LATE=1: all weight-gradient kernels are issued behind the data-gradient chain (on the side streams, next to
    // each other) instead of next to it, layer by layer
    // Schedule of the weight gradients.  Default: ALL of them behind the data-gradient chain - the 64 -> 64 layers as ONE launch
    // (blockIdx.z = layer) and the two odd-shaped layers next to it on the side streams.  Issued layer by layer next to the
    // chain (NSIG_DEC_WGRAD_EARLY=1, the round-2 schedule) every data-gradient kernel waited ~7 us instead of ~3 us for its
    // predecessor's stragglers (phase trace: profiles/r02_decoder_phase_trace.txt).
    static const bool wgrad_late = [] { const char* e = getenv("NSIG_DEC_WGRAD_EARLY"); return !(e && e[0] == '1'); }();
    auto wgrad_params = [&](int l) {
        WgradParams w{};
        w.dz = H16(d.off_dz[l]); w.dW = grads[4 * l]; w.db = grads[4 * l + 1];
        w.B = (int)B; w.H = (int)H; w.W = (int)W; w.cin_real = l == 0 ? 3 : 64; w.cout_real = l == L ? nb : 64;
        w.a = l == 0 ? H16(d.off_x0) : H16(d.off_a[l - 1]);
        if (two_phase) {
            w.partial = reinterpret_cast<float*>(ws + d.off_partial[l]);
            if (ticket_reduce) w.tickets = reinterpret_cast<unsigned int*>(ws + d.off_tickets) + 16 * l;
        }
        return w;
    };
    auto fork_to = [&](int slot, cudaStream_t& wst) -> int {   // side stream `slot` continues from the current point of st
        wst = st;
        if (side) {
            cudaError_t e = cudaEventRecord(side->fork[slot], st);
            wst = side->stream[slot % side->n];
            if (e == cudaSuccess) e = cudaStreamWaitEvent(wst, side->fork[slot], 0);
            if (e != cudaSuccess) return (int)e;
        }
        return 0;
    };
    auto wgrad_layer = [&](int l) -> int {
        cudaStream_t wst;
        if (int rc = fork_to(l, wst)) return rc;
        const WgradParams w = wgrad_params(l);
        if (l == 0) return launch_wgrad<16, 64, 64>(w, wst);
        if (l == L) return launch_wgrad<64, 16, 8>(w, wst);
        return launch_wgrad<64, 64, 64>(w, wst);
    };
    for (int l = L; l >= 0; --l) {
        BnSrc bn{D64(d.off_sums[l]), D64(d.off_bsums[l]), params[4 * l + 2], params[4 * l + 3], inv_n, l == L ? 8 : 64, l == L ? nb : 64};
        // (1) backward statistics = dbeta, dgamma.  Only the last block needs a pass of its own (its da comes from the
        //     head); for every other layer the data-gradient conv that produced da_l accumulated them in its epilogue.
        if (l == L) {
            k_dec_bwd_stats<8><<<stat_blocks, kDecThreads, 0, st>>>(da, H16(d.off_z[l]), bn, D64(d.off_bsums[l]), n_pix);
            NSIG_LAUNCH_CHECK();
        }
        // (2) data gradient da_{l-1} = conv(dz_l, rotated weights); the staged dz_l is written out for (3)
        ConvParams p{};
        p.B = (int)B; p.H = (int)H; p.W = (int)W;
        p.src = da; p.src2 = H16(d.off_z[l]); p.bn = bn; p.w = WR(l); p.act_out = H16(d.off_dz[l]);
        if (l > 0 && l < L) p.w_tc = pw ? reinterpret_cast<const __half*>(pw + wl.off_wr_tc[l]) : H16(d.off_wr_tc[l]);
        __half* out = l == 0 ? H16(d.off_dx0) : H16(d.off_da[l & 1]);
        p.dst = out; p.cout_valid = l == 0 ? 3 : 64;
        if (l == 0) p.dimage = dimage;   // the input gradient leaves the chain's last kernel directly
        if (l > 0) {   // outputs are da_{l-1}: accumulate layer l-1's backward statistics on the way out
            p.z_out = H16(d.off_z[l - 1]);
            p.bn_out = BnSrc{D64(d.off_sums[l - 1]), nullptr, params[4 * (l - 1) + 2], params[4 * (l - 1) + 3], inv_n, 64, 64};
            p.out_bsums = D64(d.off_bsums[l - 1]);
        }
        int rc;
        if (l == L) rc = launch_conv<16, 8, 64, IN_DZ>(p, st);
        else if (l == 0) rc = launch_conv<64, 64, 16, IN_DZ>(p, st);
        else if (conv_tc_fits(p)) rc = launch_conv_tc<IN_DZ>(p, st);
        else rc = launch_conv<64, 64, 64, IN_DZ>(p, st);
        if (rc) return rc;
        // (3) weight / bias gradient from the materialised a_{l-1} and dz_l, on the side stream
        if (!wgrad_late) {
            rc = wgrad_layer(l);
            if (rc) return rc;
        }
        da = out;
    }
    if (wgrad_late) {
        cudaStream_t wst;
        if (int rc = fork_to(0, wst)) return rc;
        for (int l0 = 1; l0 < L; l0 += kWgradBatchMax) {   // the 64 -> 64 layers, up to 8 per launch
            WgradParams mid[kWgradBatchMax];
            int n = 0;
            for (int l = l0; l < L && n < kWgradBatchMax; ++l) mid[n++] = wgrad_params(l);
            if (int rc = wgrad_all_fits(mid[0]) ? launch_wgrad_all(mid, n, wst) : launch_wgrad_batch<64, 64, 64>(mid, n, wst)) return rc;
        }
        if (int rc = fork_to(1, wst)) return rc;
        if (int rc = launch_wgrad<16, 64, 64>(wgrad_params(0), wst)) return rc;
        if (int rc = fork_to(2, wst)) return rc;
        if (int rc = launch_wgrad<64, 16, 8>(wgrad_params(L), wst)) return rc;
    }
    // BatchNorm parameter gradients (= the backward statistics): nothing on the caller's stream needs them - with the weight
    // gradients on a side stream (the last one used above, or the caller's stream without side streams)
    {
        cudaStream_t bst = st;
        if (side && wgrad_late) bst = side->stream[2 % side->n];
        else if (side) { if (int rc = fork_to(0, bst)) return rc; }
        BnGradParams q{};
        for (int l = 0; l <= L; ++l) {
            q.bsums[l] = D64(d.off_bsums[l]); q.dgamma[l] = grads[4 * l + 2]; q.dbeta[l] = grads[4 * l + 3];
            q.c_pad[l] = l == L ? 8 : 64; q.c_real[l] = l == L ? nb : 64;
        }
        k_dec_bn_grads<<<L + 1, 64, 0, bst>>>(q);
        NSIG_LAUNCH_CHECK();
    }
    const bool defer = g_defer_on && side != nullptr && wgrad_late;
    if (side && !defer) {
        for (int k = 0; k < side->n && k <= L; ++k) {
            cudaError_t e = cudaEventRecord(side->join[k], side->stream[k]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(st, side->join[k], 0);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (defer) {
        DeferredTail& t = g_deferred[cur_dev];
        t.pending = true;
        t.reduce = false;
        t.layers = L + 1;
    }
    if (two_phase && !ticket_reduce) {
        WgradReduceParams r{};
        for (int l = 0; l <= L; ++l) {
            const int cin_pad = l == 0 ? 16 : 64, cout_pad = l == L ? 16 : 64;
            r.partial[l] = reinterpret_cast<const float*>(ws + d.off_partial[l]);
            r.dW[l] = grads[4 * l]; r.db[l] = grads[4 * l + 1];
            r.n_elem[l] = cin_pad * cout_pad; r.cin_pad[l] = cin_pad;
            r.cin_real[l] = l == 0 ? 3 : 64; r.cout_real[l] = l == L ? nb : 64;
        }
        r.groups = wgrad_groups((int)B);
        if (defer) {
            g_deferred[cur_dev].reduce = true;
            g_deferred[cur_dev].r = r;
        } else {
            k_dec_wgrad_reduce<<<dim3(17, L + 1), 256, 0, st>>>(r);
            NSIG_LAUNCH_CHECK();
        }
    }
    return 0;
}

int nsig_decoder_defer_weight_grads(int on) {
    std::lock_guard<std::mutex> lock(g_side_mutex);
    g_defer_on = on != 0;
    return 0;
}

int nsig_decoder_finish_backward(nsig_stream_t stream) {
    std::lock_guard<std::mutex> lock(g_side_mutex);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return NSIG_EINVAL;
    return finish_deferred_locked(dev, (cudaStream_t)stream);
}

}  // extern "C"
