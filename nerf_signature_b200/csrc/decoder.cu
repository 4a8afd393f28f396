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

namespace nsig {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = 8;
constexpr float kBnEps = 1e-3f;  // hidden_models.py:24

enum InMode { IN_RAW = 0, IN_BNGELU = 1, IN_DZ = 2 };

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
template <int C, int SCH, int MODE>
__device__ __forceinline__ void stage_tile(__half* tile, const __half* __restrict__ src, const __half* __restrict__ src2,
                                           const BnCoef* __restrict__ coef, int b, int r0, int R, int H, int W) {
    constexpr int STRIDE = C + 8;
    const int TW = W + 2, n_pos = (R + 2) * TW, chunks = C / 8;
    for (int i = threadIdx.x; i < n_pos * chunks; i += blockDim.x) {
        const int pos = i / chunks, ck = i - pos * chunks;
        const int tr = pos / TW, tc = pos - tr * TW;
        const int r = r0 - 1 + tr, w = tc - 1;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (r >= 0 && r < H && w >= 0 && w < W && ck * 8 < SCH) {
            const size_t off = (((size_t)b * H + r) * W + w) * SCH + ck * 8;
            const uint4 v = *reinterpret_cast<const uint4*>(src + off);
            if (MODE == IN_RAW) {
                out = v;
            } else {
                const __half* hv = reinterpret_cast<const __half*>(&v);
                __half* ho = reinterpret_cast<__half*>(&out);
                if (MODE == IN_BNGELU) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) ho[k] = act_from_z(hv[k], coef[ck * 8 + k]);
                } else {
                    const uint4 v2 = *reinterpret_cast<const uint4*>(src2 + off);
                    const __half* hz = reinterpret_cast<const __half*>(&v2);
#pragma unroll
                    for (int k = 0; k < 8; ++k) ho[k] = dz_from(hv[k], hz[k], coef[ck * 8 + k]);
                }
            }
        }
        *reinterpret_cast<uint4*>(tile + (size_t)pos * STRIDE + ck * 8) = out;
    }
}

struct ConvParams {
    const __half* src;     // [B,H,W,CIN]   activation (IN_RAW), z of the producing layer (IN_BNGELU) or da (IN_DZ)
    const __half* src2;    // z (IN_DZ)
    BnSrc bn;              // BatchNorm state of the input transform
    const __half* w;       // [COUT][9][CIN] fp16 (already rotated/transposed for data gradients)
    const float* bias;     // [COUT] or null
    __half* dst;           // [B,H,W,COUT]
    double* out_sums;      // [2][COUT] or null: += sum / sum of squares of the (fp16-rounded) outputs
    int B, H, W, R;        // R = image rows per CTA
    int cout_valid;        // outputs >= cout_valid are written as zero (channel padding)
};

// ---------------------------------------------------------------------------------------------------------------
// 3x3 / pad 1 convolution as an implicit GEMM.  grid = (ceil(H/R), B).
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int SCH, int COUT, int MODE>
__global__ void __launch_bounds__(kDecThreads)
k_dec_conv(const ConvParams p) {
    constexpr int STRIDE = CIN + 8, WSTRIDE = 9 * CIN + 8, NT = COUT / 8, KS = CIN / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* wsm = reinterpret_cast<__half*>(smem_raw);                       // [COUT][WSTRIDE]
    __half* tile = wsm + COUT * WSTRIDE;                                     // [(R+2)*(W+2)][STRIDE]
    BnCoef* coef = reinterpret_cast<BnCoef*>(tile + (size_t)(p.R + 2) * (p.W + 2) * STRIDE);
    const int b = blockIdx.y, r0 = blockIdx.x * p.R, R = min(p.R, p.H - r0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;

    if (MODE != IN_RAW) {
        for (int ch = threadIdx.x; ch < SCH; ch += blockDim.x) coef[ch] = bn_coef(p.bn, ch, MODE == IN_DZ);
    }
    for (int i = threadIdx.x; i < COUT * (9 * CIN / 8); i += blockDim.x) {
        const int row = i / (9 * CIN / 8), ck = i - row * (9 * CIN / 8);
        *reinterpret_cast<uint4*>(wsm + row * WSTRIDE + ck * 8) = *reinterpret_cast<const uint4*>(p.w + (size_t)row * 9 * CIN + ck * 8);
    }
    __syncthreads();
    stage_tile<CIN, SCH, MODE>(tile, p.src, p.src2, coef, b, r0, R, p.H, p.W);
    __syncthreads();

    const int P = R * p.W, TW = p.W + 2, m_tiles = (P + 15) / 16;
    for (int item = warp; item < m_tiles * NT; item += kDecWarps) {
        const int mt = item / NT, nt = item - mt * NT;
        // this lane's A row for ldmatrix: pixel mt*16 + (lane & 15), clamped; centre tap position in the tile
        const int pa = min(mt * 16 + (lane & 15), P - 1);
        const int ra = pa / p.W, wa = pa - ra * p.W;
        const __half* arow = tile + ((size_t)(ra + 1) * TW + (wa + 1)) * STRIDE + (lane >> 4) * 8;
        const __half* wrow = wsm + (nt * 8 + g) * WSTRIDE + 2 * tig;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int dy = t / 3 - 1, dx = t % 3 - 1;
            const __half* at = arow + (dy * TW + dx) * STRIDE;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t a[4];
                ldsm_x4(a, at + ks * 16);
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wrow + t * CIN + ks * 16);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wrow + t * CIN + ks * 16 + 8);
                mma_16816(c, a, b0, b1);
            }
        }
        // epilogue: rows g, g+8 of the m-tile; columns nt*8 + 2tig, +1
        const int col = nt * 8 + 2 * tig;
        const float bias0 = (p.bias && col < p.cout_valid) ? p.bias[col] : 0.f;
        const float bias1 = (p.bias && col + 1 < p.cout_valid) ? p.bias[col + 1] : 0.f;
        float l1[2] = {0.f, 0.f}, l2[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pp = mt * 16 + h * 8 + g;
            if (pp < P) {
                const int rr = pp / p.W, ww = pp - rr * p.W;
                __half o0 = f2h(c[2 * h] + bias0), o1 = f2h(c[2 * h + 1] + bias1);
                if (col >= p.cout_valid) o0 = f2h(0.f);
                if (col + 1 >= p.cout_valid) o1 = f2h(0.f);
                const __half2 o = __halves2half2(o0, o1);
                *reinterpret_cast<__half2*>(p.dst + (((size_t)b * p.H + r0 + rr) * p.W + ww) * COUT + col) = o;
                const float f0 = h2f(o0), f1 = h2f(o1);
                l1[0] += f0; l1[1] += f1; l2[0] += f0 * f0; l2[1] += f1 * f1;
            }
        }
        if (p.out_sums) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a1 = l1[e], a2 = l2[e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a1 += __shfl_xor_sync(NSIG_FULL_MASK, a1, o); a2 += __shfl_xor_sync(NSIG_FULL_MASK, a2, o); }
                if (g == 0 && col + e < p.cout_valid) {
                    atomicAdd(p.out_sums + col + e, (double)a1);
                    atomicAdd(p.out_sums + COUT + col + e, (double)a2);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward statistics of one layer: bsums[0][c] += sum dy, bsums[1][c] += sum dy*yhat   (dy = da * GELU'(BN(z)))
// grid = any, grid-stride over pixels; C channels (64 or 8)
// ---------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kDecThreads)
k_dec_bwd_stats(const __half* __restrict__ da, const __half* __restrict__ z, BnSrc bn, double* __restrict__ bsums, int n_pix) {
    __shared__ BnCoef coef[C];
    __shared__ float red[2][C];
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) { coef[ch] = bn_coef(bn, ch, false); red[0][ch] = 0.f; red[1][ch] = 0.f; }
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
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&red[0][slot * 8 + k], a1[k]); atomicAdd(&red[1][slot * 8 + k], a2[k]); }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        atomicAdd(bsums + ch, (double)red[0][ch]);
        atomicAdd(bsums + C + ch, (double)red[1][ch]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient of one conv layer: dW[co][tap][ci] += sum_pixels dz[p][co] * a[p + tap][ci];  db[co] += sum dz.
// grid = (ceil(H/R), B); dW fp32 [COUT_REAL][CIN_REAL][3][3] (torch layout), accumulated with atomics.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
    const __half* a_src;   // input of the layer: x0 (IN_RAW) or z of the previous layer (IN_BNGELU)   [B,H,W,CIN]
    BnSrc a_bn;
    const __half* da;      // gradient wrt this layer's activation                                       [B,H,W,COUT]
    const __half* z;       // this layer's conv output
    BnSrc bn;              // this layer's BatchNorm state (incl. backward sums)
    float* dW;             // [cout_real][cin_real][3][3]
    float* db;             // [cout_real]
    int B, H, W, R, cin_real, cout_real;
};

// COUT = output channels padded to the MMA's M granularity (16); DCH = channels per pixel of da / z in memory.
template <int CIN, int COUT, int DCH, int AMODE>
__global__ void __launch_bounds__(kDecThreads)
k_dec_wgrad(const WgradParams p) {
    constexpr int ASTR = CIN + 8, DSTR = COUT + 8, MT = COUT / 16, NT = CIN / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* atile = reinterpret_cast<__half*>(smem_raw);                         // [(R+2)*(W+2)][ASTR]
    const int TW = p.W + 2;
    const int b = blockIdx.y, r0 = blockIdx.x * p.R, R = min(p.R, p.H - r0);
    const int P = R * p.W, Ppad = (P + 15) / 16 * 16;
    __half* dtile = atile + (size_t)(p.R + 2) * TW * ASTR;                       // [Ppad_max][DSTR]
    const int Ppad_max = (p.R * p.W + 15) / 16 * 16;
    BnCoef* coef_a = reinterpret_cast<BnCoef*>(dtile + (size_t)Ppad_max * DSTR);
    BnCoef* coef_d = coef_a + CIN;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;

    if (AMODE == IN_BNGELU)
        for (int ch = threadIdx.x; ch < CIN; ch += blockDim.x) coef_a[ch] = bn_coef(p.a_bn, ch, false);
    for (int ch = threadIdx.x; ch < DCH; ch += blockDim.x) coef_d[ch] = bn_coef(p.bn, ch, true);
    __syncthreads();
    stage_tile<CIN, CIN, AMODE>(atile, p.a_src, nullptr, coef_a, b, r0, R, p.H, p.W);
    // dz tile: rows = pixels of the strip (no halo), zero rows up to a multiple of 16
    for (int i = threadIdx.x; i < Ppad * (COUT / 8); i += blockDim.x) {
        const int pix = i / (COUT / 8), ck = i - pix * (COUT / 8);
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (pix < P && ck * 8 < DCH) {
            const int rr = pix / p.W, ww = pix - rr * p.W;
            const size_t off = (((size_t)b * p.H + r0 + rr) * p.W + ww) * DCH + ck * 8;
            const uint4 v = *reinterpret_cast<const uint4*>(p.da + off);
            const uint4 v2 = *reinterpret_cast<const uint4*>(p.z + off);
            const __half* hd = reinterpret_cast<const __half*>(&v);
            const __half* hz = reinterpret_cast<const __half*>(&v2);
            __half* ho = reinterpret_cast<__half*>(&out);
#pragma unroll
            for (int k = 0; k < 8; ++k) ho[k] = (ck * 8 + k < p.cout_real) ? dz_from(hd[k], hz[k], coef_d[ck * 8 + k]) : f2h(0.f);
        }
        *reinterpret_cast<uint4*>(dtile + (size_t)pix * DSTR + ck * 8) = out;
    }
    __syncthreads();

    // bias gradient: column sums of the dz tile
    for (int co = threadIdx.x; co < p.cout_real; co += blockDim.x) {
        float s = 0.f;
        for (int pix = 0; pix < P; ++pix) s += h2f(dtile[(size_t)pix * DSTR + co]);
        atomicAdd(p.db + co, s);
    }

    // items: (tap, m-tile of 16 couts, n-tile of 8 cins)
    for (int item = warp; item < 9 * MT * NT; item += kDecWarps) {
        const int t = item / (MT * NT), rem = item - t * (MT * NT), mt = rem / NT, nt = rem - mt * NT;
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k0 = 0; k0 < Ppad; k0 += 16) {
            // A = dz^T: matrices (m 0-7,k 0-7), (m 8-15,k 0-7), (m 0-7,k 8-15), (m 8-15,k 8-15) = transposed 8x8 blocks of
            // dtile rows (pixels) k0 + (lane&7) + 8*(lane>>4), columns mt*16 + 8*((lane>>3)&1)
            uint32_t a[4];
            ldsm_x4_trans(a, dtile + (size_t)(k0 + (lane & 7) + ((lane >> 4) << 3)) * DSTR + mt * 16 + (((lane >> 3) & 1) << 3));
            // B[k = pixel][n = ci]: transposed 8x8 blocks of the shifted a rows; lanes 0-7 -> k0..k0+7, lanes 8-15 -> k0+8..
            const int pk = min(k0 + (lane & 15), P - 1);     // rows >= P multiply zero dz rows
            const int rk = pk / p.W, wk = pk - rk * p.W;
            uint32_t bb[2];
            ldsm_x2_trans(bb, atile + ((size_t)(rk + 1 + dy) * TW + (wk + 1 + dx)) * ASTR + nt * 8);
            mma_16816(c, a, bb[0], bb[1]);
        }
        // c[0],c[1]: (co = mt*16+g, ci = nt*8+2tig, +1); c[2],c[3]: co + 8
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int co = mt * 16 + h * 8 + g;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ci = nt * 8 + 2 * tig + e;
                if (co < p.cout_real && ci < p.cin_real) atomicAdd(p.dW + ((size_t)co * p.cin_real + ci) * 9 + t, c[2 * h + e]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// parameter preparation: fp16 conv weights in [COUT_PAD][9][CIN_PAD] (forward) and the rotated / transposed copy
// [CIN_PAD][9][COUT_PAD] with tap 8-t (data gradient), from torch's fp32 [cout][cin][3][3].
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dec_prep_weights(const float* __restrict__ w, int cout, int cin, int cout_pad, int cin_pad, __half* __restrict__ wf,
                   __half* __restrict__ wr) {
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
// dx0 [B,H,W,16] fp16 -> dimg [B,H,W,3] fp32 = dx0 / std
__global__ void __launch_bounds__(256)
k_dec_input_grad(const __half* __restrict__ dx0, int n_pix, float* __restrict__ dimg) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_pix) return;
    const float std[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
    for (int k = 0; k < 3; ++k) dimg[(size_t)i * 3 + k] = __fdiv_rn(h2f(dx0[(size_t)i * 16 + k]), std[k]);
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
    for (int k = 0; k < p.nb; ++k) atomicAdd(&acc[k], s[k]);
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
            atomicAdd(p.dlin_b + o, dlo);
            for (int k = 0; k < p.nb; ++k) {
                atomicAdd(p.dlin_w + o * p.nb + k, dlo * h2f(p.pooled[b * 8 + k]));
                dp[k] += dlo * h2f(f2h(p.lin_w[o * p.nb + k]));
            }
        }
        for (int k = 0; k < 8; ++k) dpool[k] = h2f(f2h(dp[k])) / (float)p.HW;
    }
    __syncthreads();
    for (int pix = threadIdx.x; pix < p.HW; pix += blockDim.x) {
        __align__(16) __half o[8];
        for (int k = 0; k < 8; ++k) o[k] = f2h(k < p.nb ? dpool[k] : 0.f);
        *reinterpret_cast<uint4*>(p.da9 + ((size_t)b * p.HW + pix) * 8) = *reinterpret_cast<const uint4*>(o);
    }
}

__global__ void __launch_bounds__(64)
k_dec_bn_grads(const double* __restrict__ bsums, int C_pad, int c_real, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int ch = threadIdx.x;
    if (ch < c_real) { atomicAdd(dbeta + ch, (float)bsums[ch]); atomicAdd(dgamma + ch, (float)bsums[C_pad + ch]); }
}

}  // namespace nsig

using namespace nsig;

// ---------------------------------------------------------------------------------------------------------------
// host side: workspace layout and the launch chains
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxLayers = 16;

struct DecLayout {
    int B, H, W, L;            // L = number of 64-channel conv blocks (num_blocks); layer L+1 is the nb-channel block
    size_t n_pix;
    size_t off_x0, off_z[kMaxLayers + 1], off_da[2], off_da9, off_dx0, off_pooled;
    size_t off_wf[kMaxLayers + 1], off_wr[kMaxLayers + 1], off_sums[kMaxLayers + 1], off_bsums[kMaxLayers + 1];
    size_t off_sums_begin, sums_bytes, total;
};

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

DecLayout make_layout(int B, int H, int W, int L) {
    DecLayout d{};
    d.B = B; d.H = H; d.W = W; d.L = L;
    d.n_pix = (size_t)B * H * W;
    size_t o = 0;
    d.off_x0 = o; o = align_up(o + d.n_pix * 16 * 2);
    for (int l = 0; l < L; ++l) { d.off_z[l] = o; o = align_up(o + d.n_pix * 64 * 2); }
    d.off_z[L] = o; o = align_up(o + d.n_pix * 8 * 2);
    for (int k = 0; k < 2; ++k) { d.off_da[k] = o; o = align_up(o + d.n_pix * 64 * 2); }
    d.off_da9 = o; o = align_up(o + d.n_pix * 8 * 2);
    d.off_dx0 = o; o = align_up(o + d.n_pix * 16 * 2);
    d.off_pooled = o; o = align_up(o + (size_t)B * 8 * 2);
    for (int l = 0; l <= L; ++l) {
        const int cin = l == 0 ? 16 : 64, cout = l == L ? 16 : 64;
        d.off_wf[l] = o; o = align_up(o + (size_t)cout * 9 * cin * 2);
        d.off_wr[l] = o; o = align_up(o + (size_t)cout * 9 * cin * 2);
    }
    d.off_sums_begin = o;
    for (int l = 0; l <= L; ++l) {
        d.off_sums[l] = o; o += 2 * 64 * sizeof(double);
        d.off_bsums[l] = o; o += 2 * 64 * sizeof(double);
    }
    d.sums_bytes = o - d.off_sums_begin;
    d.total = align_up(o);
    return d;
}

int pick_rows(int B, int H, int W, int cin, int cout, bool wgrad) {
    // rows per CTA: enough strips to fill the 148 SMs (conv), few strips for the weight gradient (every CTA adds a full
    // dW with atomics), and staged tiles below ~96 KB next to the 75 KB of weights (conv) / the dz tile (wgrad)
    const int strips = wgrad ? 2 : (148 + B - 1) / B;
    const int want = (H + strips - 1) / strips > 0 ? (H + strips - 1) / strips : 1;
    for (int R = want; R >= 1; --R) {
        size_t bytes = (size_t)(R + 2) * (W + 2) * (cin + 8) * 2;
        if (wgrad) bytes += (size_t)((R * W + 15) / 16 * 16) * (cout + 8) * 2;
        if (bytes <= 96 * 1024) return R < H ? R : H;
    }
    return 0;
}

template <int CIN, int SCH, int COUT, int MODE>
int launch_conv(ConvParams p, cudaStream_t st) {
    p.R = pick_rows(p.B, p.H, p.W, CIN, COUT, false);
    if (p.R <= 0) return NSIG_EINVAL;
    const size_t smem = (size_t)COUT * (9 * CIN + 8) * 2 + (size_t)(p.R + 2) * (p.W + 2) * (CIN + 8) * 2 + CIN * sizeof(BnCoef);
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_dec_conv<CIN, SCH, COUT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    k_dec_conv<CIN, SCH, COUT, MODE><<<dim3((p.H + p.R - 1) / p.R, p.B), kDecThreads, smem, st>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

template <int CIN, int COUT, int DCH, int AMODE>
int launch_wgrad(WgradParams p, cudaStream_t st) {
    p.R = pick_rows(p.B, p.H, p.W, CIN, COUT, true);
    if (p.R <= 0) return NSIG_EINVAL;
    const size_t smem = (size_t)(p.R + 2) * (p.W + 2) * (CIN + 8) * 2 + (size_t)((p.R * p.W + 15) / 16 * 16) * (COUT + 8) * 2 +
                        (CIN + COUT) * sizeof(BnCoef);
    static bool set = false;
    if (!set) { cudaFuncSetAttribute(k_dec_wgrad<CIN, COUT, DCH, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); set = true; }
    k_dec_wgrad<CIN, COUT, DCH, AMODE><<<dim3((p.H + p.R - 1) / p.R, p.B), kDecThreads, smem, st>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" {

size_t nsig_decoder_workspace_bytes(uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks) {
    if (num_blocks == 0 || num_blocks > (uint32_t)kMaxLayers) return 0;
    return make_layout((int)B, (int)H, (int)W, (int)num_blocks).total;
}

// params (HOST array of device pointers, fp32), per conv block l = 0..num_blocks (the last one has nb outputs):
//   params[4l+0] conv weight [cout,cin,3,3], [4l+1] conv bias, [4l+2] BN weight, [4l+3] BN bias;
//   then linear weight [nb,nb], linear bias [nb].
int nsig_decoder_forward(const float* image, uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks, uint32_t num_bits,
                         uint32_t redundancy, const float* const* params, void* workspace, float* logits,
                         nsig_stream_t stream) {
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
    for (int l = 0; l <= L; ++l) {
        const int cin = l == 0 ? 3 : 64, cout = l == L ? nb : 64, cin_p = l == 0 ? 16 : 64, cout_p = l == L ? 16 : 64;
        k_dec_prep_weights<<<32, 256, 0, st>>>(params[4 * l], cout, cin, cout_p, cin_p, H16(d.off_wf[l]), H16(d.off_wr[l]));
        NSIG_LAUNCH_CHECK();
    }
    k_dec_prep_input<<<(unsigned)((d.n_pix + 255) / 256), 256, 0, st>>>(image, (int)d.n_pix, H16(d.off_x0));
    NSIG_LAUNCH_CHECK();
    const float inv_n = 1.0f / (float)d.n_pix;
    for (int l = 0; l <= L; ++l) {
        ConvParams p{};
        p.B = (int)B; p.H = (int)H; p.W = (int)W;
        p.w = H16(d.off_wf[l]); p.bias = params[4 * l + 1]; p.dst = H16(d.off_z[l]); p.out_sums = D64(d.off_sums[l]);
        p.cout_valid = l == L ? nb : 64;
        int rc;
        if (l == 0) {
            p.src = H16(d.off_x0);
            rc = launch_conv<16, 16, 64, IN_RAW>(p, st);
        } else {
            p.src = H16(d.off_z[l - 1]);
            p.bn = BnSrc{D64(d.off_sums[l - 1]), nullptr, params[4 * (l - 1) + 2], params[4 * (l - 1) + 3], inv_n, 64, 64};
            rc = l == L ? launch_conv<64, 64, 8, IN_BNGELU>(p, st) : launch_conv<64, 64, 64, IN_BNGELU>(p, st);
        }
        if (rc) return rc;
    }
    HeadParams h{};
    h.z9 = H16(d.off_z[L]); h.bn = BnSrc{D64(d.off_sums[L]), nullptr, params[4 * L + 2], params[4 * L + 3], inv_n, 8, nb};
    h.lin_w = params[4 * (L + 1)]; h.lin_b = params[4 * (L + 1) + 1];
    h.B = (int)B; h.HW = (int)(H * W); h.nb = nb; h.num_bits = (int)num_bits; h.redundancy = (int)redundancy;
    h.logits = logits; h.pooled = H16(d.off_pooled);
    k_dec_head_fwd<<<B, 256, 0, st>>>(h);
    NSIG_LAUNCH_CHECK();
    return 0;
}

// grads: HOST array of device pointers (fp32, ACCUMULATED into) in the order of `params`.
// dimage (optional) [B,H,W,3] fp32: gradient wrt the (un-normalised) input image.
int nsig_decoder_backward(const float* dlogits, uint32_t B, uint32_t H, uint32_t W, uint32_t num_blocks, uint32_t num_bits,
                          uint32_t redundancy, const float* const* params, float* const* grads, void* workspace,
                          float* dimage, nsig_stream_t stream) {
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

    HeadParams h{};
    h.lin_w = params[4 * (L + 1)]; h.B = (int)B; h.HW = (int)(H * W); h.nb = nb; h.num_bits = (int)num_bits;
    h.redundancy = (int)redundancy; h.pooled = H16(d.off_pooled); h.dlogits = dlogits;
    h.dlin_w = grads[4 * (L + 1)]; h.dlin_b = grads[4 * (L + 1) + 1]; h.da9 = H16(d.off_da9);
    k_dec_head_bwd<<<B, 256, 0, st>>>(h);
    NSIG_LAUNCH_CHECK();

    const __half* da = H16(d.off_da9);
    for (int l = L; l >= 0; --l) {
        const int stat_blocks = 148;
        BnSrc bn{D64(d.off_sums[l]), D64(d.off_bsums[l]), params[4 * l + 2], params[4 * l + 3], inv_n, l == L ? 8 : 64, l == L ? nb : 64};
        // (1) backward statistics = dbeta, dgamma
        if (l == L) k_dec_bwd_stats<8><<<stat_blocks, kDecThreads, 0, st>>>(da, H16(d.off_z[l]), bn, D64(d.off_bsums[l]), n_pix);
        else k_dec_bwd_stats<64><<<stat_blocks, kDecThreads, 0, st>>>(da, H16(d.off_z[l]), bn, D64(d.off_bsums[l]), n_pix);
        NSIG_LAUNCH_CHECK();
        // (2) weight / bias gradient
        WgradParams w{};
        w.da = da; w.z = H16(d.off_z[l]); w.bn = bn; w.dW = grads[4 * l]; w.db = grads[4 * l + 1];
        w.B = (int)B; w.H = (int)H; w.W = (int)W; w.cin_real = l == 0 ? 3 : 64; w.cout_real = l == L ? nb : 64;
        int rc;
        if (l == 0) {
            w.a_src = H16(d.off_x0);
            rc = launch_wgrad<16, 64, 64, IN_RAW>(w, st);
        } else {
            w.a_src = H16(d.off_z[l - 1]);
            w.a_bn = BnSrc{D64(d.off_sums[l - 1]), nullptr, params[4 * (l - 1) + 2], params[4 * (l - 1) + 3], inv_n, 64, 64};
            rc = l == L ? launch_wgrad<64, 16, 8, IN_BNGELU>(w, st) : launch_wgrad<64, 64, 64, IN_BNGELU>(w, st);
        }
        if (rc) return rc;
        // (3) data gradient: da_{l-1} = conv(dz_l, rotated weights)
        if (l > 0 || dimage) {
            ConvParams p{};
            p.B = (int)B; p.H = (int)H; p.W = (int)W;
            p.src = da; p.src2 = H16(d.off_z[l]); p.bn = bn; p.w = H16(d.off_wr[l]); p.bias = nullptr; p.out_sums = nullptr;
            __half* out = l == 0 ? H16(d.off_dx0) : H16(d.off_da[l & 1]);
            p.dst = out; p.cout_valid = l == 0 ? 3 : 64;
            if (l == L) rc = launch_conv<16, 8, 64, IN_DZ>(p, st);
            else if (l == 0) rc = launch_conv<64, 64, 16, IN_DZ>(p, st);
            else rc = launch_conv<64, 64, 64, IN_DZ>(p, st);
            if (rc) return rc;
            da = out;
        }
    }
    for (int l = 0; l <= L; ++l) {   // dbeta = sum dy, dgamma = sum dy*yhat: the backward statistics themselves
        k_dec_bn_grads<<<1, 64, 0, st>>>(D64(d.off_bsums[l]), l == L ? 8 : 64, l == L ? nb : 64, grads[4 * l + 2], grads[4 * l + 3]);
        NSIG_LAUNCH_CHECK();
    }
    if (dimage) {
        k_dec_input_grad<<<(unsigned)((d.n_pix + 255) / 256), 256, 0, st>>>(H16(d.off_dx0), n_pix, dimage);
        NSIG_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
