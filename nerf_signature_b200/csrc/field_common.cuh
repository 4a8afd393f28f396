// field_common.cuh — pieces shared by the fused field kernels (field.cu) and the occupancy-grid sweep
// (grid.cu): shared-memory weight layout, m16n8k16 layer helpers, fragment-layout hash encoding, SH4,
// the FieldParams block and its host-side validation.
#pragma once
#include "hash_common.cuh"
#include "march_common.cuh"

#include <cstdlib>

namespace nsig {

// ---- shared-memory weight layout (halfs); row strides padded by 8 halfs: conflict-free B loads
constexpr int kS32 = 40;  // stride of a [*,32] matrix
constexpr int kS64 = 72;  // stride of a [*,64] matrix
constexpr int kS16 = 24;  // stride of a [*,16] matrix
// forward copies, [out][in]
constexpr int oWs0 = 0;                    // [64][32]
constexpr int oWs1 = oWs0 + 64 * kS32;     // [16][64] rows permuted: r' <- (r'+1)%16
constexpr int oWc0 = oWs1 + 16 * kS64;     // [64][32] (input column 31 zeroed)
constexpr int oWc1 = oWc0 + 64 * kS32;     // [64][64]
constexpr int oWc2 = oWc1 + 64 * kS64;     // [8][64]  (outputs 0..7; 3..7 are padding)
constexpr int kFwdHalfs = oWc2 + 8 * kS64;
constexpr int kFwdHalfsPad = (kFwdHalfs + 7) / 8 * 8;  // 16-byte aligned end of the forward weights
// transposed copies for dgrad, [in][out]
constexpr int oWc2T = kFwdHalfs;           // [64][16] (outputs >= 3 zeroed)
constexpr int oWc1T = oWc2T + 64 * kS16;   // [64][64]
constexpr int oWc0T = oWc1T + 64 * kS64;   // [16][64] inputs 16..31 (geo part)
constexpr int oWs1T = oWc0T + 16 * kS64;   // [64][16] outputs permuted like oWs1
constexpr int oWs0T = oWs1T + 64 * kS16;   // [32][64]
constexpr int kBwdHalfs = oWs0T + 32 * kS64;
constexpr int kBwdHalfsPad = (kBwdHalfs + 7) / 8 * 8;

constexpr int kFieldWarps = 4;
constexpr int kFieldThreads = kFieldWarps * 32;
#ifndef NSIG_FWD_MINB
#define NSIG_FWD_MINB 4  // resident CTAs per SM the forward kernel is compiled for (register cap 65536/(128*MINB))
#endif

struct FieldParams {
    const float* xyzs;
    const float* dirs;
    uint32_t M;
    float bound_add;   // bound
    float bound_mul;   // fl(1 / (2*bound)): torch divides by a scalar as x * (1/s) on CUDA
    FusedTablePtrs base;  // 16 levels
    const float2* S;      // pre-summed message table or null
    LevelGeom msg_geom;
    uint32_t mask;
    const __half* sigma_w;
    const __half* color_w;
    const int32_t* M_dev;  // optional device-side sample count (march counter): M = min(M, *M_dev)
    float density_scale;   // sigma = density_scale * exp(logit)  (renderer_wtmk.py:294)
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_relu_h2(float a, float b) {
    return pack_h2(fmaxf(a, 0.0f), fmaxf(b, 0.0f));
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// C[MT][NT] (+)= A[MT][KS] x W^T, W in shared memory as [n][k] with `stride` halfs per row.
// NT0 = first n-tile computed (lets the caller skip unused output columns).
template <int MT, int KS, int NT, int NT0 = 0>
__device__ __forceinline__ void layer(float (&c)[MT][NT][4], const uint32_t (&a)[MT][KS][4],
                                      const __half* __restrict__ W, int stride, int g, int tig) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { c[mt][nt][0] = c[mt][nt][1] = c[mt][nt][2] = c[mt][nt][3] = 0.f; }
        const __half* wrow = W + ((NT0 + nt) * 8 + g) * stride + 2 * tig;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wrow + ks * 16);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wrow + ks * 16 + 8);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma16816(c[mt][nt], a[mt][ks], b0, b1);
        }
    }
}

// accumulators of a 16 x (NT*8) layer output -> A fragments of the next layer (ReLU, fp16)
template <int MT, int NT>
__device__ __forceinline__ void relu_to_a(uint32_t (&a)[MT][NT / 2][4], const float (&c)[MT][NT][4]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            a[mt][ks][0] = pack_relu_h2(c[mt][2 * ks][0], c[mt][2 * ks][1]);
            a[mt][ks][1] = pack_relu_h2(c[mt][2 * ks][2], c[mt][2 * ks][3]);
            a[mt][ks][2] = pack_relu_h2(c[mt][2 * ks + 1][0], c[mt][2 * ks + 1][1]);
            a[mt][ks][3] = pack_relu_h2(c[mt][2 * ks + 1][2], c[mt][2 * ks + 1][3]);
        }
}

// Sign masks of a 16 x 64 hidden layer, for the backward pass (k_field_bwd_masks).  Taken from the layer's A fragments
// (post-ReLU, fp16-rounded - exactly the values whose sign k_field_bwd re-derives): one HSET2 + one LOP3 per fragment
// register.  Bit layout of a row's 32-bit word: fragment register (ks, q) = n-tile nt = 2*ks + q holds hidden units
// nt*8 + 2*tig (low half) and nt*8 + 2*tig + 1 (high half); the low half's sign goes to bit nt + shift, the high half's to
// bit 16 + nt + shift.  shift = 0 / 8 lets two layers share one word.
template <int MT>
__device__ __forceinline__ void act_mask_bits(uint32_t (&bits)[MT][2], const uint32_t (&a)[MT][4][4], int shift) {
    const __half2 zero = __float2half2_rn(0.0f);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t b = bits[mt][h];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t m = __hgt2_mask(*reinterpret_cast<const __half2*>(&a[mt][ks][h + 2 * q]), zero);  // 0xffff per positive half
                    b |= m & (0x00010001u << (2 * ks + q + shift));
                }
            bits[mt][h] = b;
        }
}

// gradient accumulators -> A fragments, masked by the sign bits the forward kernel saved (act_mask_bits): the selected
// halves are kept, the others become +0 (a bitwise AND: inf/NaN in a masked-out unit cannot leak, exactly like a select)
template <int MT, int NT>
__device__ __forceinline__ void grad_to_a_bits(uint32_t (&a)[MT][NT / 2][4], const float (&c)[MT][NT][4],
                                               const uint32_t (&bits)[MT][2], int shift) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks)
#pragma unroll
            for (int r = 0; r < 4; ++r) {   // fragment register r: row half h = r & 1, n-tile 2*ks + (r >> 1)
                const int h = r & 1, nt = 2 * ks + (r >> 1);
                const uint32_t sel = (bits[mt][h] >> (nt + shift)) & 0x00010001u;
                a[mt][ks][r] = pack_h2(c[mt][nt][2 * h], c[mt][nt][2 * h + 1]) & (sel * 0xffffu);
            }
}

// gradient accumulators -> A fragments, masked by the forward activation (ReLU'(h) = h > 0)
__device__ __forceinline__ uint32_t pack_masked(float a, float b, uint32_t act) {
    const __half2 h = *reinterpret_cast<const __half2*>(&act);
    const __half2 m = __hgt2(h, __float2half2_rn(0.0f));  // 1.0 where h > 0
    __half2 v = __hmul2(__floats2half2_rn(a, b), m);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <int MT, int NT>
__device__ __forceinline__ void grad_to_a(uint32_t (&a)[MT][NT / 2][4], const float (&c)[MT][NT][4],
                                          const uint32_t (&act)[MT][NT / 2][4]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
            a[mt][ks][0] = pack_masked(c[mt][2 * ks][0], c[mt][2 * ks][1], act[mt][ks][0]);
            a[mt][ks][1] = pack_masked(c[mt][2 * ks][2], c[mt][2 * ks][3], act[mt][ks][1]);
            a[mt][ks][2] = pack_masked(c[mt][2 * ks + 1][0], c[mt][2 * ks + 1][1], act[mt][ks][2]);
            a[mt][ks][3] = pack_masked(c[mt][2 * ks + 1][2], c[mt][2 * ks + 1][3], act[mt][ks][3]);
        }
}

// ---- weight staging ---------------------------------------------------------------------
// Each CTA copies the (frozen) MLP weights into its padded shared-memory layout once.  The copies move 16 bytes
// (8 halfs) per load: the 2-byte-per-thread version spent ~7 % of the backward kernel's warp time here (ncu source view,
// profiles/r01_experiments_v5.txt).  Row lengths (32 / 64 halfs) and padded strides (80 / 144 / 48 bytes) are multiples
// of 16 bytes, so a chunk never straddles a row; the flat fp16 parameter vectors must be 16-byte aligned (the entry
// points return NSIG_EINVAL otherwise; torch allocations are 256-byte aligned).
__device__ __forceinline__ uint4 ldg16(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void sts16(__half* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ uint4 zero_last_half(uint4 v) { v.w &= 0x0000ffffu; return v; }   // halfs 0..6 kept, half 7 = 0

__device__ __forceinline__ void stage_forward_weights(__half* sm, const __half* __restrict__ sw,
                                                      const __half* __restrict__ cw, bool color) {
    for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) {          // [64][32]: 4 chunks per row
        const int r = i >> 2, c = (i & 3) * 8;
        sts16(sm + oWs0 + r * kS32 + c, ldg16(sw + r * 32 + c));
        if (color) {                                                    // input column 31 is padding: zeroed
            const uint4 v = ldg16(cw + r * 32 + c);
            sts16(sm + oWc0 + r * kS32 + c, c == 24 ? zero_last_half(v) : v);
        }
    }
    for (int i = threadIdx.x; i < 16 * 8; i += blockDim.x) {          // [16][64]: smem row r holds param row (r+1)%16
        const int r = i >> 3, c = (i & 7) * 8;                          //           = [geo0..14, logit]
        sts16(sm + oWs1 + r * kS64 + c, ldg16(sw + 2048 + ((r + 1) & 15) * 64 + c));
    }
    if (color) {
        for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
            const int r = i >> 3, c = (i & 7) * 8;
            sts16(sm + oWc1 + r * kS64 + c, ldg16(cw + 2048 + r * 64 + c));
        }
        for (int i = threadIdx.x; i < 8 * 8; i += blockDim.x) {
            const int r = i >> 3, c = (i & 7) * 8;
            sts16(sm + oWc2 + r * kS64 + c, ldg16(cw + 2048 + 4096 + r * 64 + c));
        }
    }
}

// transposed copies for the data gradients: 8 consecutive source elements (one 16-byte load) go to 8 consecutive
// shared-memory ROWS of one column
__device__ __forceinline__ void scatter8(__half* dst, int stride, uint4 v) {
    const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j * stride] = h[j];
}

__device__ __forceinline__ void stage_backward_weights(__half* sm, const __half* __restrict__ sw,
                                                       const __half* __restrict__ cw) {
    const __half zero = __float2half(0.0f);
    // oWc2T [64][16]: column k < 3 = colour output k (source row k of [16][64]), other columns zero
    // oWs1T [64][16]: column k = permuted sigma output k (source row (k+1)%16 of [16][64])
    for (int i = threadIdx.x; i < 16 * 8; i += blockDim.x) {
        const int k = i >> 3, n = (i & 7) * 8;                          // source row k, columns n..n+7
        if (k < 3) scatter8(sm + oWc2T + n * kS16 + k, kS16, ldg16(cw + 2048 + 4096 + k * 64 + n));
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) sm[oWc2T + (n + j) * kS16 + k] = zero;
        }
        scatter8(sm + oWs1T + n * kS16 + k, kS16, ldg16(sw + 2048 + ((k + 1) & 15) * 64 + n));
    }
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {          // oWc1T[n][k] = Wc1[k][n]
        const int k = i >> 3, n = (i & 7) * 8;
        scatter8(sm + oWc1T + n * kS64 + k, kS64, ldg16(cw + 2048 + k * 64 + n));
    }
    for (int i = threadIdx.x; i < 64 * 2; i += blockDim.x) {          // oWc0T[n][k] = Wc0[k][16 + n], n < 16 (n == 15: padding)
        const int k = i >> 1, n = (i & 1) * 8;
        uint4 v = ldg16(cw + k * 32 + 16 + n);
        if (n == 8) v = zero_last_half(v);
        scatter8(sm + oWc0T + n * kS64 + k, kS64, v);
    }
    for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) {          // oWs0T[n][k] = Ws0[k][n], n < 32
        const int k = i >> 2, n = (i & 3) * 8;
        scatter8(sm + oWs0T + n * kS64 + k, kS64, ldg16(sw + k * 32 + n));
    }
}

// ---- SH degree 4 at v = ((d+1)/2)*2-1 (network_wtmk_tcnn.py:114 + tcnn's [0,1] convention);
//      formulas of hash_encoding.py:162-193 -----------------------------------------------------
__device__ __forceinline__ void sh4(float dx, float dy, float dz, float (&o)[16]) {
    const float x = __fmaf_rn(__fmul_rn(__fadd_rn(dx, 1.0f), 0.5f), 2.0f, -1.0f);
    const float y = __fmaf_rn(__fmul_rn(__fadd_rn(dy, 1.0f), 0.5f), 2.0f, -1.0f);
    const float z = __fmaf_rn(__fmul_rn(__fadd_rn(dz, 1.0f), 0.5f), 2.0f, -1.0f);
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.4886025119029199f * y;
    o[2] = 0.4886025119029199f * z;
    o[3] = -0.4886025119029199f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.5462742152960396f * (xx - yy);
    o[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
    o[10] = 2.890611442640554f * xy * z;
    o[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    o[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    o[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    o[14] = 1.445305721320277f * z * (xx - yy);
    o[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// row index helpers: a warp owns rows [row0, row0 + 16*MT); thread (g,tig) owns rows
// row0 + mt*16 + h*8 + g for mt < MT, h in {0,1}; fragment register index is 2*half_k + h.
template <int MT>
__device__ __forceinline__ uint32_t frag_row(uint32_t row0, int mt, int h, int g) {
    return row0 + mt * 16 + h * 8 + g;
}

// Encode the rows of this thread: A fragments of the sigma net's first layer.
// fa[mt][ks][2*hk + h]: level = 8*ks + 4*hk + tig, row half h.
// xn[mt][h][a]: the rows' positions already normalised to the unit box.
// H2: gather the half2 shadow tables (p.base.th / inv_scale) instead of the fp32 ones.
template <int MT, bool H2 = false>
__device__ __forceinline__ void encode_positions(uint32_t (&fa)[MT][2][4], const FieldParams& p,
                                                 const float (&xn)[MT][2][3], int g, int tig) {
    float2 f[MT][2][4];  // [mt][h][j]: level tig + 4j
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int level = tig + 4 * j;
        const float2* tab = p.base.t[level];
        const __half2* tabh = p.base.th[level];
        const float inv_scale = H2 ? __ldg(p.base.inv_scale + level) : 1.0f;
        const LevelGeom L = p.base.geom[level];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const Voxel v = locate_fused(xn[mt][h][0], xn[mt][h][1], xn[mt][h][2], L);
                f[mt][h][j] = H2 ? encode_level_fused_h2(tabh, v, p.mask, inv_scale) : encode_level_fused(tab, v, p.mask);
            }
    }
    if (p.S != nullptr) {
        // message feature (hash_encoding_wtmk_bit.py, pre-summed form): quad thread `tig` evaluates row
        // (mt = tig>>1, h = tig&1) of each 2-tile group, thread tig==3 (owner of channels 30,31) collects.
#pragma unroll
        for (int q = 0; q < (MT * 2 + 3) / 4; ++q) {
            const int sel = q * 4 + tig;  // which (mt,h) this thread evaluates
            float2 mine = make_float2(0.f, 0.f);
            float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (mt * 2 + h == sel) { sx = xn[mt][h][0]; sy = xn[mt][h][1]; sz = xn[mt][h][2]; }
            if (sel < MT * 2) {
                const Voxel v = locate_fused(sx, sy, sz, p.msg_geom);
                mine = encode_level_fused(p.S, v, p.mask);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int src = mt * 2 + h - q * 4;  // quad lane that evaluated this row
                    if (src >= 0 && src < 4) {
                        const float mx = __shfl_sync(NSIG_FULL_MASK, mine.x, (g << 2) | src);
                        const float my = __shfl_sync(NSIG_FULL_MASK, mine.y, (g << 2) | src);
                        if (tig == 3) {  // x_feature[:, -2:] += msg_feature (network_wtmk_tcnn.py:106)
                            f[mt][h][3].x = __fadd_rn(f[mt][h][3].x, mx);
                            f[mt][h][3].y = __fadd_rn(f[mt][h][3].y, my);
                        }
                    }
                }
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j)  // level tig+4j -> k-step j>>1, half-k j&1
                fa[mt][j >> 1][2 * (j & 1) + h] = pack_h2(f[mt][h][j].x, f[mt][h][j].y);
}

// The same encode with the loop over the thread's 2*MT rows NOT unrolled: one copy of the per-row code (4 base levels + locate)
// instead of 2*MT.  For kernels whose warps drift apart (k_render_rays: every warp marches its own ray, so the warps of an SM
// partition sit in march, encode, MLP and composite code at the same time) the fully unrolled body does not fit the
// instruction caches: ncu showed stall_no_instruction as the top stall (28 % of the samples, profiles/r02_ncu_render.txt).
// pos(mt, h, x): the row's normalised position (a runtime (mt, h), e.g. read from shared or global memory).
template <int MT, bool H2, typename PosFn>
__device__ __forceinline__ void encode_positions_rolled(uint32_t (&fa)[MT][2][4], const FieldParams& p, PosFn pos, int g, int tig) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int r = 0; r < 4; ++r) fa[mt][ks][r] = 0u;
    float2 top[MT][2];   // level tig + 12 of every row, kept in fp32 until the message feature has been added (tig == 3)
    float xnext[3];
    pos(0, 0, xnext);
#pragma unroll 1
    for (int s = 0; s < 2 * MT; ++s) {
        // this row's position was requested one iteration ago; request the next row's now (its latency would otherwise be
        // exposed at the head of every iteration: 11 % of the warp-stall samples sat on the first use of the position)
        const float x[3] = {xnext[0], xnext[1], xnext[2]};
        if (s + 1 < 2 * MT) pos((s + 1) >> 1, (s + 1) & 1, xnext);
        uint32_t packed[3];
        float2 last = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int level = tig + 4 * j;
            const LevelGeom L = p.base.geom[level];
            const Voxel v = locate_fused(x[0], x[1], x[2], L);
            const float2 f = H2 ? encode_level_fused_h2(p.base.th[level], v, p.mask, __ldg(p.base.inv_scale + level))
                                : encode_level_fused(p.base.t[level], v, p.mask);
            if (j < 3) packed[j] = pack_h2(f.x, f.y); else last = f;
        }
#pragma unroll
        for (int m2 = 0; m2 < MT; ++m2)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
                if (s == m2 * 2 + h2) {   // level tig+4j -> k-step j>>1, register 2*(j&1)+h
                    fa[m2][0][h2] = packed[0]; fa[m2][0][2 + h2] = packed[1]; fa[m2][1][h2] = packed[2];
                    top[m2][h2] = last;
                }
    }
    if (p.S != nullptr) {   // message feature, as in encode_positions
#pragma unroll
        for (int q = 0; q < (MT * 2 + 3) / 4; ++q) {
            const int sel = q * 4 + tig;
            float2 mine = make_float2(0.f, 0.f);
            if (sel < MT * 2) {
                float x[3];
                pos(sel >> 1, sel & 1, x);
                const Voxel v = locate_fused(x[0], x[1], x[2], p.msg_geom);
                mine = encode_level_fused(p.S, v, p.mask);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int src = mt * 2 + h - q * 4;
                    if (src >= 0 && src < 4) {
                        const float mx = __shfl_sync(NSIG_FULL_MASK, mine.x, (g << 2) | src);
                        const float my = __shfl_sync(NSIG_FULL_MASK, mine.y, (g << 2) | src);
                        if (tig == 3) {
                            top[mt][h].x = __fadd_rn(top[mt][h].x, mx);
                            top[mt][h].y = __fadd_rn(top[mt][h].y, my);
                        }
                    }
                }
        }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) fa[mt][1][2 + h] = pack_h2(top[mt][h].x, top[mt][h].y);
}

template <int MT, bool H2 = false>
__device__ __forceinline__ void encode_rows(uint32_t (&fa)[MT][2][4], const FieldParams& p, uint32_t M,
                                            uint32_t row0, int g, int tig) {
    float xn[MT][2][3];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t r = min(frag_row<MT>(row0, mt, h, g), M - 1);
#pragma unroll
            for (int a = 0; a < 3; ++a)  // x = (x + bound) / (2*bound)  (network_wtmk_tcnn.py:101)
                xn[mt][h][a] = __fmul_rn(__fadd_rn(__ldg(p.xyzs + (size_t)r * 3 + a), p.bound_add), p.bound_mul);
        }
    encode_positions<MT, H2>(fa, p, xn, g, tig);
}

// encode_rows with the per-row code emitted once (encode_positions_rolled): positions are re-read per row
template <int MT, bool H2 = false>
__device__ __forceinline__ void encode_rows_rolled(uint32_t (&fa)[MT][2][4], const FieldParams& p, uint32_t M,
                                                   uint32_t row0, int g, int tig) {
    auto pos = [&](int mt, int h, float (&x)[3]) {
        const uint32_t r = min(row0 + (uint32_t)(mt * 16 + h * 8 + g), M - 1);
#pragma unroll
        for (int a = 0; a < 3; ++a) x[a] = __fmul_rn(__fadd_rn(__ldg(p.xyzs + (size_t)r * 3 + a), p.bound_add), p.bound_mul);
    };
    encode_positions_rolled<MT, H2>(fa, p, pos, g, tig);
}

// ---------------------------------------------------------------------------------------------------
// Warp-cooperative gather (v2).  ncu shows the v1 gather above bound by the L1 tag stage: one 128-byte line
// per cycle and SM, 87 % busy (profiles/r01_field_kernels_v1.txt: 31.0 M tag requests in 241.8 k cycles x 148 SMs),
// so the lever is the number of DISTINCT LINES per load instruction, not bytes.  Here a load instruction covers
// 16 consecutive rows x the two x-neighbours of one (y,z) corner pair of ONE level:
//   * the hash is x ^ y*P1 ^ z*P2, so for an even x index the two x-neighbours are adjacent table entries - the
//     lane pair (xbit = 0,1) then hits the same line: ~6 instead of 8 lines per row and level;
//   * 16 rows of a ray (instead of 8 rows x 4 levels) share voxels - hence lines - on the coarse levels.
// Each lane accumulates its 4 weighted corners, one xor-shuffle adds the x-neighbour, the even lane writes the
// feature (fp16 pair) into a per-warp shared-memory tile [32 rows][16 levels] and ldmatrix turns the tile into the
// m16n8k16 A fragments.  Interpolation is a plain weighted sum of the 8 corners (association differs from the
// reference's x,y,z lerp order by fp32 rounding only; the result is rounded to fp16 for the MMA anyway).
// ---------------------------------------------------------------------------------------------------
constexpr int kFeatStride = 20;                    // 32-bit words per row of the feature tile (64 B + 16 B pad)
constexpr int kFeatWords = 32 * kFeatStride;       // per warp
// dynamic shared memory of the 4-warp field kernels: forward weights + one feature tile per warp
#ifdef NSIG_GATHER_V2
constexpr int kFeatWordsAlloc = kFeatWords;
#else
constexpr int kFeatWordsAlloc = 0;                 // the default (v1) gather goes straight to fragments
#endif
constexpr size_t kFieldFwdSmem = kFwdHalfsPad * sizeof(__half) + (size_t)kFieldWarps * kFeatWordsAlloc * sizeof(uint32_t);

template <bool H2>
__device__ __forceinline__ float2 gather_half_voxel(const FieldParams& p, int level, float x, float y, float z,
                                                    int xbit) {
    const LevelGeom L = p.base.geom[level];
    uint32_t ix, iy, iz;
    float wx, wy, wz;
    locate_axis_fused(x, L, ix, wx);
    locate_axis_fused(y, L, iy, wy);
    locate_axis_fused(z, L, iz, wz);
    const uint32_t hx = ix + (uint32_t)xbit;
    const float fx = xbit ? wx : 1.0f - wx;
    const uint32_t hy0 = iy * kPrimeY, hy1 = hy0 + kPrimeY, hz0 = iz * kPrimeZ, hz1 = hz0 + kPrimeZ;
    const float fy0 = fx * (1.0f - wy), fy1 = fx * wy, oz = 1.0f - wz;
    const uint32_t s00 = (hx ^ hy0 ^ hz0) & p.mask, s01 = (hx ^ hy0 ^ hz1) & p.mask;
    const uint32_t s10 = (hx ^ hy1 ^ hz0) & p.mask, s11 = (hx ^ hy1 ^ hz1) & p.mask;
    float2 e00, e01, e10, e11;
    if (H2) {
        const uint32_t* t = reinterpret_cast<const uint32_t*>(p.base.th[level]);
        const uint32_t r00 = __ldg(t + s00), r01 = __ldg(t + s01), r10 = __ldg(t + s10), r11 = __ldg(t + s11);
        e00 = __half22float2(*reinterpret_cast<const __half2*>(&r00));
        e01 = __half22float2(*reinterpret_cast<const __half2*>(&r01));
        e10 = __half22float2(*reinterpret_cast<const __half2*>(&r10));
        e11 = __half22float2(*reinterpret_cast<const __half2*>(&r11));
    } else {
        const float2* t = p.base.t[level];
        e00 = __ldg(t + s00); e01 = __ldg(t + s01); e10 = __ldg(t + s10); e11 = __ldg(t + s11);
    }
    const float w00 = fy0 * oz, w01 = fy0 * wz, w10 = fy1 * oz, w11 = fy1 * wz;
    float2 a;
    a.x = fmaf(w11, e11.x, fmaf(w10, e10.x, fmaf(w01, e01.x, w00 * e00.x)));
    a.y = fmaf(w11, e11.y, fmaf(w10, e10.y, fmaf(w01, e01.y, w00 * e00.y)));
    if (H2) { const float inv = __ldg(p.base.inv_scale + level); a.x *= inv; a.y *= inv; }
    return a;
}

// message feature of one row half (pre-summed table S, fp32): same lane-pair scheme
__device__ __forceinline__ float2 gather_half_voxel_msg(const FieldParams& p, float x, float y, float z, int xbit) {
    uint32_t ix, iy, iz;
    float wx, wy, wz;
    locate_axis_fused(x, p.msg_geom, ix, wx);
    locate_axis_fused(y, p.msg_geom, iy, wy);
    locate_axis_fused(z, p.msg_geom, iz, wz);
    const uint32_t hx = ix + (uint32_t)xbit;
    const float fx = xbit ? wx : 1.0f - wx;
    const uint32_t hy0 = iy * kPrimeY, hy1 = hy0 + kPrimeY, hz0 = iz * kPrimeZ, hz1 = hz0 + kPrimeZ;
    const float fy0 = fx * (1.0f - wy), fy1 = fx * wy, oz = 1.0f - wz;
    const float2 e00 = __ldg(p.S + ((hx ^ hy0 ^ hz0) & p.mask)), e01 = __ldg(p.S + ((hx ^ hy0 ^ hz1) & p.mask));
    const float2 e10 = __ldg(p.S + ((hx ^ hy1 ^ hz0) & p.mask)), e11 = __ldg(p.S + ((hx ^ hy1 ^ hz1) & p.mask));
    const float w00 = fy0 * oz, w01 = fy0 * wz, w10 = fy1 * oz, w11 = fy1 * wz;
    float2 a;
    a.x = fmaf(w11, e11.x, fmaf(w10, e10.x, fmaf(w01, e01.x, w00 * e00.x)));
    a.y = fmaf(w11, e11.y, fmaf(w10, e10.y, fmaf(w01, e01.y, w00 * e00.y)));
    return a;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const uint32_t* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// (px,py,pz): normalised position of tile row `lane`.  feat_s: this warp's [32][kFeatStride] tile.
// On return fa holds the A fragments of both 16-row tiles and feat_s the fp16 features (row-major, 16 half2 per row).
template <bool H2>
__device__ __forceinline__ void gather_tile(uint32_t (&fa)[2][2][4], const FieldParams& p, float px, float py, float pz,
                                            uint32_t* __restrict__ feat_s, int lane) {
    const int xbit = lane & 1, r16 = lane >> 1;
    __syncwarp();  // the previous tile's readers are done with feat_s
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
        const int row = grp * 16 + r16;
        const float x = __shfl_sync(NSIG_FULL_MASK, px, row), y = __shfl_sync(NSIG_FULL_MASK, py, row),
                    z = __shfl_sync(NSIG_FULL_MASK, pz, row);
#pragma unroll 4
        for (int level = 0; level < NSIG_MAX_LEVELS; ++level) {
            float2 a = gather_half_voxel<H2>(p, level, x, y, z, xbit);
            if (level == NSIG_MAX_LEVELS - 1 && p.S != nullptr) {  // x_feature[:, -2:] += msg_feature (network_wtmk_tcnn.py:106)
                const float2 m = gather_half_voxel_msg(p, x, y, z, xbit);
                a.x += m.x; a.y += m.y;
            }
            a.x += __shfl_xor_sync(NSIG_FULL_MASK, a.x, 1);
            a.y += __shfl_xor_sync(NSIG_FULL_MASK, a.y, 1);
            if (xbit == 0) feat_s[row * kFeatStride + level] = pack_h2(a.x, a.y);
        }
    }
    __syncwarp();
    // A fragment of tile mt, k-step ks: lanes 0-15 address rows 0-15 at halfs [16ks, 16ks+8), lanes 16-31 the next 8
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            ldmatrix_x4(fa[mt][ks], feat_s + (mt * 16 + (lane & 15)) * kFeatStride + ks * 8 + (lane >> 4) * 4);
}

// A fragments of the colour net's first k-step: SH(d) columns {2tig,2tig+1,2tig+8,2tig+9}
template <int MT>
__device__ __forceinline__ void sh_rows(uint32_t (&ca)[MT][2][4], const float* __restrict__ dirs, uint32_t M,
                                        uint32_t row0, int g, int tig) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t r = min(frag_row<MT>(row0, mt, h, g), M - 1);
            float o[16];
            sh4(__ldg(dirs + (size_t)r * 3), __ldg(dirs + (size_t)r * 3 + 1), __ldg(dirs + (size_t)r * 3 + 2), o);
            float lo0 = o[0], lo1 = o[1], hi0 = o[8], hi1 = o[9];
#pragma unroll
            for (int t = 1; t < 4; ++t)
                if (tig == t) { lo0 = o[2 * t]; lo1 = o[2 * t + 1]; hi0 = o[2 * t + 8]; hi1 = o[2 * t + 9]; }
            ca[mt][0][h] = pack_h2(lo0, lo1);
            ca[mt][0][2 + h] = pack_h2(hi0, hi1);
        }
}

// same, from directions already in registers (dv[mt][h] = direction of row (mt,h))
template <int MT>
__device__ __forceinline__ void sh_rows_vals(uint32_t (&ca)[MT][2][4], const float (&dv)[MT][2][3], int tig) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float o[16];
            sh4(dv[mt][h][0], dv[mt][h][1], dv[mt][h][2], o);
            float lo0 = o[0], lo1 = o[1], hi0 = o[8], hi1 = o[9];
#pragma unroll
            for (int t = 1; t < 4; ++t)
                if (tig == t) { lo0 = o[2 * t]; lo1 = o[2 * t + 1]; hi0 = o[2 * t + 8]; hi1 = o[2 * t + 9]; }
            ca[mt][0][h] = pack_h2(lo0, lo1);
            ca[mt][0][2 + h] = pack_h2(hi0, hi1);
        }
}

// geo features (sigma net outputs in the permuted order [geo0..14, logit]) -> colour k-step 1;
// column 15 (the logit; the colour net's padded input 31) is cleared.
template <int MT>
__device__ __forceinline__ void geo_to_a(uint32_t (&ca)[MT][2][4], const float (&so)[MT][2][4], int tig) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        ca[mt][1][0] = pack_h2(so[mt][0][0], so[mt][0][1]);
        ca[mt][1][1] = pack_h2(so[mt][0][2], so[mt][0][3]);
        ca[mt][1][2] = pack_h2(so[mt][1][0], (tig == 3) ? 0.0f : so[mt][1][1]);
        ca[mt][1][3] = pack_h2(so[mt][1][2], (tig == 3) ? 0.0f : so[mt][1][3]);
    }
}
}  // namespace nsig

static inline int fill_field_params(nsig::FieldParams& p, const float* xyzs, const float* dirs, uint32_t M, float bound,
                             const float* const* tables, const float* resolutions, uint32_t log2_T,
                             const float* S, float msg_resolution, const void* sigma_w, const void* color_w,
                             const int32_t* M_dev, float density_scale, const void* const* tables_h2 = nullptr,
                             const float* h2_inv_scale = nullptr) {
    if (!xyzs || !tables || !resolutions || !sigma_w) return NSIG_EINVAL;
    if ((((uintptr_t)sigma_w) | ((uintptr_t)color_w)) & 15) return NSIG_EINVAL;   // 16-byte weight staging
    if ((tables_h2 == nullptr) != (h2_inv_scale == nullptr)) return NSIG_EINVAL;
    if (log2_T < 1 || log2_T > 30 || !(bound > 0.0f)) return NSIG_EINVAL;
    p.xyzs = xyzs;
    p.dirs = dirs;
    p.M = M;
    p.bound_add = bound;
    p.bound_mul = 1.0f / (2.0f * bound);
    for (int l = 0; l < NSIG_MAX_LEVELS; ++l) {
        if (!tables[l] || !(resolutions[l] > 0.0f)) return NSIG_EINVAL;
        p.base.t[l] = reinterpret_cast<const float2*>(tables[l]);
        p.base.geom[l] = nsig::make_level_geom(resolutions[l]);
        p.base.th[l] = tables_h2 ? reinterpret_cast<const __half2*>(tables_h2[l]) : nullptr;
        if (tables_h2 && !tables_h2[l]) return NSIG_EINVAL;
    }
    p.base.inv_scale = h2_inv_scale;
    p.S = reinterpret_cast<const float2*>(S);
    p.msg_geom = nsig::make_level_geom((msg_resolution > 0.0f) ? msg_resolution : 1.0f);
    if (S && !(msg_resolution > 0.0f)) return NSIG_EINVAL;
    p.mask = (1u << log2_T) - 1u;
    p.sigma_w = reinterpret_cast<const __half*>(sigma_w);
    p.color_w = reinterpret_cast<const __half*>(color_w);
    p.M_dev = M_dev;
    p.density_scale = density_scale;
    return 0;
}

// persistent grid: exactly one resident wave (occupancy queried from the runtime), tiles handed out grid-stride
template <typename K>
static inline int field_grid(K kernel, size_t smem, uint32_t M, uint32_t rows_per_cta) {
    int ctas_per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, nsig::kFieldThreads, smem) != cudaSuccess ||
        ctas_per_sm < 1)
        ctas_per_sm = 1;
    // experiment switch: cap the resident CTAs per SM of the persistent field kernels so that kernels of a parallel graph
    // branch (the decoder chain in the harness' `overlap` render mode) find free registers / shared memory
    static const int cap_env = [] { const char* e = getenv("NSIG_FIELD_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    if (cap_env > 0 && ctas_per_sm > cap_env) ctas_per_sm = cap_env;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t tiles = nsig::div_up(M, rows_per_cta);
    const uint32_t cap = (uint32_t)(sms * ctas_per_sm);
    return (int)(tiles < cap ? tiles : cap);
}

