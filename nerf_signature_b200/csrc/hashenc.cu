// hashenc.cu — stand-alone multiresolution hash encoders (HashEmbedder.forward and its table
// gradient), the message-table pre-sum and the reference-form per-bit message encoder.
//
// These back the drop-in `HashEmbedder` modules.  The render/train hot path uses the fused
// field kernels (field.cu), which share the per-level arithmetic in hash_common.cuh.
#include "hash_common.cuh"

namespace nsig {

// ---------------------------------------------------------------------------------------
// forward: one thread per (sample, level); the 2*n_levels outputs of a sample are written by
// n_levels consecutive lanes => each warp stores 256 contiguous bytes.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_hash_encode_fwd(const float* __restrict__ x, uint32_t B, TablePtrs tp, uint32_t n_levels, uint32_t mask,
                  float* __restrict__ out, int32_t* __restrict__ slots) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)B * n_levels;
    if (gid >= total) return;
    const uint32_t b = (uint32_t)(gid / n_levels), l = (uint32_t)(gid % n_levels);
    const float px = x[(size_t)b * 3], py = x[(size_t)b * 3 + 1], pz = x[(size_t)b * 3 + 2];
    const Voxel v = locate(px, py, pz, tp.grid_size[l]);
    float2 e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t s = corner_slot(v, k, mask);
        e[k] = __ldg(tp.t[l] + s);
        if (slots) slots[gid * 8 + k] = (int32_t)s;
    }
    const float2 r = trilerp(e, v);
    *reinterpret_cast<float2*>(out + gid * 2) = r;
}

struct GradTablePtrs {
    float* t[NSIG_MAX_LEVELS];
    float grid_size[NSIG_MAX_LEVELS];
};

// backward: scatter-add into the tables, one thread per (sample, level), 8 vector reductions
__global__ void __launch_bounds__(256)
k_hash_encode_bwd(const float* __restrict__ x, const float* __restrict__ grad_out, uint32_t B,
                  GradTablePtrs tp, uint32_t n_levels, uint32_t mask) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)B * n_levels;
    if (gid >= total) return;
    const uint32_t b = (uint32_t)(gid / n_levels), l = (uint32_t)(gid % n_levels);
    const float2 g = *reinterpret_cast<const float2*>(grad_out + gid * 2);
    if (g.x == 0.0f && g.y == 0.0f) return;  // padding rows / terminated samples
    const float px = x[(size_t)b * 3], py = x[(size_t)b * 3 + 1], pz = x[(size_t)b * 3 + 2];
    const Voxel v = locate(px, py, pz, tp.grid_size[l]);
    float* tab = tp.t[l];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t s = corner_slot(v, k, mask);
        red_add_v2(tab + (size_t)s * 2, corner_grad(v, k, g.x), corner_grad(v, k, g.y));
    }
}

// ---------------------------------------------------------------------------------------
// S = sum_i tables[2i + bit_i]  (message read on device).  Pure streaming: message_dim x 4 MiB
// read once, 4 MiB written; loads bypass L1 and are marked evict-first in L2 so they do not
// push the base tables out.
// ---------------------------------------------------------------------------------------
struct MsgTablePtrs {
    const float* t[NSIG_MAX_MSG_TABLES];
};

struct f8 { float v[8]; };

// 256-bit streaming load (sm_100): no L1 allocation, evict-first in L2
__device__ __forceinline__ f8 ld_evict_first8(const float* p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]),
                   "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}

__global__ void __launch_bounds__(256)
k_msg_table_sum(MsgTablePtrs tp, uint32_t message_dim, const float* __restrict__ message, uint32_t n_vec8,
                float* __restrict__ S, uint32_t vec8_begin) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec8) return;
    const size_t off = (size_t)(vec8_begin + i) * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    uint32_t m = 0;
    for (; m + 4 <= message_dim; m += 4) {  // 4 independent 32-byte loads in flight
        f8 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t bit = (uint32_t)(int)__ldg(message + m + u);
            v[u] = ld_evict_first8(tp.t[2 * (m + u) + (bit & 1u)] + off);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[u].v[j];
    }
    for (; m < message_dim; ++m) {
        const uint32_t bit = (uint32_t)(int)__ldg(message + m);
        const f8 v = ld_evict_first8(tp.t[2 * m + (bit & 1u)] + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
    }
    float4* dst = reinterpret_cast<float4*>(S + off);
    dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// ---- half2 shadow tables for the fused kernels (hash_common.cuh: encode_level_fused_h2) ----------------------------
struct ShadowPtrs {
    const float* src[NSIG_MAX_LEVELS];
    __half2* dst[NSIG_MAX_LEVELS];
};

// pass 1: per-level max |v| (as the bit pattern of a non-negative float, so an integer atomicMax orders it)
__global__ void __launch_bounds__(256)
k_shadow_absmax(ShadowPtrs tp, uint32_t n_vec4, uint32_t* __restrict__ absmax_bits) {
    const uint32_t level = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(tp.src[level]);
    float m = 0.f;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n_vec4; i += gridDim.x * 256u) {
        const float4 v = ld_stream4(src + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(NSIG_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(absmax_bits + level, __float_as_uint(m));
}

// pass 2: dst = fp16(src * 2^k), k chosen so that max|v| * 2^k lies in [2^14, 2^15); inv_scale[level] = 2^-k
__global__ void __launch_bounds__(256)
k_shadow_convert(ShadowPtrs tp, uint32_t n_vec4, const uint32_t* __restrict__ absmax_bits, float* __restrict__ inv_scale) {
    const uint32_t level = blockIdx.y;
    const float m = __uint_as_float(absmax_bits[level]);
    int e = 0;
    float scale = 1.0f;
    if (m > 0.f) { frexpf(m, &e); scale = scalbnf(1.0f, max(-100, min(100, 15 - e))); }  // m = f * 2^e, f in [0.5, 1)
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[level] = 1.0f / scale;
    const float4* src = reinterpret_cast<const float4*>(tp.src[level]);
    uint2* dst = reinterpret_cast<uint2*>(tp.dst[level]);
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n_vec4; i += gridDim.x * 256u) {
        const float4 v = ld_stream4(src + i);
        const __half2 a = __floats2half2_rn(v.x * scale, v.y * scale), b = __floats2half2_rn(v.z * scale, v.w * scale);
        dst[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
}

// reference-form message encoder: per bit gather + trilerp, summed in bit order
__global__ void __launch_bounds__(256)
k_msg_encode_perbit(const float* __restrict__ x, uint32_t B, MsgTablePtrs tp, uint32_t message_dim,
                    const float* __restrict__ message, float grid_size, uint32_t mask, float* __restrict__ out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const Voxel v = locate(x[(size_t)b * 3], x[(size_t)b * 3 + 1], x[(size_t)b * 3 + 2], grid_size);
    uint32_t s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = corner_slot(v, k, mask);
    float a0 = 0.f, a1 = 0.f;
    for (uint32_t m = 0; m < message_dim; ++m) {
        const uint32_t bit = (uint32_t)(int)__ldg(message + m);
        const float2* tab = reinterpret_cast<const float2*>(tp.t[2 * m + (bit & 1u)]);
        float2 e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = __ldg(tab + s[k]);
        const float2 r = trilerp(e, v);
        a0 = __fadd_rn(a0, r.x);
        a1 = __fadd_rn(a1, r.y);
    }
    out[(size_t)b * 2] = a0;
    out[(size_t)b * 2 + 1] = a1;
}


// ---- parity probe of the FUSED kernels' per-level geometry ------------------------------------------------------------
// k_field_fwd / k_render_rays / k_grid_sweep locate a sample's voxel with locate_fused (hash_common.cuh: one double
// multiply instead of the reference's IEEE fp32 division).  This kernel evaluates exactly that device function and
// writes the 8 hashed slots (and the three interpolation weights) in the layout of k_hash_encode_fwd's `slots`, so a
// test can assert slot-for-slot equality with the reference-order encoder on adversarial inputs (cell boundaries +-k ulp).
struct GeomLevels { LevelGeom g[NSIG_MAX_LEVELS]; };

__global__ void __launch_bounds__(256)
k_fused_hash_slots(const float* __restrict__ x, uint32_t B, GeomLevels gl, uint32_t n_levels, uint32_t mask,
                   int32_t* __restrict__ slots, float* __restrict__ weights) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)B * n_levels;
    if (gid >= total) return;
    const uint32_t b = (uint32_t)(gid / n_levels), l = (uint32_t)(gid % n_levels);
    const Voxel v = locate_fused(x[(size_t)b * 3], x[(size_t)b * 3 + 1], x[(size_t)b * 3 + 2], gl.g[l]);
#pragma unroll
    for (int k = 0; k < 8; ++k) slots[gid * 8 + k] = (int32_t)corner_slot(v, k, mask);
    if (weights) { weights[gid * 3] = v.wx; weights[gid * 3 + 1] = v.wy; weights[gid * 3 + 2] = v.wz; }
}

}  // namespace nsig

using namespace nsig;

extern "C" {

int nsig_hash_encode_forward(const float* x, uint32_t B, const float* const* tables, const float* resolutions,
                             uint32_t n_levels, uint32_t log2_T, float* out, int32_t* slots,
                             nsig_stream_t stream) {
    if (B == 0) return 0;
    if (!x || !tables || !resolutions || !out) return NSIG_EINVAL;
    if (n_levels == 0 || n_levels > NSIG_MAX_LEVELS || log2_T == 0 || log2_T > 30) return NSIG_EINVAL;
    TablePtrs tp;
    for (uint32_t l = 0; l < n_levels; ++l) {
        if (!tables[l] || !(resolutions[l] > 0.0f)) return NSIG_EINVAL;
        tp.t[l] = reinterpret_cast<const float2*>(tables[l]);
        tp.grid_size[l] = 1.0f / resolutions[l];  // IEEE fp32 division, hash_encoding.py:37
    }
    const uint64_t total = (uint64_t)B * n_levels;
    k_hash_encode_fwd<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, B, tp, n_levels, (1u << log2_T) - 1u, out, slots);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_hash_encode_backward(const float* x, const float* grad_out, uint32_t B, float* const* grad_tables,
                              const float* resolutions, uint32_t n_levels, uint32_t log2_T,
                              nsig_stream_t stream) {
    if (B == 0) return 0;
    if (!x || !grad_out || !grad_tables || !resolutions) return NSIG_EINVAL;
    if (n_levels == 0 || n_levels > NSIG_MAX_LEVELS || log2_T == 0 || log2_T > 30) return NSIG_EINVAL;
    GradTablePtrs tp;
    for (uint32_t l = 0; l < n_levels; ++l) {
        if (!grad_tables[l] || !(resolutions[l] > 0.0f)) return NSIG_EINVAL;
        tp.t[l] = grad_tables[l];
        tp.grid_size[l] = 1.0f / resolutions[l];
    }
    const uint64_t total = (uint64_t)B * n_levels;
    k_hash_encode_bwd<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, grad_out, B, tp, n_levels, (1u << log2_T) - 1u);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_msg_table_sum(const float* const* tables, uint32_t message_dim, const float* message, uint32_t log2_T,
                       float* S, uint32_t elem_begin, uint32_t elem_count, nsig_stream_t stream) {
    if (!tables || !message || !S) return NSIG_EINVAL;
    if (message_dim == 0 || 2 * message_dim > NSIG_MAX_MSG_TABLES || log2_T < 1 || log2_T > 30) return NSIG_EINVAL;
    MsgTablePtrs tp;
    for (uint32_t i = 0; i < 2 * message_dim; ++i) {
        if (!tables[i] || (((uintptr_t)tables[i]) & 31)) return NSIG_EINVAL;
        tp.t[i] = tables[i];
    }
    if (log2_T < 2 || (((uintptr_t)S) & 31)) return NSIG_EINVAL;
    const uint32_t total = 2u << log2_T;         // T entries x 2 floats
    if (elem_count == 0) { elem_begin = 0; elem_count = total; }
    if ((elem_begin | elem_count) & 7u || elem_begin > total || elem_count > total - elem_begin) return NSIG_EINVAL;
    const uint32_t n_vec8 = elem_count / 8;
    k_msg_table_sum<<<div_up(n_vec8, 256), 256, 0, (cudaStream_t)stream>>>(tp, message_dim, message, n_vec8, S, elem_begin / 8);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_msg_encode_forward_perbit(const float* x, uint32_t B, const float* const* tables, uint32_t message_dim,
                                   const float* message, float resolution, uint32_t log2_T, float* out,
                                   nsig_stream_t stream) {
    if (B == 0) return 0;
    if (!x || !tables || !message || !out) return NSIG_EINVAL;
    if (message_dim == 0 || 2 * message_dim > NSIG_MAX_MSG_TABLES || log2_T < 1 || log2_T > 30) return NSIG_EINVAL;
    MsgTablePtrs tp;
    for (uint32_t i = 0; i < 2 * message_dim; ++i) {
        if (!tables[i]) return NSIG_EINVAL;
        tp.t[i] = tables[i];
    }
    k_msg_encode_perbit<<<div_up(B, 256), 256, 0, (cudaStream_t)stream>>>(
        x, B, tp, message_dim, message, 1.0f / resolution, (1u << log2_T) - 1u, out);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_tables_to_half2(const float* const* tables, uint32_t n_levels, uint32_t log2_T, void* const* tables_h2,
                         float* inv_scale, uint32_t* absmax_scratch, nsig_stream_t stream) {
    if (n_levels == 0) return 0;
    if (!tables || !tables_h2 || !inv_scale || !absmax_scratch) return NSIG_EINVAL;
    if (n_levels > NSIG_MAX_LEVELS || log2_T < 1 || log2_T > 30) return NSIG_EINVAL;
    ShadowPtrs tp;
    for (uint32_t l = 0; l < n_levels; ++l) {
        if (!tables[l] || !tables_h2[l] || (((uintptr_t)tables[l]) & 15) || (((uintptr_t)tables_h2[l]) & 7)) return NSIG_EINVAL;
        tp.src[l] = tables[l];
        tp.dst[l] = reinterpret_cast<__half2*>(tables_h2[l]);
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(absmax_scratch, 0, n_levels * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    const uint32_t n_vec4 = (1u << log2_T) / 2;  // T entries x 2 floats / 4
    const dim3 grid(min(div_up(n_vec4, 256u), 148u * 2u), n_levels);
    k_shadow_absmax<<<grid, 256, 0, st>>>(tp, n_vec4, absmax_scratch);
    NSIG_LAUNCH_CHECK();
    k_shadow_convert<<<grid, 256, 0, st>>>(tp, n_vec4, absmax_scratch, inv_scale);
    NSIG_LAUNCH_CHECK();
    return 0;
}

int nsig_fused_hash_slots(const float* x, uint32_t B, const float* resolutions, uint32_t n_levels, uint32_t log2_T,
                          int32_t* slots, float* weights, nsig_stream_t stream) {
    if (B == 0) return 0;
    if (!x || !resolutions || !slots) return NSIG_EINVAL;
    if (n_levels == 0 || n_levels > NSIG_MAX_LEVELS || log2_T == 0 || log2_T > 30) return NSIG_EINVAL;
    GeomLevels gl;
    for (uint32_t l = 0; l < n_levels; ++l) {
        if (!(resolutions[l] > 0.0f)) return NSIG_EINVAL;
        gl.g[l] = make_level_geom(resolutions[l]);   // the same host helper fill_field_params uses
    }
    const uint64_t total = (uint64_t)B * n_levels;
    k_fused_hash_slots<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, B, gl, n_levels, (1u << log2_T) - 1u, slots, weights);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
