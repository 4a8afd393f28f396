// optim.cu — Adam update of the message tables selected by this step's message, from the single
// gradient G = dL/dS.
//
// Reference semantics (nerf/utils_wtmk_disen.py:1175-1181 with torch.optim.Adam over
// network_wtmk_tcnn.py:179-188 get_params): every selected table embeddings[2i + bit_i] receives the same
// gradient G, every unselected table has grad None and is skipped entirely (no moment decay,
// no step increment).  torch realises that as message_dim dense [2^19,2] gradient tensors, an unscale
// pass over each and a multi-tensor Adam over (param, grad, m, v) — 7 x 4 MiB of traffic per selected
// table plus message_dim 4 MiB copies.  Here G is read once per element and each selected table streams
// p, m, v in and out: 6 x 4 MiB per table, HBM-bound, with the table choice made on the device from the
// message vector (no host round trip, CUDA-graph capturable).
#include "nsig_common.cuh"

namespace nsig {

// one thread per table: bump the step count of the selected tables and precompute the two scalars the
// element kernel needs.  coef[t] = { lr / (1 - beta1^step), sqrt(1 - beta2^step) } (double arithmetic, as
// torch's fused Adam evaluates its bias corrections).
__global__ void k_msg_adam_prepare(uint32_t md, const float* __restrict__ message, float* __restrict__ steps,
                                   float* __restrict__ coef, const float* __restrict__ found_inf, double lr,
                                   const float* __restrict__ lr_dev, double beta1, double beta2) {
    const uint32_t i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= md) return;
    if (found_inf && *found_inf != 0.0f) return;  // GradScaler: skip the whole step
    if (lr_dev) lr = (double)*lr_dev;  // learning rate read on the device: schedulers keep working under graph replay
    const uint32_t t = 2 * i + (((uint32_t)(int)message[i]) & 1u);
    const float s = steps[t] + 1.0f;
    steps[t] = s;
    coef[2 * t] = (float)(lr / (1.0 - pow(beta1, (double)s)));
    coef[2 * t + 1] = (float)sqrt(1.0 - pow(beta2, (double)s));
}

struct AdamPtrs {  // device array layout: [3][n_tables] of pointers (param, exp_avg, exp_avg_sq)
    const uint64_t* table;
    uint32_t n_tables;
};

__device__ __forceinline__ float4 ld4_stream(const float* p) { return ld_stream4(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, float w1, float beta2, float omb2,
                                          float step_size, float bc2_sqrt, float eps) {
    m = m + w1 * (g - m);                       // lerp(exp_avg, grad, 1 - beta1)
    v = beta2 * v + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p = p - step_size * (m / denom);
}

constexpr int kAdamUnroll = 4;

__global__ void __launch_bounds__(256)
k_msg_adam(AdamPtrs ptrs, uint32_t md, const float* __restrict__ message, const float* __restrict__ G,
           const float* __restrict__ coef, const float* __restrict__ grad_scale,
           const float* __restrict__ found_inf, float beta1, float beta2, float eps, uint32_t n_vec4) {
    if (found_inf && *found_inf != 0.0f) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec4) return;
    const size_t off = (size_t)i * 4;
    float4 g = ld4_stream(G + off);
    if (grad_scale) {
        const float inv = 1.0f / *grad_scale;  // GradScaler.unscale_: grad * (1/scale)
        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    }
    const float w1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const uint64_t* P = ptrs.table;
    const uint64_t* M = ptrs.table + ptrs.n_tables;
    const uint64_t* V = ptrs.table + 2 * ptrs.n_tables;
    for (uint32_t m0 = 0; m0 < md; m0 += kAdamUnroll) {
        float4 p[kAdamUnroll], m[kAdamUnroll], v[kAdamUnroll];
        uint32_t t[kAdamUnroll];
#pragma unroll
        for (int u = 0; u < kAdamUnroll; ++u) {
            const uint32_t mi = min(m0 + u, md - 1);
            t[u] = 2 * mi + (((uint32_t)(int)__ldg(message + mi)) & 1u);
            p[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(P + t[u])) + off);
            m[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(M + t[u])) + off);
            v[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(V + t[u])) + off);
        }
#pragma unroll
        for (int u = 0; u < kAdamUnroll; ++u) {
            if (m0 + u >= md) break;
            const float step_size = __ldg(coef + 2 * t[u]), bc2_sqrt = __ldg(coef + 2 * t[u] + 1);
            adam_elem(p[u].x, m[u].x, v[u].x, g.x, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p[u].y, m[u].y, v[u].y, g.y, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p[u].z, m[u].z, v[u].z, g.z, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p[u].w, m[u].w, v[u].w, g.w, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(__ldg(P + t[u])) + off) = p[u];
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(__ldg(M + t[u])) + off) = m[u];
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(__ldg(V + t[u])) + off) = v[u];
        }
    }
}

}  // namespace nsig

using namespace nsig;

extern "C" int nsig_msg_adam_step(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                                  const float* message, const float* G, float* steps, float* coef,
                                  const float* grad_scale, const float* found_inf, float lr, float beta1,
                                  float beta2, float eps, uint32_t log2_T, const float* lr_dev,
                                  nsig_stream_t stream) {
    if (!ptr_table || !message || !G || !steps || !coef) return NSIG_EINVAL;
    if (message_dim == 0 || 2 * message_dim > n_tables || n_tables > NSIG_MAX_MSG_TABLES) return NSIG_EINVAL;
    if (log2_T < 1 || log2_T > 30 || (((uintptr_t)G) & 15)) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    k_msg_adam_prepare<<<div_up(message_dim, 128), 128, 0, st>>>(message_dim, message, steps, coef, found_inf,
                                                                 (double)lr, lr_dev, (double)beta1, (double)beta2);
    NSIG_LAUNCH_CHECK();
    const uint32_t n_vec4 = (1u << log2_T) / 2;  // T entries x 2 floats / 4
    AdamPtrs ptrs{ptr_table, n_tables};
    k_msg_adam<<<div_up(n_vec4, 256), 256, 0, st>>>(ptrs, message_dim, message, G, coef, grad_scale, found_inf,
                                                    beta1, beta2, eps, n_vec4);
    NSIG_LAUNCH_CHECK();
    return 0;
}
