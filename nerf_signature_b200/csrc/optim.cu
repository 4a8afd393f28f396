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
#include <cstdlib>

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

// Streaming accesses of the table Adam: no L1 allocation and EVICT-FIRST in L2, loads and stores alike.  The kernel moves
// ~0.8 GB per step through a 126 MB L2; with the default policy it evicts everything else - the half2 shadow tables the field
// forward gathers, the occupancy bitfield and rays of the march running beside it, and the instruction lines of every
// kernel that starts while it runs (a 2 us kernel of the parallel branch then took 50-100 us to get going because each
// instruction-cache miss went to a saturated HBM).
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld4_stream(const float* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st4_stream(float* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// One rounded operation per line, spelled with intrinsics: the update kernel and the look-ahead sum must produce the SAME
// bits, and left to itself the compiler contracts `beta2*v + omb2*g*g` into different FMAs in the two kernels.
__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, float w1, float beta2, float omb2,
                                          float step_size, float bc2_sqrt, float eps) {
    m = __fmaf_rn(w1, __fsub_rn(g, m), m);                                  // lerp(exp_avg, grad, 1 - beta1)
    v = __fmaf_rn(__fmul_rn(omb2, g), g, __fmul_rn(beta2, v));              // beta2 * v + (1 - beta2) * g * g
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = __fmaf_rn(-step_size, __fdiv_rn(m, denom), p);
}

constexpr int kAdamUnroll = 4;

__global__ void __launch_bounds__(256)
k_msg_adam(AdamPtrs ptrs, uint32_t md, const float* __restrict__ message, const float* __restrict__ G,
           const float* __restrict__ coef, const float* __restrict__ grad_scale,
           const float* __restrict__ found_inf, float beta1, float beta2, float eps, uint32_t n_vec4,
           uint32_t vec4_begin) {
    if (found_inf && *found_inf != 0.0f) return;
    // Grid-stride over a SMALL resident grid (a couple of CTAs per SM, chosen by the host): the kernel is HBM-bound and needs
    // ~50 KB of loads in flight per SM, not the whole register file.  Launched as one CTA per 256 elements it parks 5 long-lived
    // (~100 us) CTAs on every SM, and the kernels of the parallel graph branch (near/far, march) only get SM slots once its
    // last wave has been dispatched - the two branches ran back to back instead of side by side (profiles/r02_timeline_step*).
    const uint64_t pol = evict_first_policy();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += gridDim.x * blockDim.x) {
    const size_t off = (size_t)(vec4_begin + i) * 4;   // this rank's slice of every table (sharded optimizer) or 0
    float4 g = ld4_stream(G + off, pol);
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
            p[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(P + t[u])) + off, pol);
            m[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(M + t[u])) + off, pol);
            v[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(V + t[u])) + off, pol);
        }
#pragma unroll
        for (int u = 0; u < kAdamUnroll; ++u) {
            if (m0 + u >= md) break;
            const float step_size = __ldg(coef + 2 * t[u]), bc2_sqrt = __ldg(coef + 2 * t[u] + 1);
            adam_elem(p[u].x, m[u].x, v[u].x, g.x, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p[u].y, m[u].y, v[u].y, g.y, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p[u].z, m[u].z, v[u].z, g.z, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p[u].w, m[u].w, v[u].w, g.w, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            st4_stream(reinterpret_cast<float*>(__ldg(P + t[u])) + off, p[u], pol);
            st4_stream(reinterpret_cast<float*>(__ldg(M + t[u])) + off, m[u], pol);
            st4_stream(reinterpret_cast<float*>(__ldg(V + t[u])) + off, v[u], pol);
        }
    }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The same update with the data in flight held in SHARED MEMORY instead of registers: bulk asynchronous copies (TMA,
// cp.async.bulk + mbarrier) stream [G | p | m | v] chunks of 256 float4 through a 4-stage ring, 256 threads update a chunk in
// place and a bulk copy writes it back.  k_msg_adam needs ~100 registers x 256 threads per 48 KB of loads in flight, i.e. two
// resident CTAs own 52 k of an SM's 64 k registers for the ~145 us the update streams, and the march of the parallel graph
// branch (14 k registers per CTA) does not get a single CTA in: the two branches ran back to back (profiles/
// r02_graph_offsets_final.txt: the march ends at 253 us whenever it starts).  Here a CTA holds 48 KB in flight with ~40
// registers per thread, so HBM stays saturated while most of the register file is free for the latency-bound neighbour.
// Items = (chunk, table) pairs, dealt to the CTAs as contiguous ranges; arithmetic = adam_elem (bit-identical to k_msg_adam).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTmaChunk = 256;                       // float4 per array and item
constexpr int kTmaStages = 4;
constexpr int kTmaStageBytes = 4 * kTmaChunk * 16;   // G, p, m, v

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
}

__global__ void __launch_bounds__(256)
k_msg_adam_tma(AdamPtrs ptrs, uint32_t md, const float* __restrict__ message, const float* __restrict__ G,
               const float* __restrict__ coef, const float* __restrict__ grad_scale, const float* __restrict__ found_inf,
               float beta1, float beta2, float eps, uint32_t n_vec4, uint32_t vec4_begin) {
    if (found_inf && *found_inf != 0.0f) return;
    extern __shared__ __align__(128) unsigned char ring[];      // kTmaStages x [G | p | m | v]
    __shared__ __align__(8) uint64_t full[kTmaStages];
    __shared__ uint64_t s_ptr[3][NSIG_MAX_MSG_TABLES / 2];       // (param, exp_avg, exp_avg_sq) of the table each bit selects
    __shared__ float s_coef[NSIG_MAX_MSG_TABLES / 2][2];
    const uint32_t tid = threadIdx.x;
    for (uint32_t mi = tid; mi < md; mi += blockDim.x) {
        const uint32_t t = 2 * mi + (((uint32_t)(int)message[mi]) & 1u);
        s_ptr[0][mi] = ptrs.table[t];
        s_ptr[1][mi] = ptrs.table[ptrs.n_tables + t];
        s_ptr[2][mi] = ptrs.table[2 * ptrs.n_tables + t];
        s_coef[mi][0] = coef[2 * t];
        s_coef[mi][1] = coef[2 * t + 1];
    }
    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t pol = evict_first_policy();
    uint64_t pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    const uint32_t n_chunks = (n_vec4 + kTmaChunk - 1) / kTmaChunk;
    const uint64_t n_items = (uint64_t)n_chunks * md;
    const uint64_t i0 = n_items * blockIdx.x / gridDim.x, i1 = n_items * (blockIdx.x + 1) / gridDim.x;
    const float inv = grad_scale ? 1.0f / *grad_scale : 1.0f;
    const float w1 = 1.0f - beta1, omb2 = 1.0f - beta2;

    auto issue_loads = [&](uint64_t item, int s) {              // thread 0 only
        const uint32_t chunk = (uint32_t)(item / md), mi = (uint32_t)(item - (uint64_t)chunk * md);
        const uint32_t v0 = chunk * kTmaChunk, nv = min((uint32_t)kTmaChunk, n_vec4 - v0), bytes = nv * 16;
        const size_t off = ((size_t)vec4_begin + v0) * 4;
        const uint32_t bar = smem_addr(&full[s]), base = smem_addr(ring + (size_t)s * kTmaStageBytes);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(4 * bytes) : "memory");
        bulk_load(base, G + off, bytes, bar, pol_keep);   // G is re-read for every table: it must stay in L2
        bulk_load(base + kTmaChunk * 16, reinterpret_cast<const float*>(s_ptr[0][mi]) + off, bytes, bar, pol);
        bulk_load(base + 2 * kTmaChunk * 16, reinterpret_cast<const float*>(s_ptr[1][mi]) + off, bytes, bar, pol);
        bulk_load(base + 3 * kTmaChunk * 16, reinterpret_cast<const float*>(s_ptr[2][mi]) + off, bytes, bar, pol);
    };
    if (tid == 0)
        for (int k = 0; k < kTmaStages && i0 + k < i1; ++k) issue_loads(i0 + k, k);

    for (uint64_t item = i0; item < i1; ++item) {
        const uint32_t k = (uint32_t)(item - i0), s = k % kTmaStages, parity = (k / kTmaStages) & 1u;
        const uint32_t chunk = (uint32_t)(item / md), mi = (uint32_t)(item - (uint64_t)chunk * md);
        const uint32_t v0 = chunk * kTmaChunk, nv = min((uint32_t)kTmaChunk, n_vec4 - v0);
        {   // wait for the stage's four copies
            const uint32_t bar = smem_addr(&full[s]);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        float4* st = reinterpret_cast<float4*>(ring + (size_t)s * kTmaStageBytes);
        if (tid < nv) {
            float4 g = st[tid], p = st[kTmaChunk + tid], m = st[2 * kTmaChunk + tid], v = st[3 * kTmaChunk + tid];
            g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
            const float step_size = s_coef[mi][0], bc2_sqrt = s_coef[mi][1];
            adam_elem(p.x, m.x, v.x, g.x, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p.y, m.y, v.y, g.y, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p.z, m.z, v.z, g.z, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            adam_elem(p.w, m.w, v.w, g.w, w1, beta2, omb2, step_size, bc2_sqrt, eps);
            st[kTmaChunk + tid] = p; st[2 * kTmaChunk + tid] = m; st[3 * kTmaChunk + tid] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
            const size_t off = ((size_t)vec4_begin + v0) * 4;
            const uint32_t base = smem_addr(st), bytes = nv * 16;
            bulk_store(reinterpret_cast<float*>(s_ptr[0][mi]) + off, base + kTmaChunk * 16, bytes, pol);
            bulk_store(reinterpret_cast<float*>(s_ptr[1][mi]) + off, base + 2 * kTmaChunk * 16, bytes, pol);
            bulk_store(reinterpret_cast<float*>(s_ptr[2][mi]) + off, base + 3 * kTmaChunk * 16, bytes, pol);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (k >= 1) {   // the PREVIOUS item's write-back has left shared memory: refill its stage
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                const uint64_t nxt = item - 1 + kTmaStages;
                if (nxt < i1) issue_loads(nxt, (int)((k - 1) % kTmaStages));
            }
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}


// ---------------------------------------------------------------------------------------------------------------
// Look-ahead table sum: S = sum_i table[2i + next_i] AS IT WILL BE once k_msg_adam has applied the pending update (the one
// selected by `applied`), without writing anything but S.  A table that the pending update touches AND the next message
// selects (next_i == applied_i) contributes p_new computed from (p, m, v, G) with the very same adam_elem; any other table
// contributes its current p.  The accumulation order is k_msg_table_sum's (bit 0 first), so S is bit-identical to running
// the update and then the table sum - but the field forward of the next step no longer has to wait for the 0.8 GB of
// Adam traffic, which can then run anywhere before the next backward touches G (harness: next to the decoder).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_msg_sum_ahead(AdamPtrs ptrs, uint32_t md, const float* __restrict__ applied, const float* __restrict__ next,
                const float* __restrict__ G, const float* __restrict__ coef, const float* __restrict__ grad_scale,
                const float* __restrict__ found_inf, float beta1, float beta2, float eps, uint32_t n_vec4,
                uint32_t vec4_begin, float* __restrict__ S) {
    const bool skip = found_inf && *found_inf != 0.0f;   // the pending update will be skipped: sum the tables as they are
    const uint64_t pol = evict_first_policy();
    const float w1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const uint64_t* P = ptrs.table;
    const uint64_t* M = ptrs.table + ptrs.n_tables;
    const uint64_t* V = ptrs.table + 2 * ptrs.n_tables;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += gridDim.x * blockDim.x) {
        const size_t off = (size_t)(vec4_begin + i) * 4;
        float4 g = ld4_stream(G + off, pol);
        if (grad_scale) {
            const float inv = 1.0f / *grad_scale;
            g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t m0 = 0; m0 < md; m0 += kAdamUnroll) {
            float4 p[kAdamUnroll], m[kAdamUnroll], v[kAdamUnroll];
            uint32_t t[kAdamUnroll];
            bool upd[kAdamUnroll];
#pragma unroll
            for (int u = 0; u < kAdamUnroll; ++u) {
                const uint32_t mi = min(m0 + u, md - 1);
                const uint32_t bn = ((uint32_t)(int)__ldg(next + mi)) & 1u, ba = ((uint32_t)(int)__ldg(applied + mi)) & 1u;
                t[u] = 2 * mi + bn;
                upd[u] = !skip && bn == ba;       // uniform over the grid
                p[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(P + t[u])) + off, pol);
                if (upd[u]) {
                    m[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(M + t[u])) + off, pol);
                    v[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(V + t[u])) + off, pol);
                }
            }
#pragma unroll
            for (int u = 0; u < kAdamUnroll; ++u) {
                if (m0 + u >= md) break;
                if (upd[u]) {
                    const float step_size = __ldg(coef + 2 * t[u]), bc2_sqrt = __ldg(coef + 2 * t[u] + 1);
                    adam_elem(p[u].x, m[u].x, v[u].x, g.x, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    adam_elem(p[u].y, m[u].y, v[u].y, g.y, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    adam_elem(p[u].z, m[u].z, v[u].z, g.z, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    adam_elem(p[u].w, m[u].w, v[u].w, g.w, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                }
                acc.x += p[u].x; acc.y += p[u].y; acc.z += p[u].z; acc.w += p[u].w;
            }
        }
        *reinterpret_cast<float4*>(S + off) = acc;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Update AND the next message's table sum in one pass: the update of the tables `applied` selects (k_msg_adam's loop,
// adam_elem) with S = sum_i table[2i + next_i] accumulated on the way - from the freshly updated value where the next message
// selects the table being updated (next_i == applied_i), from one extra read of the sibling table otherwise (on average
// half of the bits: +64 MiB of reads at message_dim 32 next to the update's 0.8 GB).  Replaces k_msg_adam followed by
// k_msg_table_sum (message_dim x 4 MiB read again, one more kernel and one more cross-branch dependency on the step's
// critical path: the field forward waits for S).  Accumulation order = k_msg_table_sum's, so S has the bits of
// update-then-sum.  A skipped update (found_inf) leaves the tables alone and sums them as they are.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_msg_adam_sum(AdamPtrs ptrs, uint32_t md, const float* __restrict__ applied, const float* __restrict__ next,
               const float* __restrict__ G, const float* __restrict__ coef, const float* __restrict__ grad_scale,
               const float* __restrict__ found_inf, float beta1, float beta2, float eps, uint32_t n_vec4,
               uint32_t vec4_begin, float* __restrict__ S) {
    const bool skip = found_inf && *found_inf != 0.0f;
    const uint64_t pol = evict_first_policy();
    const float w1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    const uint64_t* P = ptrs.table;
    const uint64_t* M = ptrs.table + ptrs.n_tables;
    const uint64_t* V = ptrs.table + 2 * ptrs.n_tables;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += gridDim.x * blockDim.x) {
        const size_t off = (size_t)(vec4_begin + i) * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!skip) {
            g = ld4_stream(G + off, pol);
            if (grad_scale) {
                const float inv = 1.0f / *grad_scale;
                g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
            }
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t m0 = 0; m0 < md; m0 += kAdamUnroll) {
            float4 p[kAdamUnroll], m[kAdamUnroll], v[kAdamUnroll], q[kAdamUnroll];
            uint32_t t[kAdamUnroll];
            bool same[kAdamUnroll];
#pragma unroll
            for (int u = 0; u < kAdamUnroll; ++u) {
                const uint32_t mi = min(m0 + u, md - 1);
                const uint32_t ba = ((uint32_t)(int)__ldg(applied + mi)) & 1u, bn = ((uint32_t)(int)__ldg(next + mi)) & 1u;
                t[u] = 2 * mi + ba;
                same[u] = !skip && bn == ba;      // uniform over the grid
                if (!skip) {
                    p[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(P + t[u])) + off, pol);
                    m[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(M + t[u])) + off, pol);
                    v[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(V + t[u])) + off, pol);
                }
                if (!same[u]) q[u] = ld4_stream(reinterpret_cast<const float*>(__ldg(P + 2 * mi + bn)) + off, pol);
            }
#pragma unroll
            for (int u = 0; u < kAdamUnroll; ++u) {
                if (m0 + u >= md) break;
                if (!skip) {
                    const float step_size = __ldg(coef + 2 * t[u]), bc2_sqrt = __ldg(coef + 2 * t[u] + 1);
                    adam_elem(p[u].x, m[u].x, v[u].x, g.x, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    adam_elem(p[u].y, m[u].y, v[u].y, g.y, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    adam_elem(p[u].z, m[u].z, v[u].z, g.z, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    adam_elem(p[u].w, m[u].w, v[u].w, g.w, w1, beta2, omb2, step_size, bc2_sqrt, eps);
                    st4_stream(reinterpret_cast<float*>(__ldg(P + t[u])) + off, p[u], pol);
                    st4_stream(reinterpret_cast<float*>(__ldg(M + t[u])) + off, m[u], pol);
                    st4_stream(reinterpret_cast<float*>(__ldg(V + t[u])) + off, v[u], pol);
                }
                const float4 s = same[u] ? p[u] : q[u];
                acc.x += s.x; acc.y += s.y; acc.z += s.z; acc.w += s.w;
            }
        }
        *reinterpret_cast<float4*>(S + off) = acc;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// torch.amp.GradScaler's per-step work as ONE kernel over the flat gradient bucket [dL/dS | decoder gradients]
// (nerf/utils_wtmk_disen.py:1175-1181: scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update()):
//   * the non-finite check of every gradient (GradScaler._unscale_grads_ -> _amp_foreach_non_finite_check_and_unscale_:
//     a multi-tensor launch over ~40 tensors, 15 us) as one vectorised pass over 5 MB that is still in L2;
//   * the last CTA to finish publishes found_inf and the scale THIS step's optimizer kernels must divide by, bumps the
//     Adam step count of the flat parameter group, and applies _amp_update_scale_ (backoff on overflow, growth after
//     growth_interval clean steps) right away - the optimizer kernels read the published copy, so nothing has to run
//     after them.  Replaces 2 fills + the check + a reduction of found_inf + amp_update_scale (6 launches).
// scratch: uint32[2] {flag, ticket}, zero-initialised once; the kernel leaves it zeroed.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_grad_check_update_scale(const float* __restrict__ flat, uint32_t n4, uint32_t n, float* __restrict__ scale,
                          int32_t* __restrict__ growth_tracker, float growth_factor, float backoff_factor,
                          int32_t growth_interval, float* __restrict__ found_inf, float* __restrict__ step_scale,
                          float* __restrict__ adam_step, uint32_t* __restrict__ scratch,
                          const uint32_t* __restrict__ enabled) {
    bool bad = false;
    const float4* f4 = reinterpret_cast<const float4*>(flat);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 v = f4[i];
        // non-finite <=> exponent bits all set
        bad |= ((__float_as_uint(v.x) & 0x7f800000u) == 0x7f800000u) | ((__float_as_uint(v.y) & 0x7f800000u) == 0x7f800000u) |
               ((__float_as_uint(v.z) & 0x7f800000u) == 0x7f800000u) | ((__float_as_uint(v.w) & 0x7f800000u) == 0x7f800000u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) bad |= (__float_as_uint(flat[n4 * 4 + threadIdx.x]) & 0x7f800000u) == 0x7f800000u;
    const bool any_bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (any_bad) atomicOr(scratch, 1u);
        __threadfence();
        const uint32_t ticket = atomicAdd(scratch + 1, 1u);
        if (ticket == gridDim.x - 1) {   // last CTA: every flag update is visible
            __threadfence();
            const bool inf = atomicOr(scratch, 0u) != 0u;
            const float s = *scale;
            *step_scale = s;
            if (enabled && *enabled == 0u) {
                // no optimizer step is pending (deferred-step mode after a flush): make the optimizer kernels skip and leave
                // the scaler state alone
                *found_inf = 1.0f;
                scratch[0] = 0u;
                scratch[1] = 0u;
                return;
            }
            *found_inf = inf ? 1.0f : 0.0f;
            if (inf) {
                *scale = s * backoff_factor;
                *growth_tracker = 0;
            } else {
                if (adam_step) *adam_step += 1.0f;
                const int32_t ok = *growth_tracker + 1;
                if (ok == growth_interval) {
                    const float grown = s * growth_factor;
                    if (isfinite(grown)) *scale = grown;
                    *growth_tracker = 0;
                } else {
                    *growth_tracker = ok;
                }
            }
            scratch[0] = 0u;
            scratch[1] = 0u;
        }
    }
}

// Adam (torch.optim.Adam semantics: no weight decay, no amsgrad) over ONE flat parameter vector - the HiDDeN decoder's
// 38 tensors re-pointed at a single buffer - instead of torch's two multi-tensor launches.  step: device float already
// incremented for this step by k_grad_check_update_scale.
__global__ void __launch_bounds__(256)
k_flat_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, uint32_t n,
            const float* __restrict__ step, const float* __restrict__ grad_scale, const float* __restrict__ found_inf,
            float lr, const float* __restrict__ lr_dev, float beta1, float beta2, float eps) {
    if (found_inf && *found_inf != 0.0f) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = (double)*step;
    const double lr_now = lr_dev ? (double)*lr_dev : (double)lr;
    const float step_size = (float)(lr_now / (1.0 - pow((double)beta1, s)));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, s));
    const float inv = grad_scale ? 1.0f / *grad_scale : 1.0f;
    float pp = p[i], mm = m[i], vv = v[i];
    adam_elem(pp, mm, vv, g[i] * inv, 1.0f - beta1, beta2, 1.0f - beta2, step_size, bc2_sqrt, eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
}

}  // namespace nsig

using namespace nsig;

namespace {
uint32_t adam_ctas_per_sm() {
    static const uint32_t per_sm = [] { const char* e = getenv("NSIG_ADAM_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return (uint32_t)(v > 0 ? v : 4); }();
    return per_sm;
}
int adam_range(uint32_t log2_T, uint32_t& elem_begin, uint32_t& elem_count) {
    const uint32_t total = 2u << log2_T;         // T entries x 2 floats
    if (elem_count == 0) { elem_begin = 0; elem_count = total; }
    if ((elem_begin | elem_count) & 3u || elem_begin > total || elem_count > total - elem_begin) return NSIG_EINVAL;
    return 0;
}
}  // namespace

extern "C" int nsig_msg_adam_step(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                                  const float* message, const float* G, float* steps, float* coef,
                                  const float* grad_scale, const float* found_inf, float lr, float beta1,
                                  float beta2, float eps, uint32_t log2_T, const float* lr_dev,
                                  uint32_t elem_begin, uint32_t elem_count, uint32_t steps_prepared,
                                  nsig_stream_t stream) {
    if (!ptr_table || !message || !G || !steps || !coef) return NSIG_EINVAL;
    if (message_dim == 0 || 2 * message_dim > n_tables || n_tables > NSIG_MAX_MSG_TABLES) return NSIG_EINVAL;
    if (log2_T < 1 || log2_T > 30 || (((uintptr_t)G) & 15)) return NSIG_EINVAL;
    if (adam_range(log2_T, elem_begin, elem_count)) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (!steps_prepared) {   // (a preceding nsig_msg_adam_lookahead_sum for this update already advanced steps / coef)
        k_msg_adam_prepare<<<div_up(message_dim, 128), 128, 0, st>>>(message_dim, message, steps, coef, found_inf,
                                                                     (double)lr, lr_dev, (double)beta1, (double)beta2);
        NSIG_LAUNCH_CHECK();
    }
    const uint32_t n_vec4 = elem_count / 4;
    AdamPtrs ptrs{ptr_table, n_tables};
    // NSIG_ADAM_TMA=<CTAs per SM> selects the shared-memory-staged kernel.  Measured (round 2, call X): bit-identical; alone 164 us
    // vs 157 us for the register kernel; inside the step the march does overlap it then (ends at 200 us instead of 253 us) but
    // the update stretches to 192 us beside it and the field forward starts at 251 us instead of 257 us - step 0.972-0.991 ms vs
    // 0.955 ms.  The register kernel stays the default.
    static const int tma_per_sm = [] { const char* e = getenv("NSIG_ADAM_TMA"); return e ? atoi(e) : 0; }();
    if (tma_per_sm > 0 && 2 * message_dim <= NSIG_MAX_MSG_TABLES) {
        const size_t smem = (size_t)kTmaStages * kTmaStageBytes;
        static bool set = false;
        if (!set) { cudaFuncSetAttribute(k_msg_adam_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set = true; }
        const uint64_t items = (uint64_t)div_up(n_vec4, (uint32_t)kTmaChunk) * message_dim;
        const uint64_t cap = 148ull * (uint64_t)tma_per_sm;
        const uint32_t grid = (uint32_t)(items < cap ? items : cap);
        k_msg_adam_tma<<<grid, 256, smem, st>>>(ptrs, message_dim, message, G, coef, grad_scale, found_inf,
                                                beta1, beta2, eps, n_vec4, elem_begin / 4);
        NSIG_LAUNCH_CHECK();
        return 0;
    }
    const uint32_t grid = min(div_up(n_vec4, 256u), 148u * adam_ctas_per_sm());
    k_msg_adam<<<grid, 256, 0, st>>>(ptrs, message_dim, message, G, coef, grad_scale, found_inf,
                                     beta1, beta2, eps, n_vec4, elem_begin / 4);
    NSIG_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsig_msg_adam_step_sum(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                                      const float* message_applied, const float* message_next, const float* G,
                                      float* steps, float* coef, const float* grad_scale, const float* found_inf,
                                      float lr, float beta1, float beta2, float eps, uint32_t log2_T,
                                      const float* lr_dev, uint32_t elem_begin, uint32_t elem_count, float* S,
                                      nsig_stream_t stream) {
    if (!ptr_table || !message_applied || !message_next || !G || !steps || !coef || !S) return NSIG_EINVAL;
    if (message_dim == 0 || 2 * message_dim > n_tables || n_tables > NSIG_MAX_MSG_TABLES) return NSIG_EINVAL;
    if (log2_T < 1 || log2_T > 30 || (((uintptr_t)G) & 15) || (((uintptr_t)S) & 15)) return NSIG_EINVAL;
    if (adam_range(log2_T, elem_begin, elem_count)) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    k_msg_adam_prepare<<<div_up(message_dim, 128), 128, 0, st>>>(message_dim, message_applied, steps, coef, found_inf,
                                                                 (double)lr, lr_dev, (double)beta1, (double)beta2);
    NSIG_LAUNCH_CHECK();
    const uint32_t n_vec4 = elem_count / 4;
    AdamPtrs ptrs{ptr_table, n_tables};
    const uint32_t grid = min(div_up(n_vec4, 256u), 148u * adam_ctas_per_sm());
    k_msg_adam_sum<<<grid, 256, 0, st>>>(ptrs, message_dim, message_applied, message_next, G, coef, grad_scale, found_inf,
                                         beta1, beta2, eps, n_vec4, elem_begin / 4, S);
    NSIG_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsig_msg_adam_lookahead_sum(const uint64_t* ptr_table, uint32_t n_tables, uint32_t message_dim,
                                           const float* message_applied, const float* message_next, const float* G,
                                           float* steps, float* coef, const float* grad_scale, const float* found_inf,
                                           float lr, float beta1, float beta2, float eps, uint32_t log2_T,
                                           const float* lr_dev, uint32_t elem_begin, uint32_t elem_count, float* S,
                                           nsig_stream_t stream) {
    if (!ptr_table || !message_applied || !message_next || !G || !steps || !coef || !S) return NSIG_EINVAL;
    if (message_dim == 0 || 2 * message_dim > n_tables || n_tables > NSIG_MAX_MSG_TABLES) return NSIG_EINVAL;
    if (log2_T < 1 || log2_T > 30 || (((uintptr_t)G) & 15) || (((uintptr_t)S) & 15)) return NSIG_EINVAL;
    if (adam_range(log2_T, elem_begin, elem_count)) return NSIG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    k_msg_adam_prepare<<<div_up(message_dim, 128), 128, 0, st>>>(message_dim, message_applied, steps, coef, found_inf,
                                                                 (double)lr, lr_dev, (double)beta1, (double)beta2);
    NSIG_LAUNCH_CHECK();
    const uint32_t n_vec4 = elem_count / 4;
    AdamPtrs ptrs{ptr_table, n_tables};
    static const uint32_t per_sm = [] { const char* e = getenv("NSIG_SUM_AHEAD_CTAS_PER_SM"); const int v = e ? atoi(e) : 0; return (uint32_t)(v > 0 ? v : 2); }();
    const uint32_t grid = min(div_up(n_vec4, 256u), 148u * per_sm);
    k_msg_sum_ahead<<<grid, 256, 0, st>>>(ptrs, message_dim, message_applied, message_next, G, coef, grad_scale, found_inf,
                                          beta1, beta2, eps, n_vec4, elem_begin / 4, S);
    NSIG_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsig_grad_check_update_scale(const float* flat, uint32_t n, float* scale, int32_t* growth_tracker,
                                            float growth_factor, float backoff_factor, int32_t growth_interval,
                                            float* found_inf, float* step_scale, float* adam_step, uint32_t* scratch,
                                            const uint32_t* enabled, nsig_stream_t stream) {
    if (!flat || !scale || !growth_tracker || !found_inf || !step_scale || !scratch) return NSIG_EINVAL;
    if ((((uintptr_t)flat) & 15) || growth_interval < 1) return NSIG_EINVAL;
    const uint32_t n4 = n / 4;
    const uint32_t grid = max(1u, min(div_up(n4, 256u * 4u), 148u * 4u));
    k_grad_check_update_scale<<<grid, 256, 0, (cudaStream_t)stream>>>(flat, n4, n, scale, growth_tracker, growth_factor,
                                                                        backoff_factor, growth_interval, found_inf,
                                                                        step_scale, adam_step, scratch, enabled);
    NSIG_LAUNCH_CHECK();
    return 0;
}

extern "C" int nsig_flat_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, uint32_t n,
                                   const float* step, const float* grad_scale, const float* found_inf, float lr,
                                   const float* lr_dev, float beta1, float beta2, float eps, nsig_stream_t stream) {
    if (n == 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !step) return NSIG_EINVAL;
    k_flat_adam<<<div_up(n, 256u), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, step, grad_scale,
                                                                    found_inf, lr, lr_dev, beta1, beta2, eps);
    NSIG_LAUNCH_CHECK();
    return 0;
}
