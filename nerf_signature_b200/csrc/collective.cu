// collective.cu — the one exchange step of the ray-sharded training path: the once-per-step
// all-reduce (mean) of the flat gradient bucket [dL/dS | decoder gradients] (~5 MB), as ONE kernel over NVLink /
// NVSwitch peer memory instead of an NCCL call:
//
//   entry barrier  ->  two-shot all-reduce in place  ->  exit barrier
//
// Rank r owns the r-th slice of the bucket.  With NVSwitch multicast (NVLS) it reads the slice already summed by
// the switch (`multimem.ld_reduce.add.v4.f32` on the multicast address), scales it and broadcasts it with
// `multimem.st`; without multicast it sums the slice from every peer's buffer with system-scope loads and stores
// the result into every peer's buffer (plain P2P).  Each address is read and then written by exactly one rank, so
// the reduction is in place.  The barriers are self-resetting flag exchanges in a symmetric flag buffer (one slot
// per (CTA, peer)), which makes the kernel safe to capture in the step's CUDA graph and replay.
//
// Measured motivation (gpurun --gpus 2, bench.py): the NCCL all-reduce inside the captured step costs 0.21 ms of a
// 2.17 ms step; the bucket is only 5.2 MB, so the cost is latency, not bandwidth.
#include "nsig_common.cuh"

namespace nsig {

constexpr int kArMaxWorld = 16;
constexpr int kArThreads = 512;

struct ArParams {
    float* bufs[kArMaxWorld];     // every rank's bucket (peer-mapped), bufs[rank] is the local one
    uint32_t* flags[kArMaxWorld]; // every rank's flag buffer: [grid][world] words, zero-initialised once
    float* mc;                    // multicast address of the bucket, or null
    uint32_t n4;                  // bucket length in float4 units
    uint32_t rank, world;
    float scale;
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_sys4(const float* addr) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys4(float* addr, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// CTA b of every rank meets CTA b of every other rank.  Thread t < world raises the flag (b, my rank) at peer t
// (waiting for it to be clear first) and then takes down the flag (b, t) the peer raised here.
__device__ __forceinline__ void barrier_all_ranks(const ArParams& p) {
    __syncthreads();
    if (threadIdx.x < p.world) {
        __threadfence_system();  // release: this CTA's stores (and, transitively, earlier kernels') before the flag
        const uint32_t peer = threadIdx.x;
        uint32_t* theirs = p.flags[peer] + blockIdx.x * p.world + p.rank;
        while (atomicCAS_system(theirs, 0u, 1u) != 0u) { }
        uint32_t* mine = p.flags[p.rank] + blockIdx.x * p.world + peer;
        while (atomicCAS_system(mine, 1u, 0u) != 1u) { }
        __threadfence_system();  // acquire
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kArThreads)
k_allreduce_two_shot(const ArParams p) {
    barrier_all_ranks(p);  // every rank's gradients are complete
    const uint32_t per = (p.n4 + p.world - 1) / p.world;
    const uint32_t lo = min(p.rank * per, p.n4), hi = min(lo + per, p.n4);
    for (uint32_t i = lo + blockIdx.x * kArThreads + threadIdx.x; i < hi; i += gridDim.x * kArThreads) {
        float4 v;
        if (p.mc) {
            v = multimem_ld_reduce_add(p.mc + (size_t)i * 4);
        } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t r = 0; r < p.world; ++r) {
                const float4 a = ld_sys4(p.bufs[r] + (size_t)i * 4);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
        }
        v.x *= p.scale; v.y *= p.scale; v.z *= p.scale; v.w *= p.scale;
        if (p.mc) {
            multimem_st(p.mc + (size_t)i * 4, v);
        } else {
            for (uint32_t r = 0; r < p.world; ++r) st_sys4(p.bufs[r] + (size_t)i * 4, v);
        }
    }
    barrier_all_ranks(p);  // every slice has landed in every bucket
}

}  // namespace nsig

using namespace nsig;

extern "C" {

uint32_t nsig_allreduce_grid(void) { return 32; }

int nsig_allreduce_mean_inplace(void* const* bufs, void* const* flags, void* multicast, uint32_t n, uint32_t rank,
                                uint32_t world, nsig_stream_t stream) {
    if (n == 0 || world <= 1) return 0;
    if (!bufs || !flags || world > (uint32_t)kArMaxWorld || rank >= world || (n & 3u)) return NSIG_EINVAL;
    ArParams p;
    for (uint32_t r = 0; r < world; ++r) {
        if (!bufs[r] || !flags[r] || (((uintptr_t)bufs[r]) & 15)) return NSIG_EINVAL;
        p.bufs[r] = reinterpret_cast<float*>(bufs[r]);
        p.flags[r] = reinterpret_cast<uint32_t*>(flags[r]);
    }
    p.mc = reinterpret_cast<float*>(multicast);
    p.n4 = n / 4; p.rank = rank; p.world = world; p.scale = 1.0f / (float)world;
    k_allreduce_two_shot<<<nsig_allreduce_grid(), kArThreads, 0, (cudaStream_t)stream>>>(p);
    NSIG_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
