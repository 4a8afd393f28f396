// tc_probe.cu - bring-up probe for the tcgen05 path of the fused field backward (csrc/field_tc.cu).
//
// One CTA of 128 threads computes D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T (fp16) with tcgen05.mma
// (cta_group::1, kind::f16, both operands K-major, NO swizzle), exactly the way the production kernel stages its operands:
// thread r writes row r of A as 16-byte chunks into the canonical "interleaved" core-matrix layout
//   byte(r, k) = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
// the MMA is issued by one thread per K=16 slice, completion arrives on an mbarrier through tcgen05.commit, and every
// thread reads its own accumulator row back with tcgen05.ld.32x32b.  The result is checked against a host GEMM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -cudart static -o tools/scratch/tc_probe tools/tc_probe.cu
//   tools/scratch/tc_probe [variant]      variant 0: LBO = K-direction stride (128 B), SBO = M/N-direction stride
//                                         variant 1: the two swapped (diagnosis only)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

#define TMEM_LD16(taddr, v)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),        \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])   \
                 : "r"(taddr))

__global__ void __launch_bounds__(128)
k_probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int N, int K, int variant,
        int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sA = smem;                         // 128 x K halfs
    uint8_t* sB = smem + 128 * K * 2;           // N x K halfs
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t chunks = K / 8;
    const uint32_t k_stride = 128;                    // adjacent K-chunks: consecutive 128-byte core matrices
    const uint32_t mn_stride = chunks * 128;          // adjacent 8-row groups
    for (uint32_t c = 0; c < chunks; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)t * K + c * 8);
        *reinterpret_cast<uint4*>(sA + (t / 8) * mn_stride + c * k_stride + (t % 8) * 16) = v;
    }
    if (t < N)
        for (uint32_t c = 0; c < chunks; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)t * K + c * 8);
            *reinterpret_cast<uint4*>(sB + (t / 8) * mn_stride + c * k_stride + (t % 8) * 16) = v;
        }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (t == 0) {
        // instruction descriptor: D fp32 (bit 4), A/B fp16 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t lbo = variant == 0 ? k_stride : mn_stride, sbo = variant == 0 ? mn_stride : k_stride;
        for (int k = 0; k < K / 16; ++k) {
            const uint64_t da = make_desc(smem_u32(sA) + k * 2 * k_stride, lbo, sbo);
            const uint64_t db = make_desc(smem_u32(sB) + k * 2 * k_stride, lbo, sbo);
            mma_f16_ss(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    // bounded wait on phase 0
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    if (!done && t == 0) status[0] = 1;   // timeout
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            TMEM_LD16(tmem_base + lane_base + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) D[(size_t)t * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u));
}

static int run_case(int N, int K, int variant) {
    std::vector<__half> A(128 * K), B(N * K);
    std::vector<float> Af(128 * K), Bf(N * K), ref(128 * N), out(128 * N, -777.f);
    srand(1234 + N * 131 + K);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Bf[i] = __half2float(B[i]); }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)Af[m * K + k] * Bf[n * K + k];
            ref[m * N + n] = (float)s;
        }
    __half *dA, *dB; float* dD; int* dS;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, out.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, out.data(), out.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    const size_t smem = (size_t)(128 + N) * K * 2;
    k_probe<<<1, 128, smem>>>(dA, dB, dD, N, K, variant, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    if (e == cudaSuccess) { cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); }
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < out.size(); ++i) { double d = fabs((double)out[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-2) ++bad; }
    printf("variant %d  N=%3d K=%3d : cuda=%s timeout=%d max_abs_err=%.5f bad=%d/%zu  D[0..3]=%.3f %.3f %.3f %.3f ref=%.3f %.3f %.3f %.3f\n",
           variant, N, K, cudaGetErrorString(e), st, maxerr, bad, out.size(), out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    return (e == cudaSuccess && st == 0 && bad == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    int fails = 0;
    const int cases[][2] = {{64, 32}, {16, 64}, {64, 64}, {64, 16}, {16, 16}, {32, 64}};
    for (auto& c : cases) fails += run_case(c[0], c[1], variant);
    printf("variant %d: %s\n", variant, fails ? "FAIL" : "ALL OK");
    return fails ? 1 : 0;
}
