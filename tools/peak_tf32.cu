// TF32 tensor-core peak probe for B200 (sm_100a): back-to-back tcgen05.mma.kind::tf32, M=128 N=256 K=8, operands resident
// in shared memory (K-major, SWIZZLE_128B), two alternating 256-column TMEM accumulators, one CTA per SM.
// Prints one JSON line: the TF32 roofline denominator for the fp32 leaf GEMM (MEASURED_PEAKS.json has bf16 only); the
// fp32 result rate of the 3xTF32 split scheme is bounded by a third of it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_tf32 tools/peak_tf32.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) k_tf32_peak(int iters, float* sink) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float* a = reinterpret_cast<float*>(smem);                 // [128 mn][32 k] K-major slab, 16 KiB
    float* b = reinterpret_cast<float*>(smem + 16 * 1024);     // [N mn][32 k], N * 128 B
    for (int i = threadIdx.x; i < (16 * 1024 + N * 128) / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        a[i] = ((int)(h >> 20) - 2048) * (1.0f / 4194304.0f);  // small values of both signs: the accumulators stay finite
    }
    (void)b;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) {
        constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(a), sb = smem_u32(b);
        constexpr int NACC = 512 / N;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                mma_tf32(tmem_base + (uint32_t)((it % NACC) * N), umma_desc(sa + ks * 32, 16, 1024, 2), umma_desc(sb + ks * 32, 16, 1024, 2),
                         IDESC, (it >= NACC || ks) ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x < 32) {
        uint32_t r;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(tmem_base) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (sink && threadIdx.x == 0) sink[blockIdx.x] = __uint_as_float(r);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int N>
void run(int sms, float* sink, const char* tag, bool sustained) {
    const int smem = 1024 + 16 * 1024 + N * 128;
    CK(cudaFuncSetAttribute(k_tf32_peak<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 8192;   // x 4 K-steps
    const double flop = 2.0 * 128 * N * 8 * 4.0 * iters * sms;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) k_tf32_peak<N><<<sms, 128, smem>>>(iters, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 10; ++r) {
        CK(cudaEventRecord(e0)); k_tf32_peak<N><<<sms, 128, smem>>>(iters, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf(", \"tf32_m128n%d_burst_tflops\": %.1f, \"%s_clk_per_mma\": %.1f", N, flop / best / 1e9, tag, best * 1e-3 * 1.965e9 / (4.0 * iters));
    if (sustained) {
        int n = 0; float ms = 0;
        CK(cudaEventRecord(e0));
        for (; n < 1200; ++n) k_tf32_peak<N><<<sms, 128, smem>>>(iters, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"tf32_m128n%d_sustained_tflops\": %.1f, \"sustained_ms\": %.0f", N, flop * n / ms / 1e9, ms);
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    float* sink; CK(cudaMalloc(&sink, sizeof(float) * sms));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    run<256>(sms, sink, "n256", true);
    run<128>(sms, sink, "n128", false);
    run<64>(sms, sink, "n64", false);
    printf(", \"how\": \"tcgen05.mma.cta_group::1.kind::tf32 M=128 K=8 from resident smem operands (K-major, SWIZZLE_128B), alternating TMEM accumulators, "
           "1 CTA/SM, CUDA events; burst = best of 10 launches, sustained = back to back for ~2 s; clk_per_mma assumes 1965 MHz\"}\n");
    return 0;
}
