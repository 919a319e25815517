// tcgen05.mma.kind::tf32 issue-cost probe (B200, sm_100a): clocks per MMA instruction for the shapes / operand layouts /
// accumulator patterns the fp32 leaf GEMM can choose from.  Operands are resident in shared memory; results are not checked
// (descriptors are valid, data arbitrary but finite).  One JSON line.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
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

// MODE 0: rotate the accumulator every KSTEPS MMAs (a new product each time); 1: every MMA into the SAME accumulator;
// 2: rotate the accumulator every MMA (K-steps of NACC products interleaved)
// MNMAJOR: both operands MN-major (32-byte-atom swizzle, layout type 1) instead of K-major SWIZZLE_128B
// COMMIT_EVERY > 0: a tcgen05.commit (to a barrier nobody waits on) after every COMMIT_EVERY-th product;
// POLL: the other warps of the CTA spin on mbarrier.try_wait of a barrier that never completes (what the consumer roles of a
// warp-specialised kernel do) instead of sleeping at a CTA barrier
// ROT: operand base address rotates over ROT slots of 16 KiB (A) / 16 KiB (B) from product to product (fresh data per product)
template <int M, int N, int MODE, bool MNMAJOR, int KSTEPS, int COMMIT_EVERY = 0, bool POLL = false, int THREADS = 128, int ROT = 1>
__global__ void __launch_bounds__(THREADS, 1) k_probe(int iters, float* sink, long long* clocks) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2[8], never;
    __shared__ volatile int done_flag;
    __shared__ uint32_t tmem_base_s;
    constexpr int ABYTES = ROT > 1 ? 96 * 1024 : 128 * 32 * 4 * 2, BBYTES = ROT > 1 ? 96 * 1024 : 256 * 32 * 4 * 2;
    float* a = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < (ABYTES + BBYTES) / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        a[i] = ((int)(h >> 20) - 2048) * (1.0f / 4194304.0f);
    }
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        mbar_init(smem_u32(&never), 1);
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar2[i]), 1);
        done_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((MNMAJOR ? 1u : 0u) << 15) | ((MNMAJOR ? 1u : 0u) << 16) |
                                   ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + ABYTES);
        constexpr int NACC = 512 / N;
        constexpr uint32_t LBO = MNMAJOR ? 32 * 128 : 16, SBO = MNMAJOR ? 512 : 1024, LT = MNMAJOR ? 1 : 2;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const uint32_t off = (MNMAJOR ? (uint32_t)(ks * 1024) : (uint32_t)((ks >> 2) * 8192 + (ks & 3) * 32)) + (uint32_t)((it % ROT) * 16384);
                const uint32_t acc = MODE == 1 ? 0u : (MODE == 0 ? (uint32_t)(it % NACC) : (uint32_t)((it * KSTEPS + ks) % NACC));
                mma_tf32(tmem_base + acc * N, umma_desc(sa + off, LBO, SBO, LT), umma_desc(sb + off, LBO, SBO, LT), IDESC,
                         (MODE == 0 ? ks != 0 : it + ks != 0) ? 1u : 0u);
            }
            if (COMMIT_EVERY > 0 && (it % COMMIT_EVERY) == COMMIT_EVERY - 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(it / COMMIT_EVERY) & 7])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) *clocks = t1 - t0;
        done_flag = 1;
    } else if (POLL && threadIdx.x >= 32) {
        while (!done_flag) {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&never)), "r"(0u) : "memory");
        }
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

template <int M, int N, int MODE, bool MNMAJOR, int KSTEPS, int COMMIT_EVERY = 0, bool POLL = false, int THREADS = 128, int ROT = 1>
void run(int sms, float* sink, long long* dclk, const char* tag, bool first) {
    const int smem = ROT > 1 ? 1024 + 192 * 1024 + 4096 : 1024 + 96 * 1024 + 4096;
    auto kfn = k_probe<M, N, MODE, MNMAJOR, KSTEPS, COMMIT_EVERY, POLL, THREADS, ROT>;
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 4096;
    for (int w = 0; w < 2; ++w) kfn<<<sms, THREADS, smem>>>(iters, sink, dclk);
    CK(cudaDeviceSynchronize());
    long long best = 1ll << 60;
    for (int r = 0; r < 5; ++r) {
        kfn<<<sms, THREADS, smem>>>(iters, sink, dclk);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dclk, sizeof c, cudaMemcpyDeviceToHost));
        if (c < best) best = c;
    }
    printf("%s\"%s\": %.1f", first ? "" : ", ", tag, (double)best / ((double)iters * KSTEPS));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    float* sink; CK(cudaMalloc(&sink, sizeof(float) * sms));
    long long* dclk; CK(cudaMalloc(&dclk, sizeof(long long)));
    printf("{\"gpu\": \"%s\", \"unit\": \"SM clocks per tcgen05.mma kind::tf32 K=8, all %d SMs busy\", ", p.name, sms);
    // K-major, new accumulator per 4-K-step product
    run<64, 64, 0, false, 4>(sms, sink, dclk, "m64n64_k_rot4", true);
    run<128, 64, 0, false, 4>(sms, sink, dclk, "m128n64_k_rot4", false);
    run<64, 128, 0, false, 4>(sms, sink, dclk, "m64n128_k_rot4", false);
    run<128, 128, 0, false, 4>(sms, sink, dclk, "m128n128_k_rot4", false);
    run<128, 128, 0, false, 8>(sms, sink, dclk, "m128n128_k_rot8", false);
    run<64, 256, 0, false, 4>(sms, sink, dclk, "m64n256_k_rot4", false);
    run<128, 256, 0, false, 4>(sms, sink, dclk, "m128n256_k_rot4", false);
    // same accumulator for everything (dependent chain)
    run<64, 64, 1, false, 4>(sms, sink, dclk, "m64n64_k_chain", false);
    run<128, 64, 1, false, 4>(sms, sink, dclk, "m128n64_k_chain", false);
    run<128, 128, 1, false, 4>(sms, sink, dclk, "m128n128_k_chain", false);
    // accumulator rotates every MMA
    run<64, 64, 2, false, 4>(sms, sink, dclk, "m64n64_k_rot1", false);
    run<128, 128, 2, false, 4>(sms, sink, dclk, "m128n128_k_rot1", false);
    // MN-major operands (32-byte-atom swizzle)
    run<64, 64, 0, true, 4>(sms, sink, dclk, "m64n64_mn_rot4", false);
    run<128, 64, 0, true, 4>(sms, sink, dclk, "m128n64_mn_rot4", false);
    run<128, 128, 0, true, 4>(sms, sink, dclk, "m128n128_mn_rot4", false);
    run<128, 256, 0, true, 4>(sms, sink, dclk, "m128n256_mn_rot4", false);
    // commits in the MMA stream; consumer warps polling mbarriers
    run<64, 64, 0, false, 4, 1>(sms, sink, dclk, "m64n64_k_rot4_commit_every_product", false);
    run<64, 64, 0, false, 4, 4>(sms, sink, dclk, "m64n64_k_rot4_commit_every_4_products", false);
    run<128, 128, 0, false, 8, 1>(sms, sink, dclk, "m128n128_k_rot8_commit_every_product", false);
    run<64, 64, 0, false, 4, 0, true, 512>(sms, sink, dclk, "m64n64_k_rot4_15_warps_polling", false);
    run<64, 64, 0, false, 4, 4, true, 512>(sms, sink, dclk, "m64n64_k_rot4_commit4_15_warps_polling", false);
    run<128, 128, 0, false, 8, 1, true, 512>(sms, sink, dclk, "m128n128_k_rot8_commit1_15_warps_polling", false);
    // fresh operand addresses per product (6 slots of 16 KiB per operand)
    run<64, 64, 0, false, 4, 0, false, 128, 6>(sms, sink, dclk, "m64n64_k_rot4_fresh_operands", false);
    run<64, 64, 0, false, 4, 4, true, 512, 6>(sms, sink, dclk, "m64n64_k_rot4_fresh_commit4_polling", false);
    run<64, 128, 0, false, 4, 0, false, 128, 6>(sms, sink, dclk, "m64n128_k_rot4_fresh_operands", false);
    run<128, 64, 0, false, 4, 0, false, 128, 6>(sms, sink, dclk, "m128n64_k_rot4_fresh_operands", false);
    run<128, 128, 0, false, 4, 0, false, 128, 4>(sms, sink, dclk, "m128n128_k_rot4_fresh_operands", false);
    run<128, 64, 0, false, 4, 2, true, 512, 6>(sms, sink, dclk, "m128n64_k_rot4_fresh_commit2_polling", false);
    printf("}\n");
    return 0;
}
