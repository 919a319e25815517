// FP64 peak probe for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) shapes.
// Prints one JSON line; used to fill the FP64 roofline denominator (not in MEASURED_PEAKS.json).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 1 reg, B 1 reg, C 2 regs
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4: A 2 regs, B 1, C 4
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: A 4 regs, B 2, C 4
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: A 8 regs, B 4, C 4
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * 256 * sms * 16));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        double t;
        t = time_ms([&] { k_dfma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dfma_ilp8_bps%d_tflops\": %.2f", bps, 2.0 * 8 * iters * 256.0 * grid / t / 1e9);
        t = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dmma884_ilp8_bps%d_tflops\": %.2f", bps, 512.0 * 8 * iters * 8.0 * grid / t / 1e9);
        t = time_ms([&] { k_dmma1684<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dmma1684_ilp4_bps%d_tflops\": %.2f", bps, 1024.0 * 4 * iters * 8.0 * grid / t / 1e9);
        t = time_ms([&] { k_dmma1688<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dmma1688_ilp4_bps%d_tflops\": %.2f", bps, 2048.0 * 4 * iters * 8.0 * grid / t / 1e9);
        t = time_ms([&] { k_dmma16816<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dmma16816_ilp4_bps%d_tflops\": %.2f", bps, 4096.0 * 4 * iters * 8.0 * grid / t / 1e9);
    }
    // sustained (about 2 s) for the best candidate shapes
    {
        int grid = sms * 2;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 400; ++n) k_dmma884<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dmma884_sustained_tflops\": %.2f, \"sustained_ms\": %.0f", 512.0 * 8 * iters * 8.0 * grid * n / ms / 1e9, ms);
        CK(cudaEventRecord(e0));
        for (n = 0; n < 400; ++n) k_dfma<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dfma_sustained_tflops\": %.2f", 2.0 * 8 * iters * 256.0 * grid * n / ms / 1e9);
    }
    printf("}\n");
    return 0;
}
