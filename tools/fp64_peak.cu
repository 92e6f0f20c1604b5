// FP64 peak micro-benchmark for B200 (sm_100a): DFMA chains vs DMMA.8x8x4 (mma.sync m8n8k4 f64).
// Gives the measured FP64 denominators for the newref distance kernel's roofline.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int TILES>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c0[TILES], c1[TILES];
#pragma unroll
    for (int i = 0; i < TILES; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

// DMMA with 4 A frags x 8 B frags register blocking (32 DMMAs per 12 operand regs), operands refreshed from smem
__global__ void __launch_bounds__(256) dmma_blocked_kernel(double* out, int iters) {
    __shared__ double sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 1e-3 * (i % 17);
    __syncthreads();
    double c0[32], c1[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { c0[i] = 0; c1[i] = 0; }
    int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        double a[4], b[8];
        int base = (it * 64) & 1023;
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sm[base + i * 32 + lane];
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = sm[1024 + ((base + i * 32 + lane) & 1023)];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma884(c0[i * 8 + j], c1[i * 8 + j], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    int iters = 1 << 14;
    for (int wps = 4; wps <= 16; wps *= 2) {   // warps per SM via blocks per SM (256 thr = 8 warps per block)
        int blocks = sms * wps / 8; if (blocks < sms) blocks = sms;
        int threads = (wps < 8) ? wps * 32 : 256;
        if (wps < 8) blocks = sms;
        {
            float ms = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 16 * (double)iters * blocks * threads;
            printf(", \"dfma16_w%d_tflops\": %.2f", wps, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 8 * (double)iters * blocks * (threads / 32);
            printf(", \"dmma8_w%d_tflops\": %.2f", wps, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 16 * (double)iters * blocks * (threads / 32);
            printf(", \"dmma16_w%d_tflops\": %.2f", wps, fl / ms * 1e-9);
        }
        {
            int it2 = iters / 4;
            float ms = time_ms([&] { dmma_blocked_kernel<<<blocks, threads>>>(out, it2); }, 5);
            double fl = 2.0 * 256 * 32 * (double)it2 * blocks * (threads / 32);
            printf(", \"dmma_blk_w%d_tflops\": %.2f", wps, fl / ms * 1e-9);
        }
    }
    // sustained (about 2 s each) at 8 warps/SM
    {
        int blocks = sms, threads = 256;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int n = 0; CK(cudaEventRecord(e0));
        float ms = 0;
        while (ms < 2000.f) { for (int i = 0; i < 20; ++i) dfma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); n += 20;
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); }
        printf(", \"dfma_sustained_tflops\": %.2f", 2.0 * 16 * (double)iters * blocks * threads * n / ms * 1e-9);
        n = 0; CK(cudaEventRecord(e0)); ms = 0;
        while (ms < 2000.f) { for (int i = 0; i < 20; ++i) dmma_blocked_kernel<<<blocks, threads>>>(out, iters / 4); n += 20;
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); }
        printf(", \"dmma_blk_sustained_tflops\": %.2f", 2.0 * 256 * 32 * (double)(iters / 4) * blocks * (threads / 32) * n / ms * 1e-9);
    }
    printf("}\n");
    return 0;
}
