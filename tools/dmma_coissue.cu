// Does a warp issuing DMMA.8x8x4 back to back starve the other warp of its SM sub-partition?
// Warps 0-3 (one per sub-partition) run a DMMA loop (or nothing); warps 4-7 run a fixed integer/LDS instruction
// stream and report cycles per instruction.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256) k(double* out, long long* cyc, int mma_iters, int work_iters, int mode) {
    __shared__ int sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 4) {
        if (mma_iters == 0) return;
        double c0[16], c1[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { c0[i] = lane; c1[i] = i; }
        for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma(c0[i], c1[i], 1.0000001, 1e-9);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
        if (s == 12345.678) out[0] = s;
    } else {
        int a = lane, b = warp, c = 3, d = 7;
        double fa = lane, fb = 1.5;
        long long t0 = clock64();
        for (int it = 0; it < work_iters; ++it) {
            if (mode == 0) {           // independent integer ops (4 chains)
#pragma unroll
                for (int u = 0; u < 16; ++u) { a = a * 3 + 1; b = b ^ (b >> 3); c = c + d; d = d * 5 + c; }
            } else if (mode == 1) {    // LDS stream
#pragma unroll
                for (int u = 0; u < 16; ++u) { a = sm[(a + u) & 1023]; b = sm[(b + 2 * u) & 1023]; c += a; d ^= b; }
            } else {                   // plain FP64 ops (DADD / DSETP)
#pragma unroll
                for (int u = 0; u < 16; ++u) { fa = fa + fb; fb = fb * 1.0000001; if (fa > 1e300) c++; d += u; }
            }
        }
        long long t1 = clock64();
        if (lane == 0) cyc[blockIdx.x * 4 + (warp - 4)] = t1 - t0;
        if (a + b + c + d == 123456789 || fa == 1.2345) out[1] = a;
    }
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; long long* cyc; CK(cudaMalloc(&out, 16)); CK(cudaMalloc(&cyc, sms * 4 * 8));
    long long* h = (long long*)malloc(sms * 4 * 8);
    printf("{");
    for (int mode = 0; mode < 3; ++mode) {
        for (int with = 0; with < 2; ++with) {
            int work = 2000;
            k<<<sms, 256>>>(out, cyc, with ? 400000 : 0, work, mode);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, cyc, sms * 4 * 8, cudaMemcpyDeviceToHost));
            double tot = 0; for (int i = 0; i < sms * 4; ++i) tot += h[i];
            printf("\"mode%d_%s_cycles_per_iter16\": %.1f, ", mode, with ? "with_dmma" : "alone", tot / (sms * 4) / work);
        }
    }
    printf("\"sms\": %d}\n", sms);
    return 0;
}
