// Micro-benchmark of K5's inner loop shape (per double-step: 2 + 16 LDS.128 fragment loads, 64 DMMA.8x8x4) with 1 or 2
// warps per SM sub-partition, to see how much of the FP64 tensor pipe a single warp can drive on its own.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void lds2(uint32_t addr, double& v0, double& v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
}

template <int MT, int NT>   // warp tile = MT*8 rows x NT*8 cols
__global__ void __launch_bounds__(256) loop_kernel(double* out, int chunks) {
    extern __shared__ __align__(1024) unsigned char smem[];
    double* sd = reinterpret_cast<double*>(smem);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sd[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q4 = lane & 3, pg = ((g & 1) << 2) | (g >> 1);
    const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t a_off = (uint32_t)((warp % (16 / MT)) * MT * 8 + pg) * 128u;
    const uint32_t b_off = 16384u + (uint32_t)pg * 128u;
    uint32_t sw[2];
    for (int h = 0; h < 2; ++h) sw[h] = (uint32_t)(((4 * h + q4) ^ pg) << 4);
    double acc[MT][NT][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
    for (int c = 0; c < chunks; ++c) {
        const uint32_t base = base0 + (uint32_t)(c & 0) * 32768u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double fa[MT][2];
#pragma unroll
            for (int m = 0; m < MT; ++m) lds2(base + a_off + m * 1024 + sw[h], fa[m][0], fa[m][1]);
#pragma unroll
            for (int half = 0; half < NT / 8; ++half) {
                double fb[8][2];
#pragma unroll
                for (int i = 0; i < 8; ++i) lds2(base + b_off + (half * 8 + i) * 1024 + sw[h], fb[i][0], fb[i][1]);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int m = 0; m < MT; ++m)
#pragma unroll
                        for (int i = 0; i < 8; ++i) dmma(acc[m][half * 8 + i][0], acc[m][half * 8 + i][1], fa[m][u], fb[i][u]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) s += acc[m][n][0] + acc[m][n][1];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    const int chunks = 20000;
    CK(cudaFuncSetAttribute(loop_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(loop_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    printf("{");
    for (int warps = 4; warps <= 8; warps += 4) {
        float ms = time_ms([&] { loop_kernel<2, 16><<<sms, warps * 32, 65536>>>(out, chunks); });
        double fl = 2.0 * 256 * 128 * (double)chunks * sms * warps;
        printf("\"tile16x128_w%d_tflops\": %.2f, ", warps, fl / ms * 1e-9);
        ms = time_ms([&] { loop_kernel<4, 8><<<sms, warps * 32, 65536>>>(out, chunks); });
        printf("\"tile32x64_w%d_tflops\": %.2f, ", warps, fl / ms * 1e-9);
    }
    printf("\"sms\": %d}\n", sms);
    return 0;
}
