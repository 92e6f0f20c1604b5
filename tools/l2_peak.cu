// L2 -> SM bandwidth micro-benchmark for B200 (sm_100a): the roofline denominator of the gather kernels whose
// working set is L2-resident (K6c at 600 x 50 kb re-reads a 277 MB matrix of which the hot ~60 % stays in the 126 MB L2).
//   ldg   : every thread streams 128-bit read-only loads over a buffer that fits the L2
//   bulk  : one persistent CTA per SM, producer warps issue cp.async.bulk copies of `chunk` bytes into a shared-memory ring
//           (the same instruction, SASS UBLKCP, the re-score kernel uses), nothing consumes the data
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_peak l2_peak.cu ; run: ./l2_peak [buffer MiB]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ buf, size_t n16, int passes, unsigned* out) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n16; i += 4 * stride) {
            uint4 a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(buf + i));
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(buf + i + stride));
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(buf + i + 2 * stride));
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(buf + i + 3 * stride));
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// NW producer warps; every lane owns a 2-deep ring of `chunk`-byte slots and an mbarrier per slot
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) bulk_kernel(const unsigned char* __restrict__ buf, size_t bytes, int chunk, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);                  // [NW * 32][2]
    unsigned char* slots = sm + NW * 32 * 2 * 8;
    const int tid = threadIdx.x;
    for (int s = 0; s < 2; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[tid * 2 + s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t nchunks = bytes / chunk;
    size_t pos = ((size_t)blockIdx.x * blockDim.x + tid) * 977 % nchunks;   // scattered starts
    for (int it = 0; it < iters; ++it) {
        const int s = it & 1;
        if (it >= 2) {
            uint32_t ok = 0;
            const uint32_t par = (uint32_t)(((it - 2) >> 1) & 1);
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(smem_u32(&bars[tid * 2 + s])), "r"(par) : "memory");
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[tid * 2 + s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(slots + ((size_t)tid * 2 + s) * chunk)),
                     "l"(buf + pos * chunk), "r"(chunk), "r"(smem_u32(&bars[tid * 2 + s]))
                     : "memory");
        pos += 7919;
        if (pos >= nchunks) pos -= nchunks;
    }
    for (int it = iters; it < iters + 2; ++it) {                         // drain
        const int s = it & 1;
        if (it >= 2) {
            uint32_t ok = 0;
            const uint32_t par = (uint32_t)(((it - 2) >> 1) & 1);
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(smem_u32(&bars[tid * 2 + s])), "r"(par) : "memory");
        }
    }
}

int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? (size_t)atoi(argv[1]) : 48;
    const size_t bytes = mib << 20;
    unsigned char* buf;
    unsigned* out;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&out, 4));
    CK(cudaMemset(buf, 1, bytes));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("{\"device\": \"%s\", \"sms\": %d, \"buffer_mib\": %zu", pr.name, sms, mib);
    {
        const int passes = 40;
        ldg_kernel<<<sms * 4, 512>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, 2, out);      // warm the L2
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            ldg_kernel<<<sms * 4, 512>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, passes, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        printf(", \"ldg128_gbs\": %.1f", (double)bytes * passes / (best * 1e-3) / 1e9);
    }
    const int chunks[4] = {256, 512, 800, 1024};
    for (int ci = 0; ci < 4; ++ci) {
        const int chunk = chunks[ci];
        constexpr int NW = 4;
        const size_t smem = NW * 32 * 2 * 8 + (size_t)NW * 32 * 2 * chunk;
        if (smem > 227 * 1024) continue;
        CK(cudaFuncSetAttribute(bulk_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int iters = 4000;
        bulk_kernel<NW><<<sms, NW * 32, smem>>>(buf, bytes, chunk, 200);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            bulk_kernel<NW><<<sms, NW * 32, smem>>>(buf, bytes, chunk, iters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        printf(", \"bulk%d_gbs\": %.1f", chunk, (double)sms * NW * 32 * iters * chunk / (best * 1e-3) / 1e9);
    }
    printf("}\n");
    return 0;
}
