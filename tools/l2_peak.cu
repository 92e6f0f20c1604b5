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
#include <cuda.h>

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

// gather4: one TMA request brings 4 rows x `C` doubles (tile::gather4 of a 2-D tensor map with a 1-row box)
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) gather4_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int C, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);                  // [NW * 32][2]
    const int slot = (4 * C * 8 + 127) / 128 * 128;
    unsigned char* slots = sm + ((NW * 32 * 2 * 8 + 127) / 128 * 128);
    const int tid = threadIdx.x;
    for (int s = 0; s < 2; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[tid * 2 + s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned pos = (blockIdx.x * blockDim.x + tid) * 977u;
    for (int it = 0; it < iters + 2; ++it) {
        const int s = it & 1;
        if (it >= 2) {
            uint32_t ok = 0;
            const uint32_t par = (uint32_t)(((it - 2) >> 1) & 1);
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(smem_u32(&bars[tid * 2 + s])), "r"(par) : "memory");
        }
        if (it >= iters) continue;
        const int r0 = (int)(pos % (unsigned)rows), r1 = (int)((pos * 3u + 1u) % (unsigned)rows), r2 = (int)((pos * 5u + 2u) % (unsigned)rows),
                  r3 = (int)((pos * 7u + 3u) % (unsigned)rows);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[tid * 2 + s])), "r"(4 * C * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(smem_u32(slots + ((size_t)tid * 2 + s) * slot)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
                       "r"(smem_u32(&bars[tid * 2 + s]))
                     : "memory");
        pos += 7919u;
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
    {   // gather4 requests: rows of 128 doubles, box of C columns x 1 row
        typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        const int rowlen = 128;
        const int nrows = (int)(bytes / (rowlen * 8));
        const int Cs[3] = {32, 64, 100};
        for (int ci = 0; ci < 3; ++ci) {
            const int C = Cs[ci];
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)rowlen, (cuuint64_t)nrows};
            cuuint64_t strides[1] = {(cuuint64_t)rowlen * 8};
            cuuint32_t box[2] = {(cuuint32_t)C, 1};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = reinterpret_cast<PFN>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf(", \"gather4_%d_err\": %d", C, (int)r); continue; }
            constexpr int NW = 1;
            const int slot = (4 * C * 8 + 127) / 128 * 128;
            const size_t smem = ((NW * 32 * 2 * 8 + 127) / 128 * 128) + (size_t)NW * 32 * 2 * slot;
            if (smem > 227 * 1024) continue;
            CK(cudaFuncSetAttribute(gather4_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int iters = 4000;
            gather4_kernel<NW><<<sms, NW * 32, smem>>>(tm, nrows, C, 100);
            CK(cudaDeviceSynchronize());
            float best = 1e30f;
            for (int rr = 0; rr < 5; ++rr) {
                CK(cudaEventRecord(e0));
                gather4_kernel<NW><<<sms, NW * 32, smem>>>(tm, nrows, C, iters);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                best = ms < best ? ms : best;
            }
            printf(", \"gather4_%dB_gbs\": %.1f, \"gather4_%dB_requests_per_us_per_sm\": %.2f", 4 * C * 8, (double)sms * NW * 32 * iters * 4.0 * C * 8 / (best * 1e-3) / 1e9,
                   4 * C * 8, (double)NW * 32 * iters / (best * 1e3));
        }
    }
    printf("}\n");
    return 0;
}
