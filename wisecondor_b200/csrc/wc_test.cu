// wisecondor_b200 - batched `test` front half on B200 (sm_100a): sample preparation and within-sample z-scores.
//
// Replaces, for a batch of B samples at once,
//   toNumpyRefFormat + applyPCA   (/root/reference/wisetools.py:267-278, 104-113)            -> K7 wc_test_prep
//   trySample / repeatTest         (/root/reference/wisetools.py:407-448)                     -> K8 wc_zscore_batch
// and, once per reference, the `index[distances[i] < cutoff]` selection and the other-chromosome concatenation
// trySample rebuilds for every chromosome (wisetools.py:420-424)                               -> wc_test_table.
//
// Data layout in HBM: the corrected test values live as T[bin][sample] ("sample-minor", leading dimension ldb = B
// rounded up to 32): a warp owns 32 consecutive samples of one target bin, so every gather of a reference bin is
// one 256-byte coalesced read shared by the warp, and the whole working set of a sample tile (N x 256 B) stays in
// the 126 MB L2 across the k gathers per bin.  Marked (aberrant) bins are -1 in a working copy, exactly as in the
// reference (wisetools.py:446).  Arithmetic follows numpy's operation order (wc_numpy_order.cuh), so z, r and the
// per-sample average sigma are bit-identical to the reference's.
#include "wc_common.cuh"
#include "wc_numpy_order.cuh"

namespace {

constexpr int ZS_WARPS = 4;             // max warps per CTA of the z-score kernel (each warp: 32 samples x 1 bin at a time)
constexpr int ZS_BINS_PER_CTA = 64;

// ---------------------------------------------------------------------------------------------------------
// gather table: table[i][0..count[i]) = global masked-bin ids of bin i's usable reference bins, in stored order
// ---------------------------------------------------------------------------------------------------------
__global__ void wc_table_kernel(const int* __restrict__ indexes, const double* __restrict__ distances, int N, int k,
                                const int* __restrict__ row_cs, const int* __restrict__ row_ce, double cutoff,
                                int* __restrict__ table, int* __restrict__ count) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N) return;
    const int cs = row_cs[warp], ce = row_ce[warp];
    const int nother = N - (ce - cs);
    int w = 0;
    for (int base = 0; base < k; base += 32) {
        const int m = base + lane;
        bool keep = false;
        int g = 0;
        if (m < k) {
            keep = distances[(size_t)warp * k + m] < cutoff;          // wisetools.py:424 (NaN never passes)
            int j = indexes[(size_t)warp * k + m];
            if (j < 0) j += nother;                                    // numpy's negative index wraps (filler -1)
            g = j >= cs ? j + (ce - cs) : j;                           // position in chromData -> global bin
            if (j < 0 || j >= nother) keep = false;                    // numpy would raise IndexError; never stored
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) table[(size_t)warp * k + w + __popc(bal & ((1u << lane) - 1u))] = g;
        w += __popc(bal);
    }
    if (lane == 0) count[warp] = w;
}

// ---------------------------------------------------------------------------------------------------------
// K7: sample preparation
// ---------------------------------------------------------------------------------------------------------
__global__ void wc_totals_kernel(const int* __restrict__ counts, int Nraw, double* __restrict__ totals) {
    __shared__ long long red[8];
    const int b = blockIdx.x;
    long long acc = 0;
    for (int i = threadIdx.x; i < Nraw; i += blockDim.x) acc += counts[(size_t)b * Nraw + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        totals[b] = (double)t;              // integer-valued: the order of the reference's float sum is irrelevant
    }
}

// Value of masked bin n of sample b: counts are divided by the sample total (wisetools.py:275); float input is taken
// as already normalised (the applyPCA-only entry point).
template <class TIn>
__device__ __forceinline__ double prep_value(const TIn* in, int ld_in, const int* map, const double* totals, int b, int n) {
    const TIn raw = in[(size_t)b * ld_in + (map ? map[n] : n)];
    return totals ? (double)raw / totals[b] : (double)raw;
}

// proj[b][j] = sum_n (x_n - mean_n) * C[j][n]        (pca.transform, wisetools.py:109)
template <class TIn>
__global__ void wc_project_kernel(const TIn* __restrict__ in, int ld_in, const int* __restrict__ map, int N,
                                  const double* __restrict__ totals, const double* __restrict__ mean,
                                  const double* __restrict__ comps, int ncomp, double* __restrict__ proj) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    for (int j = 0; j < ncomp; ++j) {
        double acc = 0.0;
        for (int n = threadIdx.x; n < N; n += blockDim.x)
            acc = fma(prep_value(in, ld_in, map, totals, b, n) - mean[n], comps[(size_t)j * N + n], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
            proj[(size_t)b * ncomp + j] = t;
        }
        __syncthreads();
    }
}

// T[n][b] = x / (proj . C[:, n] + mean_n)              (wisetools.py:111-113), transposed to sample-minor
template <class TIn>
__global__ void wc_residual_kernel(const TIn* __restrict__ in, int ld_in, const int* __restrict__ map, int N, int B,
                                   int ldb, const double* __restrict__ totals, const double* __restrict__ mean,
                                   const double* __restrict__ comps, int ncomp, const double* __restrict__ proj,
                                   double* __restrict__ T) {
    __shared__ double tile[32][33];
    const int n0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    {
        const int n = n0 + threadIdx.x, b = b0 + threadIdx.y;
        double v = 1.0;                      // padding samples: harmless finite values
        if (n < N && b < B) {
            const double x = prep_value(in, ld_in, map, totals, b, n);
            if (ncomp > 0) {
                double recon = 0.0;
                for (int j = 0; j < ncomp; ++j) recon = fma(proj[(size_t)b * ncomp + j], comps[(size_t)j * N + n], recon);
                v = x / (recon + mean[n]);
            } else {
                v = x;
            }
        }
        tile[threadIdx.x][threadIdx.y] = v;
    }
    __syncthreads();
    const int n = n0 + threadIdx.y, b = b0 + threadIdx.x;
    if (n < N && b < ldb) T[(size_t)n * ldb + b] = tile[threadIdx.y][threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------------
// K8: one z-score pass over all (bin, sample) pairs
// ---------------------------------------------------------------------------------------------------------
struct ZArgs {
    const double* test;      // [N][ldb] numerators (never change)
    const double* copy;      // [N][ldb] working copy: marked bins hold -1
    const int* table;        // [N][k]
    const int* count;        // [N]
    int N, B, ldb, k;
    double* z;               // [N][ldb]
    double* r;
    int* refsz;
    double* sd;
    const int* tile_active;  // [ntiles] 0 = nothing changed for these 32 samples since the previous pass: skip
};

template <bool DEV>
__device__ __forceinline__ double zs_leaf(const double* p, int n, double mean) {
    return np_sum_leaf([&](int i) {
        const double v = p[i * 32];
        if (!DEV) return v;
        const double d = __dsub_rn(v, mean);
        return __dmul_rn(d, d);
    }, n);
}
// n <= 512: numpy splits at most twice
template <bool DEV>
__device__ __noinline__ double zs_sum_large(const double* p, int n, double mean) {
    auto half = [&](const double* q, int len) {
        if (len <= 128) return zs_leaf<DEV>(q, len, mean);
        int h = len / 2;
        h -= h & 7;
        return __dadd_rn(zs_leaf<DEV>(q, h, mean), zs_leaf<DEV>(q + (size_t)h * 32, len - h, mean));
    };
    int h = n / 2;
    h -= h & 7;
    return __dadd_rn(half(p, h), half(p + (size_t)h * 32, n - h));
}

__global__ void __launch_bounds__(ZS_WARPS * 32) wc_zscore_kernel(const ZArgs a) {
    extern __shared__ __align__(16) unsigned char zs_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    const int tile = blockIdx.y;
    if (a.tile_active != nullptr && a.tile_active[tile] == 0) return;
    const int s = tile * 32 + lane;                      // < ldb by construction
    double* buf = reinterpret_cast<double*>(zs_raw) + (size_t)warp * a.k * 32 + lane;   // entry p at buf[p * 32]
    int* sidx = reinterpret_cast<int*>(reinterpret_cast<double*>(zs_raw) + (size_t)nwarps * a.k * 32) + warp * a.k;
    const double* cp = a.copy + s;
    const int bin_end = min(a.N, (int)(blockIdx.x + 1) * ZS_BINS_PER_CTA);
    for (int i = blockIdx.x * ZS_BINS_PER_CTA + warp; i < bin_end; i += nwarps) {
        const int cnt = a.count[i];
        __syncwarp();
        for (int m = lane; m < cnt; m += 32) sidx[m] = a.table[(size_t)i * a.k + m];
        __syncwarp();
        // gather the reference values of this bin from the same sample, dropping marked (negative) ones
        int p = 0;
        for (int m0 = 0; m0 < cnt; m0 += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int m = min(m0 + u, cnt - 1);
                v[u] = cp[(size_t)sidx[m] * a.ldb];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (m0 + u < cnt && v[u] >= 0.0) {       // wisetools.py:425
                    buf[p * 32] = v[u];
                    ++p;
                }
            }
        }
        const int n = p;
        const double sum = n <= 128 ? zs_leaf<false>(buf, n, 0.0) : zs_sum_large<false>(buf, n, 0.0);
        const double mean = __ddiv_rn(sum, (double)n);                       // np_mean (wisetools.py:426)
        const double ssq = n <= 128 ? zs_leaf<true>(buf, n, mean) : zs_sum_large<true>(buf, n, mean);
        const double sd = sqrt(__ddiv_rn(ssq, (double)n));                   // np_std, ddof 0 (wisetools.py:427)
        const double x = a.test[(size_t)i * a.ldb + s];
        const size_t o = (size_t)i * a.ldb + s;
        a.z[o] = __ddiv_rn(__dsub_rn(x, mean), sd);                          // wisetools.py:431
        a.r[o] = __ddiv_rn(x, mean);                                         // wisetools.py:432
        a.refsz[o] = n;
        a.sd[o] = sd;
    }
}

// testCopy[abs(z) >= threshold] = -1 (wisetools.py:446), applied between passes; flags the sample tiles that changed
__global__ void wc_mark_kernel(const double* __restrict__ z, double* __restrict__ copy, int N, int B, int ldb, double thr,
                               int* __restrict__ next_active) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)N * ldb;
    bool changed = false;
    int s = 0;
    if (idx < total) {
        s = (int)(idx % ldb);
        if (s < B && fabs(z[idx]) >= thr && copy[idx] != -1.0) {
            copy[idx] = -1.0;
            changed = true;
        }
    }
    if (__any_sync(0xffffffffu, changed) && (threadIdx.x & 31) == 0) next_active[s >> 5] = 1;
}

// stdDevSum / stdDevNum accumulated bin by bin like the reference's Python loop (wisetools.py:428-430, 435)
__global__ void wc_sigma_kernel(const double* __restrict__ sd, int N, int B, int ldb, double* __restrict__ asdef) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    double sum = 0.0;
    int num = 0;
    int i = 0;
    for (; i + 8 <= N; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = sd[(size_t)(i + u) * ldb + s];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (!isnan(v[u])) { sum = __dadd_rn(sum, v[u]); ++num; }
    }
    for (; i < N; ++i) {
        const double v = sd[(size_t)i * ldb + s];
        if (!isnan(v)) { sum = __dadd_rn(sum, v); ++num; }
    }
    asdef[s] = __ddiv_rn(sum, (double)num);
}

// [N][ldb] sample-minor -> [B][N] sample-major (what the host copies out and what the segmentation reads)
template <class T>
__global__ void wc_transpose_kernel(const T* __restrict__ in, int N, int B, int ldb, T* __restrict__ out) {
    __shared__ T tile[32][33];
    const int n0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    {
        const int n = n0 + threadIdx.y, b = b0 + threadIdx.x;
        if (n < N && b < ldb) tile[threadIdx.y][threadIdx.x] = in[(size_t)n * ldb + b];
    }
    __syncthreads();
    const int n = n0 + threadIdx.x, b = b0 + threadIdx.y;
    if (n < N && b < B) out[(size_t)b * N + n] = tile[threadIdx.x][threadIdx.y];
}

int upload_row_ranges(wc_ctx* ctx, int N, const int* chrom_bins_h, int nchrom, int slot_cs, int slot_ce, int** cs_d,
                      int** ce_d, cudaStream_t stream) {
    std::vector<int> row_cs(N), row_ce(N);
    int pos = 0;
    for (int c = 0; c < nchrom; ++c) {
        for (int i = 0; i < chrom_bins_h[c]; ++i) { row_cs[pos + i] = pos; row_ce[pos + i] = pos + chrom_bins_h[c]; }
        pos += chrom_bins_h[c];
    }
    int rc;
    if ((rc = wc_reserve(ctx, slot_cs, (size_t)N * sizeof(int), (void**)cs_d))) return rc;
    if ((rc = wc_reserve(ctx, slot_ce, (size_t)N * sizeof(int), (void**)ce_d))) return rc;
    WC_CUDA(cudaMemcpyAsync(*cs_d, row_cs.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(*ce_d, row_ce.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaStreamSynchronize(stream));      // the host vectors die with this frame
    return WC_OK;
}

}  // namespace

extern "C" int wc_test_table(wc_ctx* ctx, const int32_t* indexes_d, const double* distances_d, int N, int k,
                             const int* chrom_bins_h, int nchrom, double cutoff, int32_t* table_d, int32_t* count_d,
                             void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && indexes_d != nullptr && distances_d != nullptr && chrom_bins_h != nullptr);
    WC_CHECK_ARG(table_d != nullptr && count_d != nullptr);
    WC_CHECK_ARG(N > 0 && k > 0 && nchrom > 0);
    long long tot = 0;
    for (int c = 0; c < nchrom; ++c) { WC_CHECK_ARG(chrom_bins_h[c] >= 0); tot += chrom_bins_h[c]; }
    WC_CHECK_ARG(tot == N);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    int* cs_d; int* ce_d;
    int rc;
    if ((rc = upload_row_ranges(ctx, N, chrom_bins_h, nchrom, SLOT_ROWCS, SLOT_ROWCE, &cs_d, &ce_d, stream))) return rc;
    const int blocks = (int)(((size_t)N * 32 + 255) / 256);
    wc_table_kernel<<<blocks, 256, 0, stream>>>(indexes_d, distances_d, N, k, cs_d, ce_d, cutoff, table_d, count_d);
    WC_CUDA(cudaGetLastError());
    return WC_OK;
}

template <class TIn>
static int prep_common(wc_ctx* ctx, const TIn* in_d, int ld_in, const int* map_d, bool normalise, int B, int N,
                       const double* pca_mean_d, const double* pca_components_d, int ncomp, double* test_d, int ldb,
                       cudaStream_t stream) {
    double* totals = nullptr; double* proj;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_T_PROJ, (size_t)B * std::max(ncomp, 1) * sizeof(double), (void**)&proj))) return rc;
    WC_CUDA(cudaEventRecord(ctx->ev[12], stream));
    if (normalise) {
        if ((rc = wc_reserve(ctx, SLOT_T_TOTALS, (size_t)B * sizeof(double), (void**)&totals))) return rc;
        wc_totals_kernel<<<B, 256, 0, stream>>>(reinterpret_cast<const int*>(in_d), ld_in, totals);
    }
    if (ncomp > 0)
        wc_project_kernel<TIn><<<B, 256, 0, stream>>>(in_d, ld_in, map_d, N, totals, pca_mean_d, pca_components_d, ncomp, proj);
    dim3 grid((N + 31) / 32, ldb / 32);
    wc_residual_kernel<TIn><<<grid, dim3(32, 32), 0, stream>>>(in_d, ld_in, map_d, N, B, ldb, totals, pca_mean_d,
                                                               pca_components_d, ncomp, proj, test_d);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[13], stream));
    ctx->timed_mask |= 1u << 6;
    return WC_OK;
}

extern "C" int wc_test_prep(wc_ctx* ctx, const int32_t* counts_d, int B, int Nraw, const int32_t* masked_raw_d, int N,
                            const double* pca_mean_d, const double* pca_components_d, int ncomp, double* test_d, int ldb,
                            void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && counts_d != nullptr && masked_raw_d != nullptr && test_d != nullptr);
    WC_CHECK_ARG(ncomp == 0 || (pca_mean_d != nullptr && pca_components_d != nullptr));
    WC_CHECK_ARG(B > 0 && Nraw > 0 && N > 0 && N <= Nraw && ncomp >= 0 && ldb >= B && ldb % 32 == 0);
    WC_CUDA(cudaSetDevice(ctx->device));
    return prep_common<int>(ctx, counts_d, Nraw, masked_raw_d, true, B, N, pca_mean_d, pca_components_d, ncomp, test_d,
                            ldb, static_cast<cudaStream_t>(stream_v));
}

extern "C" int wc_apply_pca(wc_ctx* ctx, const double* x_d, int B, int N, const double* pca_mean_d,
                            const double* pca_components_d, int ncomp, double* test_d, int ldb, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && x_d != nullptr && test_d != nullptr && pca_mean_d != nullptr && pca_components_d != nullptr);
    WC_CHECK_ARG(B > 0 && N > 0 && ncomp > 0 && ldb >= B && ldb % 32 == 0);
    WC_CUDA(cudaSetDevice(ctx->device));
    return prep_common<double>(ctx, x_d, N, nullptr, false, B, N, pca_mean_d, pca_components_d, ncomp, test_d, ldb,
                               static_cast<cudaStream_t>(stream_v));
}

extern "C" int wc_zscore_batch(wc_ctx* ctx, const double* test_d, const double* copy_init_d, int N, int B, int ldb, const int32_t* table_d,
                               const int32_t* count_d, int k, double z_threshold, int repeats, double* z_d, double* r_d,
                               int32_t* refsizes_d, double* asdef_d, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && test_d != nullptr && table_d != nullptr && count_d != nullptr);
    WC_CHECK_ARG(z_d != nullptr && r_d != nullptr && refsizes_d != nullptr && asdef_d != nullptr);
    WC_CHECK_ARG(N > 0 && B > 0 && ldb >= B && ldb % 32 == 0 && k >= 1 && k <= 512 && repeats >= 1);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    const int ntiles = ldb / 32;
    const size_t elems = (size_t)N * ldb;
    double* copy; double* zt; double* rt; int* nt; double* sd; int* flags;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_T_COPY, elems * sizeof(double), (void**)&copy))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_T_ZT, elems * sizeof(double), (void**)&zt))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_T_RT, elems * sizeof(double), (void**)&rt))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_T_NT, elems * sizeof(int), (void**)&nt))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_T_SD, elems * sizeof(double), (void**)&sd))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_T_FLAGS, (size_t)(repeats + 1) * ntiles * sizeof(int), (void**)&flags))) return rc;
    WC_CUDA(cudaEventRecord(ctx->ev[8], stream));
    WC_CUDA(cudaMemcpyAsync(copy, copy_init_d ? copy_init_d : test_d, elems * sizeof(double), cudaMemcpyDeviceToDevice,
                            stream));                                                                   // wisetools.py:442
    WC_CUDA(cudaMemsetAsync(flags, 0, (size_t)(repeats + 1) * ntiles * sizeof(int), stream));
    const size_t per_warp = (size_t)k * 32 * sizeof(double) + (size_t)k * sizeof(int);
    int warps = (int)std::min<size_t>(ZS_WARPS, (size_t)(227 * 1024) / per_warp);
    if (warps < 1) { wc_set_error("z-score kernel: refsize %d needs %zu bytes of shared memory per warp", k, per_warp); return WC_ERR_ARG; }
    if (warps == 3) warps = 2;
    const size_t smem = (size_t)warps * per_warp;
    WC_CUDA(cudaFuncSetAttribute(wc_zscore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ZArgs a;
    a.test = test_d; a.copy = copy; a.table = table_d; a.count = count_d; a.N = N; a.B = B; a.ldb = ldb; a.k = k;
    a.z = zt; a.r = rt; a.refsz = nt; a.sd = sd;
    long long launches = 0;
    for (int rep = 0; rep < repeats; ++rep) {
        a.tile_active = rep == 0 ? nullptr : flags + (size_t)rep * ntiles;
        dim3 grid((N + ZS_BINS_PER_CTA - 1) / ZS_BINS_PER_CTA, ntiles);
        wc_zscore_kernel<<<grid, warps * 32, smem, stream>>>(a);
        ++launches;
        if (rep + 1 < repeats) {     // the marks of the last pass are never read (wisetools.py:443-448)
            wc_mark_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, stream>>>(zt, copy, N, B, ldb, z_threshold,
                                                                                  flags + (size_t)(rep + 1) * ntiles);
            ++launches;
        }
    }
    wc_sigma_kernel<<<(B + 63) / 64, 64, 0, stream>>>(sd, N, B, ldb, asdef_d);
    dim3 tgrid((N + 31) / 32, ldb / 32);
    wc_transpose_kernel<double><<<tgrid, dim3(32, 32), 0, stream>>>(zt, N, B, ldb, z_d);
    wc_transpose_kernel<double><<<tgrid, dim3(32, 32), 0, stream>>>(rt, N, B, ldb, r_d);
    wc_transpose_kernel<int><<<tgrid, dim3(32, 32), 0, stream>>>(nt, N, B, ldb, refsizes_d);
    launches += 4;
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[9], stream));
    ctx->timed_mask |= 1u << 4;
    ctx->counter[5] = launches;
    return WC_OK;
}
