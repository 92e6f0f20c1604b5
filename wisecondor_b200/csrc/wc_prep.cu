// wisecondor_b200 - newref preparation on B200 (sm_100a): per-sample normalisation, nonzero-bin mask, 3-component PCA
// residual.
//
// Replaces toNumpyArray (/root/reference/wisetools.py:240-264) and trainPCA (/root/reference/wisetools.py:89-101) as
// called from toolNewrefPrep (/root/reference/wisecondor.py:91-96).
//
//   K1  wc_anycount_kernel / wc_residual-style transpose     mask = (sum over samples of the normalised bin) > 0, which
//        for non-negative counts is "any sample has a read there"; maskedData[bin][sample] = count / sample total   (HBM)
//   K2  wc_binmean_kernel       pca.mean_: per-bin mean over samples in numpy's pairwise order (bit-identical)       (HBM)
//       wc_gram_kernel          G = Xc^T Xc (S x S, contraction over all N bins) in fp64, centring fused into the tile
//                               load, split over bin slabs with a fixed-order reduction (deterministic)              (FP64)
//       -- host: the top-ncomp eigenpairs of the S x S Gram matrix (LAPACK through numpy/scipy; S = 600 or 2000) --
//       wc_components_kernel    V_j = Xc u_j / sigma_j: the right singular vectors, i.e. pca.components_           (HBM)
//   K3  wc_transform_kernel     t[s][j] = sum_n Xc[n][s] V[j][n]                       (pca.transform)               (HBM)
//       wc_correct_kernel       corrected = X / (t V + mean)      (inverse_transform + wisetools.py:96), bin-major    (HBM)
//
// The reference obtains V from a LAPACK SVD of the S x N centred matrix (scikit-learn PCA, svd_solver 'full' - see
// DESIGN.md for why that solver is the parity target); the Gram route gives the same subspace to ~1e-14 and keeps all
// N-sized work on the device.  Component signs follow scikit-learn's svd_flip (largest |entry| of each component
// positive) and are fixed on the host; they do not affect the corrected data.
#include "wc_common.cuh"
#include "wc_numpy_order.cuh"

namespace {

constexpr int GT = 64;          // Gram tile edge
constexpr int GK = 16;          // bins per shared-memory stage

// ---- K1 ---------------------------------------------------------------------------------------------------------
__global__ void wc_totals64_kernel(const int* __restrict__ counts, int Nraw, double* __restrict__ totals) {
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
        totals[b] = (double)t;
    }
}

__global__ void wc_anycount_kernel(const int* __restrict__ counts, int S, int Nraw, unsigned char* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Nraw) return;
    int any = 0;
    for (int s = 0; s < S; ++s) any |= counts[(size_t)s * Nraw + i] > 0;
    mask[i] = (unsigned char)any;
}

// maskedData[n][s] = counts[s][raw_n] / totals[s]   (wisetools.py:255-256, 261), transposed to bin-major
__global__ void wc_normalize_kernel(const int* __restrict__ counts, int S, int Nraw, const int* __restrict__ masked_raw,
                                    int N, const double* __restrict__ totals, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int n0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    {
        const int n = n0 + threadIdx.x, s = s0 + threadIdx.y;
        if (n < N && s < S) tile[threadIdx.x][threadIdx.y] = (double)counts[(size_t)s * Nraw + masked_raw[n]] / totals[s];
    }
    __syncthreads();
    const int n = n0 + threadIdx.y, s = s0 + threadIdx.x;
    if (n < N && s < S) out[(size_t)n * S + s] = tile[threadIdx.y][threadIdx.x];
}

// ---- K2 ---------------------------------------------------------------------------------------------------------
// np.mean(X, axis=0) of the (samples x bins) view of a bin-major matrix reduces each bin's contiguous run of S
// values with numpy's pairwise sum.
__global__ void wc_binmean_kernel(const double* __restrict__ X, int N, int S, double* __restrict__ mean) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double* row = X + (size_t)n * S;
    mean[n] = __ddiv_rn(np_sum_thread([&](int i) { return row[i]; }, S), (double)S);
}

// partial[split][S][S] tile (ti <= tj) over bins [n0, n1): 16 x 16 threads, 4 x 4 outputs each
__global__ void __launch_bounds__(256) wc_gram_kernel(const double* __restrict__ X, const double* __restrict__ mean, int N,
                                                      int S, int bins_per_split, double* __restrict__ partial) {
    __shared__ double As[GK][GT + 1], Bs[GK][GT + 1];
    // blockIdx.x enumerates the upper-triangular tile pairs
    const int ntile = (S + GT - 1) / GT;
    int ti = 0, rem = blockIdx.x;
    while (rem >= ntile - ti) { rem -= ntile - ti; ++ti; }
    const int tj = ti + rem;
    const int split = blockIdx.y;
    const int n0 = split * bins_per_split, n1 = min(N, n0 + bins_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int nb = n0; nb < n1; nb += GK) {
        for (int e = threadIdx.x; e < GK * GT; e += 256) {
            const int kk = e / GT, c = e % GT;
            const int n = nb + kk;
            double a = 0.0, b = 0.0;
            if (n < n1) {
                const double m = mean[n];
                const int sa = ti * GT + c, sb = tj * GT + c;
                if (sa < S) a = X[(size_t)n * S + sa] - m;
                if (sb < S) b = X[(size_t)n * S + sb] - m;
            }
            As[kk][c] = a;
            Bs[kk][c] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double* out = partial + (size_t)split * S * S;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ti * GT + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = tj * GT + tx * 4 + j;
            if (r < S && c < S) out[(size_t)r * S + c] = acc[i][j];
        }
    }
}

// G = sum over splits in a fixed order; mirror the upper tiles into the lower triangle
__global__ void wc_gram_reduce_kernel(const double* __restrict__ partial, int S, int nsplit, double* __restrict__ G) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)S * S) return;
    int r = (int)(idx / S), c = (int)(idx % S);
    if (r / GT > c / GT) { const int t = r; r = c; c = t; }      // tile below the diagonal: read its transpose
    double acc = 0.0;
    for (int k = 0; k < nsplit; ++k) acc += partial[(size_t)k * S * S + (size_t)r * S + c];
    G[idx] = acc;
}

// V[j][n] = sum_s Xc[n][s] U[s][j] / sigma_j : one warp per bin
__global__ void wc_components_kernel(const double* __restrict__ X, const double* __restrict__ mean, int N, int S,
                                     const double* __restrict__ U, const double* __restrict__ inv_sigma, int ncomp,
                                     double* __restrict__ V) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const double m = mean[n];
    const double* row = X + (size_t)n * S;
    for (int j = 0; j < ncomp; ++j) {
        double acc = 0.0;
        for (int s = lane; s < S; s += 32) acc = fma(row[s] - m, U[(size_t)s * ncomp + j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) V[(size_t)j * N + n] = acc * inv_sigma[j];
    }
}

// ---- K3 ---------------------------------------------------------------------------------------------------------
// partial_t[chunk][j][s] = sum over the chunk's bins of Xc[n][s] V[j][n]; thread = sample
__global__ void wc_transform_kernel(const double* __restrict__ X, const double* __restrict__ mean, int N, int S,
                                    const double* __restrict__ V, int ncomp, int bins_per_chunk,
                                    double* __restrict__ partial) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    const int n0 = chunk * bins_per_chunk, n1 = min(N, n0 + bins_per_chunk);
    if (s >= S) return;
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0;
    for (int n = n0; n < n1; ++n) {
        const double xc = X[(size_t)n * S + s] - mean[n];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < ncomp) acc[j] = fma(xc, V[(size_t)j * N + n], acc[j]);
    }
    for (int j = 0; j < ncomp; ++j) partial[((size_t)chunk * ncomp + j) * S + s] = acc[j];
}

__global__ void wc_transform_reduce_kernel(const double* __restrict__ partial, int S, int ncomp, int nchunk,
                                           double* __restrict__ T) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * ncomp) return;
    double acc = 0.0;
    for (int c = 0; c < nchunk; ++c) acc += partial[(size_t)c * ncomp * S + idx];
    T[idx] = acc;                        // T[j][s]
}

// corrected[n][s] = X[n][s] / (sum_j T[j][s] V[j][n] + mean[n])
__global__ void wc_correct_kernel(const double* __restrict__ X, const double* __restrict__ mean, int N, int S,
                                  const double* __restrict__ V, const double* __restrict__ T, int ncomp,
                                  double* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)N * S) return;
    const int n = (int)(idx / S), s = (int)(idx % S);
    double recon = 0.0;
    for (int j = 0; j < ncomp; ++j) recon = fma(T[(size_t)j * S + s], V[(size_t)j * N + n], recon);
    out[idx] = X[idx] / (recon + mean[n]);
}

enum { SLOT_P_TOTALS = SLOT_P_FIRST, SLOT_P_PARTIAL, SLOT_P_T, SLOT_P_ISIG };

}  // namespace

extern "C" int wc_newref_mask(wc_ctx* ctx, const int32_t* counts_d, int S, int Nraw, uint8_t* mask_d, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && counts_d != nullptr && mask_d != nullptr && S > 0 && Nraw > 0);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    wc_anycount_kernel<<<(Nraw + 255) / 256, 256, 0, stream>>>(counts_d, S, Nraw, mask_d);
    WC_CUDA(cudaGetLastError());
    return WC_OK;
}

extern "C" int wc_newref_normalize(wc_ctx* ctx, const int32_t* counts_d, int S, int Nraw, const int32_t* masked_raw_d,
                                   int N, double* masked_d, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && counts_d != nullptr && masked_raw_d != nullptr && masked_d != nullptr);
    WC_CHECK_ARG(S > 0 && Nraw > 0 && N > 0 && N <= Nraw);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    double* totals;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_P_TOTALS, (size_t)S * sizeof(double), (void**)&totals))) return rc;
    wc_totals64_kernel<<<S, 256, 0, stream>>>(counts_d, Nraw, totals);
    dim3 grid((N + 31) / 32, (S + 31) / 32);
    wc_normalize_kernel<<<grid, dim3(32, 32), 0, stream>>>(counts_d, S, Nraw, masked_raw_d, N, totals, masked_d);
    WC_CUDA(cudaGetLastError());
    return WC_OK;
}

extern "C" int wc_pca_gram(wc_ctx* ctx, const double* masked_d, int N, int S, double* mean_d, double* gram_d,
                           void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && masked_d != nullptr && mean_d != nullptr && gram_d != nullptr && N > 0 && S > 0);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    const int ntile = (S + GT - 1) / GT;
    const int pairs = ntile * (ntile + 1) / 2;
    int nsplit = std::max(1, std::min(64, (4 * ctx->sm_count + pairs - 1) / pairs));
    int bins_per_split = ((N + nsplit - 1) / nsplit + GK - 1) / GK * GK;
    nsplit = (N + bins_per_split - 1) / bins_per_split;
    double* partial;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_P_PARTIAL, (size_t)nsplit * S * S * sizeof(double), (void**)&partial))) return rc;
    WC_CUDA(cudaEventRecord(ctx->ev[14], stream));
    wc_binmean_kernel<<<(N + 127) / 128, 128, 0, stream>>>(masked_d, N, S, mean_d);
    wc_gram_kernel<<<dim3(pairs, nsplit), 256, 0, stream>>>(masked_d, mean_d, N, S, bins_per_split, partial);
    wc_gram_reduce_kernel<<<(unsigned)(((size_t)S * S + 255) / 256), 256, 0, stream>>>(partial, S, nsplit, gram_d);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[15], stream));
    ctx->timed_mask |= 1u << 7;
    ctx->counter[7] = 3;
    return WC_OK;
}

extern "C" int wc_pca_apply(wc_ctx* ctx, const double* masked_d, int N, int S, const double* mean_d,
                            const double* eigvec_d, const double* sigma_h, int ncomp, double* components_d,
                            double* corrected_d, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && masked_d != nullptr && mean_d != nullptr && eigvec_d != nullptr && sigma_h != nullptr);
    WC_CHECK_ARG(components_d != nullptr && corrected_d != nullptr && N > 0 && S > 0 && ncomp >= 1 && ncomp <= 8);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    double isig[8];
    for (int j = 0; j < ncomp; ++j) {
        WC_CHECK_ARG(sigma_h[j] > 0.0);
        isig[j] = 1.0 / sigma_h[j];
    }
    const int bins_per_chunk = 256;
    const int nchunk = (N + bins_per_chunk - 1) / bins_per_chunk;
    double* isig_d; double* partial; double* T;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_P_ISIG, 8 * sizeof(double), (void**)&isig_d))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_P_PARTIAL, (size_t)nchunk * ncomp * S * sizeof(double), (void**)&partial))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_P_T, (size_t)ncomp * S * sizeof(double), (void**)&T))) return rc;
    WC_CUDA(cudaMemcpyAsync(isig_d, isig, ncomp * sizeof(double), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaStreamSynchronize(stream));     // isig lives on this frame
    wc_components_kernel<<<(unsigned)(((size_t)N * 32 + 255) / 256), 256, 0, stream>>>(masked_d, mean_d, N, S, eigvec_d,
                                                                                         isig_d, ncomp, components_d);
    wc_transform_kernel<<<dim3((S + 127) / 128, nchunk), 128, 0, stream>>>(masked_d, mean_d, N, S, components_d, ncomp,
                                                                            bins_per_chunk, partial);
    wc_transform_reduce_kernel<<<(S * ncomp + 255) / 256, 256, 0, stream>>>(partial, S, ncomp, nchunk, T);
    wc_correct_kernel<<<(unsigned)(((size_t)N * S + 255) / 256), 256, 0, stream>>>(masked_d, mean_d, N, S, components_d,
                                                                                     T, ncomp, corrected_d);
    WC_CUDA(cudaGetLastError());
    ctx->counter[7] += 4;
    return WC_OK;
}
