/*
 * wisecondor_b200 - C ABI of the B200-native WISECONDOR hot path (libwisecondor_b200.so).
 *
 * The reference (VUmcCGP/wisecondor) has no FFI: its hot path is a set of Python functions in wisetools.py /
 * triarray.py called from wisecondor.py's tool functions.  Each entry point below replaces the body of one of
 * those functions; the citation after each declaration is the reference interface it stands in for
 * (file:line under /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - plain C types only; every pointer is either a DEVICE pointer (suffix _d) or a HOST pointer (suffix _h);
 *   - every function returns 0 on success or a negative wc_status; wc_last_error() gives the text;
 *   - work is enqueued on the cudaStream_t passed as `stream` (a void*; NULL = legacy default stream);
 *     functions documented as "synchronous" wait for their own work before returning;
 *   - nothing here falls back to the CPU: without a CUDA device every compute entry point fails with
 *     WC_ERR_CUDA.
 */
#ifndef WISECONDOR_B200_H
#define WISECONDOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wc_ctx wc_ctx;

typedef enum wc_status {
    WC_OK = 0,
    WC_ERR_ARG = -1,      /* bad argument */
    WC_ERR_CUDA = -2,     /* CUDA runtime / driver error, or no device */
    WC_ERR_NOMEM = -3,    /* device allocation failed */
    WC_ERR_INTERNAL = -4  /* invariant violated (please report) */
} wc_status;

/* One segmentation call, the tuple the reference's TriArr.segmentTri returns (triarray.py:78):
 * (value, (x, y)) with inclusive cleaned-bin coordinates inside the chromosome. */
typedef struct wc_call {
    int32_t sample;   /* index into the batch */
    int32_t chrom;    /* 0-based index into the chromosome list given to wc_segment_batch */
    int32_t x;        /* first cleaned bin of the run */
    int32_t y;        /* last cleaned bin of the run (inclusive) */
    double z;         /* Stouffer z of the run */
} wc_call;

/* ---- context ------------------------------------------------------------------------------------------ */
/* Create a context bound to CUDA device `device`.  Returns NULL on failure (see wc_last_error). */
wc_ctx* wc_create(int device);
void wc_destroy(wc_ctx* ctx);
const char* wc_last_error(void);
/* Library version string and number of SMs of the bound device (0 if none). */
const char* wc_version(void);
int wc_sm_count(const wc_ctx* ctx);
/* Device-time (ms, CUDA events on `stream`) of the named phases of the most recent call:
 * which: 0 = centre+norms (K4), 1 = distance+streaming top-k (K5), 2 = exact re-score/finalise (K6),
 *        3 = exhaustive fallback rows, 4 = z-score passes (K8), 5 = segmentation (K9), 6 = prep (K1-K3). */
double wc_last_phase_ms(const wc_ctx* ctx, int which);
/* Counters of the most recent wc_newref_topk call: which: 0 = kernel launches, 1 = rows sent to the exhaustive
 * fallback, 2 = candidate entries emitted by K5, 3 = tiles computed, 4 = CTAs launched for K5. */
long long wc_last_counter(const wc_ctx* ctx, int which);

/* ---- newref: reference-bin search --------------------------------------------------------------------- */
/* Replaces getReference + getRefForBins (wisetools.py:364-398, 298-325) for target rows
 * [row_begin, row_end) - the rows wisetools.getPart (wisetools.py:358-361) hands one part.
 *   corrected_d  DEVICE  N x S float64, bin-major C order (row = bin, S contiguous): the `correctedData` of
 *                        the prep npz (wisecondor.py:106)
 *   chrom_bins_h HOST    nchrom masked bins per chromosome (`maskedChromBins`, wisecondor.py:93); sum == N
 *   refsize              `selectRefAmount` / -refsize (wisecondor.py:379), 1..384
 *   idx_d        DEVICE  (row_end-row_begin) x refsize int32: positions in the other-chromosome concatenation
 *   dist_d       DEVICE  (row_end-row_begin) x refsize float64: squared distances, ascending
 * Rows are the first `refsize` candidates ordered by (distance, index); unfilled slots hold -1 / 1e10.
 * Distances are re-scored in the reference's operation order (sequential over samples, separately rounded
 * subtract/multiply/add) and are bit-identical to the reference's.  Synchronous. */
int wc_newref_topk(wc_ctx* ctx, const double* corrected_d, int N, int S, const int* chrom_bins_h, int nchrom,
                   int row_begin, int row_end, int refsize, int32_t* idx_d, double* dist_d, void* stream);

/* Same call with HOST buffers (copies in and out inside the call): the form toolNewrefPart
 * (wisecondor.py:111-132) would bind. */
int wc_newref_topk_host(wc_ctx* ctx, const double* corrected_h, int N, int S, const int* chrom_bins_h,
                        int nchrom, int row_begin, int row_end, int refsize, int32_t* idx_h, double* dist_h);

/* ---- newref: normalise, mask, PCA --------------------------------------------------------------------- */
/* Replaces toNumpyArray's arithmetic (wisetools.py:255-261): counts_d is S x Nraw int32 (sample-major, the
 * stacked chr1..22 arrays).  Writes sample totals, the nonzero-bin mask (1 byte per raw bin) and returns the
 * masked bin count through N_out.  Synchronous. */
int wc_normalize_mask(wc_ctx* ctx, const int32_t* counts_d, int S, int Nraw, double* totals_d,
                      uint8_t* mask_d, int* N_out, void* stream);
/* Builds the masked, normalised matrix in both layouts: Xs_d (S x N, sample-major) and, when Xb_d != NULL,
 * Xb_d (N x S, bin-major) - `maskedData` (wisetools.py:261). */
int wc_gather_masked(wc_ctx* ctx, const int32_t* counts_d, int S, int Nraw, const double* totals_d,
                     const uint8_t* mask_d, int N, double* Xs_d, double* Xb_d, void* stream);
/* Replaces trainPCA (wisetools.py:89-101): per-bin mean over samples, Gram matrix of the centred data
 * (S x S, fp64 DMMA tiles), eigen-decomposition of the S x S Gram on the host (LAPACK-free Jacobi is not used:
 * the caller supplies the top-ncomp eigenvectors through wc_pca_finish), projection and residual.
 * Step 1: mean_d[N] and gram_d[S x S] from Xs_d (S x N). */
int wc_pca_gram(wc_ctx* ctx, const double* Xs_d, int S, int N, double* mean_d, double* gram_d, void* stream);
/* Step 2: given the top-ncomp unit eigenvectors U (ncomp x S, row-major) and eigenvalues of gram, compute
 * components_d (ncomp x N, sign-normalised like sklearn's svd_flip), and corrected_d (N x S bin-major) =
 * X / (mean + (Xc V^T) V).  */
int wc_pca_finish(wc_ctx* ctx, const double* Xs_d, int S, int N, const double* mean_d, const double* U_h,
                  const double* eigval_h, int ncomp, double* components_d, double* corrected_d, void* stream);

/* ---- test: sample prep, z-scores, segmentation --------------------------------------------------------- */
/* Replaces toNumpyRefFormat + applyPCA (wisetools.py:267-278, 104-113) for a batch: counts_d is B x Nraw int32
 * already padded/truncated to the reference's chromosome sizes; out T_d is N x ldB float64 ([bin][sample]). */
int wc_test_prep(wc_ctx* ctx, const int32_t* counts_d, int B, int Nraw, const uint8_t* mask_d, int N,
                 const double* mean_d, const double* components_d, int ncomp, double* T_d, int ldB, void* stream);
/* Replaces repeatTest / trySample (wisetools.py:407-448) for a batch of B samples laid out [bin][sample].
 *   ref_idx_d  global masked-bin index of every kept reference bin (those with distance < cutoff), CSR
 *   ref_off_d  N+1 offsets into ref_idx_d
 *   Z_d, R_d   N x ldB outputs; refsizes_d N x ldB int32; asdef_d B (average reference sigma, wisetools.py:435)
 *   work_d     2 * N * ldB doubles of scratch (the evolving testCopy, double-buffered) */
int wc_test_batch(wc_ctx* ctx, const double* T_d, int N, int B, int ldB, const int32_t* ref_idx_d,
                  const int32_t* ref_off_d, double z_threshold, int repeats, double* Z_d, double* R_d,
                  int32_t* refsizes_d, double* asdef_d, double* work_d, void* stream);
/* Replaces fillTri + TriArr.segmentTri (wisetools.py:466-472, triarray.py:59-84) for a batch: Zc_d holds, per
 * sample, the cleaned z-scores of the requested chromosomes back to back; off_d is B x (nchrom+1) offsets into
 * Zc_d.  Writes cwz_d (B x nchrom chromosome-wide z), up to max_calls wc_call records (device) and the number
 * found (device int; > max_calls means truncated). */
int wc_segment_batch(wc_ctx* ctx, const double* Zc_d, const int64_t* off_d, int B, int nchrom,
                     double z_threshold, int min_search, double* cwz_d, wc_call* calls_d, int* ncalls_d,
                     int max_calls, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WISECONDOR_B200_H */
