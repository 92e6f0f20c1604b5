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

/* Debug aid: enable (1) / disable (0) per-CTA cycle counters in the distance kernel and copy the counters of the
 * most recent search to out_h (grid x 8 int64: total, wait-on-TMA, epilogue, prune, tiles, prunes, emitted by
 * thread 0, reserved).  Returns the number of CTAs copied, or a negative wc_status. */
int wc_debug_profile(wc_ctx* ctx, int enable, long long* out_h, int max_ctas);

/* Tuning knob for experiments: key "k5_lag" = chunks (0..4) by which half of the distance kernel's MMA warps
 * trail the other half.  Results never depend on it. */
int wc_set_option(wc_ctx* ctx, const char* key, double value);

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

#ifdef __cplusplus
}
#endif
#endif /* WISECONDOR_B200_H */
