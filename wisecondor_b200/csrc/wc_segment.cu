// wisecondor_b200 - Stouffer segmentation of the z-scores on B200 (sm_100a).
//
// Replaces fillTri + TriArr.segmentTri (/root/reference/wisetools.py:466-472, /root/reference/triarray.py:59-84) and
// the chromosome loop of toolTest (/root/reference/wisecondor.py:233-238) for a batch of samples: per (sample,
// chromosome) the z-scores of the kept bins are compacted, every contiguous run [x..y] is scored
// sum(z[x..y]) / sqrt(y-x+1), the most significant run above the threshold is called and the search recurses to its
// left and right.  The reference materialises the n(n+1)/2 run values (99 MB for chr1 at 50 kb) with an O(n) numpy sum
// each; here one CTA keeps the chromosome's prefix sums in shared memory and sweeps the triangle in registers.
//
// Exactness.  The sweep scores runs from prefix-sum differences times a reciprocal-sqrt table - not numpy's summation
// order - so it is used only to *locate*: every run whose score lies within a rigorous error window `delta` of the sweep's
// maximum (minimum) is re-scored in numpy's own pairwise order (wc_numpy_order.cuh) and divided by sqrt(len) exactly
// as the reference does; champion and tie-breaking (first occurrence in the reference's row-major triangle order:
// smallest x, then smallest y, triarray.py:62-66 argmax/argmin) are decided on those exact values.  Calls and their z
// are therefore identical to the reference's, not merely close.
//
// K9 wc_segment_kernel   one CTA per (sample, chromosome): compaction -> prefix sums -> [sweep -> exact re-score of the
//                        windowed rows -> call -> push left/right ranges]*                       (FP64 ALU / shared memory)
#include "wc_common.cuh"
#include "wc_numpy_order.cuh"

namespace {

constexpr int SEG_THREADS = 256;
constexpr int SEG_WARPS = SEG_THREADS / 32;
constexpr int SEG_R = 5;                      // rows per lane in the sweep (odd: conflict-free table reads)
constexpr int SEG_TILE = 32 * SEG_R;          // rows per warp tile
constexpr int SEG_PAD = SEG_TILE;             // NaN entries in front of the reciprocal-sqrt table
constexpr int SEG_STACK = 128;                // pending ranges per chromosome
constexpr int SEG_ROWCAP = 1024;              // rows re-scored exactly per search before falling back to all rows
constexpr double SEG_EPS = 1.1102230246251565e-16;

struct SegArgs {
    const double* z;         // [B][N]
    const int* refsz;        // [B][N]
    int N, B;
    const int* sel_start;    // [nsel] first masked bin of each selected chromosome (sorted by descending length)
    const int* sel_len;      // [nsel]
    const int* sel_slot;     // [nsel] position of the chromosome in the caller's list
    int nsel;
    int minrefbins;
    double thr;
    int min_search;
    const double* isq;       // [SEG_PAD + maxlen]: isq[SEG_PAD + L - 1] = 1/sqrt(L); NaN in the pad
    double* cwz;             // [B][nsel]
    int* cleaned;            // [B][nsel]
    wc_call* calls;          // [B][max_calls]
    int* ncalls;             // [B]
    int max_calls;
    int* status;             // [0] |= 1 call overflow, 2 range-stack overflow, 4 non-finite z in a kept bin
    int zcap;                // capacity (doubles) of each shared array
};

struct Best {                // lexicographic champion: value, then first occurrence (x, then y)
    double v;
    int x, y;
};
__device__ __forceinline__ bool better_max(const Best& a, const Best& b) {   // a beats b
    return a.v > b.v || (a.v == b.v && (a.x < b.x || (a.x == b.x && a.y < b.y)));
}
__device__ __forceinline__ bool better_min(const Best& a, const Best& b) {
    return a.v < b.v || (a.v == b.v && (a.x < b.x || (a.x == b.x && a.y < b.y)));
}
__device__ __forceinline__ Best shfl_best(const Best& b, int o) {
    Best r;
    r.v = __shfl_xor_sync(0xffffffffu, b.v, o);
    r.x = __shfl_xor_sync(0xffffffffu, b.x, o);
    r.y = __shfl_xor_sync(0xffffffffu, b.y, o);
    return r;
}

__global__ void wc_isq_kernel(double* isq, int maxlen) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < SEG_PAD) isq[i] = __longlong_as_double(0x7ff8000000000000ll);
    if (i < maxlen) isq[SEG_PAD + i] = __ddiv_rn(1.0, sqrt((double)(i + 1)));
}

__global__ void __launch_bounds__(SEG_THREADS) wc_segment_kernel(const SegArgs a) {
    extern __shared__ __align__(16) unsigned char seg_raw[];
    double* zc = reinterpret_cast<double*>(seg_raw);              // [zcap] kept z-scores of this chromosome
    double* P = zc + a.zcap;                                      // [zcap + 8] prefix sums, P[i] = sum zc[0..i)
    __shared__ double s_red[2][SEG_WARPS];
    __shared__ Best s_best[2][SEG_WARPS];
    __shared__ int s_scan[SEG_WARPS];
    __shared__ int s_lo[SEG_STACK], s_hi[SEG_STACK];
    __shared__ int s_rows[SEG_ROWCAP];
    __shared__ int s_nrows, s_sp, s_n, s_bad;
    __shared__ double s_A;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sel = blockIdx.x / a.B, b = blockIdx.x % a.B;     // long chromosomes first
    const int cstart = a.sel_start[sel], clen = a.sel_len[sel], slot = a.sel_slot[sel];
    const double* zrow = a.z + (size_t)b * a.N + cstart;
    const int* nrow = a.refsz + (size_t)b * a.N + cstart;

    // ---- 1. compaction of the kept bins (wisecondor.py:215-218: refSizes >= minrefbins) -------------------------
    if (tid == 0) { s_n = 0; s_bad = 0; }
    __syncthreads();
    for (int base = 0; base < clen; base += SEG_THREADS) {
        const int i = base + tid;
        const bool keep = i < clen && nrow[i] >= a.minrefbins;
        const double v = keep ? zrow[i] : 0.0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int off = s_n;
        for (int w = 0; w < warp; ++w) off += s_scan[w];
        if (keep) {
            zc[off + __popc(bal & ((1u << lane) - 1u))] = v;
            if (!isfinite(v)) s_bad = 1;
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < SEG_WARPS; ++w) t += s_scan[w];
            s_n += t;
        }
        __syncthreads();
    }
    const int n = s_n;
    if (tid == 0) a.cleaned[(size_t)b * a.nsel + slot] = n;
    if (n == 0) {                                   // the reference raises on an empty chromosome (triarray.py:29)
        if (tid == 0) a.cwz[(size_t)b * a.nsel + slot] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    if (s_bad) {                                    // +-inf / NaN z (a kept bin whose reference sigma is 0)
        if (tid == 0) {
            atomicOr(a.status, 4);
            a.cwz[(size_t)b * a.nsel + slot] = __longlong_as_double(0x7ff8000000000000ll);
        }
        return;
    }

    // ---- 2. prefix sums and A = sum |z| (the scale of the error window) ---------------------------------------------
    {
        const int per = (n + SEG_THREADS - 1) / SEG_THREADS;
        const int i0 = min(n, tid * per), i1 = min(n, i0 + per);
        double loc = 0.0, la = 0.0;
        for (int i = i0; i < i1; ++i) { loc += zc[i]; la += fabs(zc[i]); }
        double inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) la += __shfl_xor_sync(0xffffffffu, la, o);
        if (lane == 31) s_red[0][warp] = inc;
        if (lane == 0) s_red[1][warp] = la;
        __syncthreads();
        double base = 0.0;
        for (int w = 0; w < warp; ++w) base += s_red[0][w];
        double run = base + (inc - loc);
        for (int i = i0; i < i1; ++i) { P[i] = run; run += zc[i]; }
        if (i1 == n && i0 < n) P[n] = run;
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < SEG_WARPS; ++w) t += s_red[1][w];
            s_A = t;
            s_sp = 1;
            s_lo[0] = 0;
            s_hi[0] = n;
        }
        __syncthreads();
    }
    const double A = s_A;
    // |sweep value - numpy value| <= (2 (per + 16) + 40) eps A + 3 eps |v|: prefix sums (sequential chunk of `per`, warp
    // scan, warp bases) on both ends of the run, numpy's own pairwise rounding, the reciprocal multiply.  delta = 2 x that.
    const double delta_A = (4.0 * ((n + SEG_THREADS - 1) / SEG_THREADS) + 160.0) * SEG_EPS * A;

    // chromosome-wide value: entry (0, n-1) of the triangle (wisecondor.py:237), numpy order, by one thread
    if (tid == SEG_THREADS - 1) {
        const double tot = np_sum_thread([&](int i) { return zc[i]; }, n);
        a.cwz[(size_t)b * a.nsel + slot] = __ddiv_rn(tot, sqrt((double)n));
    }

    const double* isq0 = a.isq + SEG_PAD;           // isq0[L - 1] = 1/sqrt(L); isq0[-1 .. -SEG_PAD] = NaN
    const double NaN = __longlong_as_double(0x7ff8000000000000ll);

    // ---- 3. iterative most-significant-run search (triarray.py:59-84) --------------------------------------------------
    while (true) {
        __syncthreads();
        const int sp = s_sp;
        if (sp == 0) break;
        const int lo = s_lo[sp - 1], hi = s_hi[sp - 1];
        __syncthreads();
        if (tid == 0) s_sp = sp - 1;
        const int m = hi - lo;

        // -- sweep: per-thread max / min of (P[y+1] - P[x]) * isq[y - x] over its rows --
        double vmax = -INFINITY, vmin = INFINITY;
        const int ntiles = (m + SEG_TILE - 1) / SEG_TILE;
        for (int round = 0; round * SEG_WARPS < ntiles; ++round) {
            const int t = round * SEG_WARPS + ((round & 1) ? SEG_WARPS - 1 - warp : warp);   // boustrophedon: balance
            if (t >= ntiles) continue;
            const int xb = lo + t * SEG_TILE;
            const int x0 = xb + lane * SEG_R;
            double px[SEG_R];
#pragma unroll
            for (int r = 0; r < SEG_R; ++r) px[r] = (x0 + r < hi) ? P[x0 + r] : NaN;
            // window of reciprocal square roots: slot (t mod 5) holds isq0[q0 + t], q = y - x0
            double ww[SEG_R];
            int q0 = xb - x0;                         // <= 0
#pragma unroll
            for (int r = 1; r < SEG_R; ++r) ww[SEG_R - r] = isq0[q0 - r];
            int y = xb;
            for (; y + SEG_R <= hi; y += SEG_R, q0 += SEG_R) {
#pragma unroll
                for (int j = 0; j < SEG_R; ++j) {
                    ww[j] = isq0[q0 + j];
                    const double py = P[y + j + 1];
#pragma unroll
                    for (int r = 0; r < SEG_R; ++r) {
                        const double v = __dmul_rn(__dsub_rn(py, px[r]), ww[(j - r + SEG_R) % SEG_R]);
                        vmax = fmax(vmax, v);
                        vmin = fmin(vmin, v);
                    }
                }
            }
            for (; y < hi; ++y) {
                const double py = P[y + 1];
#pragma unroll
                for (int r = 0; r < SEG_R; ++r) {
                    const double v = __dmul_rn(__dsub_rn(py, px[r]), isq0[y - x0 - r]);
                    vmax = fmax(vmax, v);
                    vmin = fmin(vmin, v);
                }
            }
        }
        // -- block-wide extremes --
        double bmax = vmax, bmin = vmin;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
            bmin = fmin(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
        }
        if (lane == 0) { s_red[0][warp] = bmax; s_red[1][warp] = bmin; }
        if (tid == 0) s_nrows = 0;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < SEG_WARPS; ++w) { bmax = fmax(bmax, s_red[0][w]); bmin = fmin(bmin, s_red[1][w]); }
        const double big = fmax(fabs(bmax), fabs(bmin));
        const double delta = delta_A + 16.0 * SEG_EPS * big;
        if (big + delta < a.thr) continue;            // abs(champVal) < threshold for certain (triarray.py:72-73)

        // -- rows that can hold the exact champion: those of the threads whose own extreme lies in the window --
        const bool cand = vmax >= bmax - delta || vmin <= bmin + delta;
        if (cand) {
            for (int round = 0; round * SEG_WARPS < ntiles; ++round) {
                const int t = round * SEG_WARPS + ((round & 1) ? SEG_WARPS - 1 - warp : warp);
                if (t >= ntiles) continue;
                const int x0 = lo + t * SEG_TILE + lane * SEG_R;
                for (int r = 0; r < SEG_R; ++r) {
                    if (x0 + r < hi) {
                        const int pos = atomicAdd(&s_nrows, 1);
                        if (pos < SEG_ROWCAP) s_rows[pos] = x0 + r;
                    }
                }
            }
        }
        __syncthreads();
        const int nrows_raw = s_nrows;
        const bool all_rows = nrows_raw > SEG_ROWCAP;
        const int nrows = all_rows ? m : nrows_raw;

        // -- exact re-score (numpy order) of every windowed run of those rows --
        Best cmax = {-INFINITY, 0x7fffffff, 0x7fffffff}, cmin = {INFINITY, 0x7fffffff, 0x7fffffff};
        for (int ri = 0; ri < nrows; ++ri) {
            const int x = all_rows ? lo + ri : s_rows[ri];
            const double pxv = P[x];
            for (int y = x + tid; y < hi; y += SEG_THREADS) {
                const double v = __dmul_rn(__dsub_rn(P[y + 1], pxv), isq0[y - x]);
                const bool wmax = v >= bmax - delta, wmin = v <= bmin + delta;
                if (wmax || wmin) {
                    const int len = y - x + 1;
                    const double sum = np_sum_thread([&](int i) { return zc[x + i]; }, len);   // np_sum(region[x:y+1])
                    Best e = {__ddiv_rn(sum, sqrt((double)len)), x, y};                          // / np_sqrt(y-x+1)
                    if (wmax && better_max(e, cmax)) cmax = e;
                    if (wmin && better_min(e, cmin)) cmin = e;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const Best om = shfl_best(cmax, o), on = shfl_best(cmin, o);
            if (better_max(om, cmax)) cmax = om;
            if (better_min(on, cmin)) cmin = on;
        }
        if (lane == 0) { s_best[0][warp] = cmax; s_best[1][warp] = cmin; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 0; w < SEG_WARPS; ++w) {
                if (better_max(s_best[0][w], cmax)) cmax = s_best[0][w];
                if (better_min(s_best[1][w], cmin)) cmin = s_best[1][w];
            }
            Best champ = cmax;                                          // triarray.py:62-70
            if (fabs(cmin.v) > champ.v) champ = cmin;
            if (!(fabs(champ.v) < a.thr)) {                             // triarray.py:72-73
                const int slot_i = atomicAdd(&a.ncalls[b], 1);
                if (slot_i < a.max_calls) {
                    wc_call c;
                    c.sample = b; c.chrom = slot; c.x = champ.x; c.y = champ.y; c.z = champ.v;
                    a.calls[(size_t)b * a.max_calls + slot_i] = c;
                } else {
                    atomicOr(a.status, 1);
                }
                int spn = s_sp;
                const int xr = champ.x - lo, yr = champ.y - lo;        // coordinates inside the sub-triangle
                if (yr + 1 < m - a.min_search) {                        // triarray.py:79
                    if (spn < SEG_STACK) { s_lo[spn] = champ.y + 1; s_hi[spn] = hi; ++spn; } else atomicOr(a.status, 2);
                }
                if (xr > a.min_search) {                                // triarray.py:76
                    if (spn < SEG_STACK) { s_lo[spn] = lo; s_hi[spn] = champ.x; ++spn; } else atomicOr(a.status, 2);
                }
                s_sp = spn;
            }
        }
    }
}

}  // namespace

extern "C" int wc_segment_batch(wc_ctx* ctx, const double* z_d, const int32_t* refsizes_d, int N, int B,
                                const int* chrom_bins_h, int nchrom, const int* chromosomes_h, int nsel, int minrefbins,
                                double z_threshold, int min_search, double* cwz_d, int32_t* cleaned_bins_d,
                                wc_call* calls_d, int32_t* ncalls_d, int max_calls, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && z_d != nullptr && refsizes_d != nullptr && chrom_bins_h != nullptr);
    WC_CHECK_ARG(chromosomes_h != nullptr && cwz_d != nullptr && cleaned_bins_d != nullptr);
    WC_CHECK_ARG(calls_d != nullptr && ncalls_d != nullptr && max_calls > 0);
    WC_CHECK_ARG(N > 0 && B > 0 && nchrom > 0 && nsel > 0 && min_search >= 0);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    std::vector<int> start(nchrom);
    long long tot = 0;
    for (int c = 0; c < nchrom; ++c) { WC_CHECK_ARG(chrom_bins_h[c] >= 0); start[c] = (int)tot; tot += chrom_bins_h[c]; }
    WC_CHECK_ARG(tot == N);
    // selected chromosomes, longest first (they are the longest-running CTAs)
    std::vector<int> order(nsel);
    int maxlen = 1;
    for (int i = 0; i < nsel; ++i) {
        WC_CHECK_ARG(chromosomes_h[i] >= 0 && chromosomes_h[i] < nchrom);
        order[i] = i;
        maxlen = std::max(maxlen, chrom_bins_h[chromosomes_h[i]]);
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return chrom_bins_h[chromosomes_h[x]] > chrom_bins_h[chromosomes_h[y]];
    });
    std::vector<int> meta(3 * nsel);
    for (int i = 0; i < nsel; ++i) {
        const int c = chromosomes_h[order[i]];
        meta[i] = start[c];
        meta[nsel + i] = chrom_bins_h[c];
        meta[2 * nsel + i] = order[i];
    }
    double* isq; int* meta_d; int* status_d;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_S_ISQ, (size_t)(SEG_PAD + maxlen + 8) * sizeof(double), (void**)&isq))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_S_META, (size_t)3 * nsel * sizeof(int), (void**)&meta_d))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_S_STATUS, 4 * sizeof(int), (void**)&status_d))) return rc;
    const int zcap = (maxlen + 8 + 1) & ~1;
    const size_t smem = (size_t)(2 * zcap + 8) * sizeof(double);
    if (smem > 220 * 1024) {
        wc_set_error("segmentation: a chromosome of %d bins needs %zu bytes of shared memory (limit 220 KiB)", maxlen, smem);
        return WC_ERR_ARG;
    }
    WC_CUDA(cudaEventRecord(ctx->ev[10], stream));
    WC_CUDA(cudaMemcpyAsync(meta_d, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemsetAsync(status_d, 0, 4 * sizeof(int), stream));
    WC_CUDA(cudaMemsetAsync(ncalls_d, 0, (size_t)B * sizeof(int), stream));
    wc_isq_kernel<<<(SEG_PAD + maxlen + 255) / 256, 256, 0, stream>>>(isq, maxlen);
    SegArgs a;
    a.z = z_d; a.refsz = refsizes_d; a.N = N; a.B = B; a.sel_start = meta_d; a.sel_len = meta_d + nsel;
    a.sel_slot = meta_d + 2 * nsel; a.nsel = nsel; a.minrefbins = minrefbins; a.thr = z_threshold;
    a.min_search = min_search; a.isq = isq; a.cwz = cwz_d; a.cleaned = cleaned_bins_d; a.calls = calls_d;
    a.ncalls = ncalls_d; a.max_calls = max_calls; a.status = status_d; a.zcap = zcap;
    WC_CUDA(cudaFuncSetAttribute(wc_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wc_segment_kernel<<<(unsigned)((size_t)nsel * B), SEG_THREADS, smem, stream>>>(a);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[11], stream));
    int status = 0;
    WC_CUDA(cudaMemcpyAsync(&status, status_d, sizeof(int), cudaMemcpyDeviceToHost, stream));
    WC_CUDA(cudaStreamSynchronize(stream));       // meta/status host buffers; the call is documented as synchronous
    ctx->timed_mask |= 1u << 5;
    ctx->counter[6] = 2;
    if (status & 4) {
        wc_set_error("segmentation: a kept bin has a non-finite z-score (reference sigma 0); not supported yet");
        return WC_ERR_ARG;
    }
    if (status & 2) { wc_set_error("segmentation: more than %d pending ranges on one chromosome", SEG_STACK); return WC_ERR_INTERNAL; }
    if (status & 1) { wc_set_error("segmentation: more than max_calls=%d calls for one sample", max_calls); return WC_ERR_ARG; }
    return WC_OK;
}
