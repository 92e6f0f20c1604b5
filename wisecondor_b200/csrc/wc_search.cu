// wisecondor_b200 - newref reference-bin search on B200 (sm_100a).
//
// Replaces getReference / getRefForBins (/root/reference/wisetools.py:364-398, 298-325): for every target bin,
// the squared Euclidean distance over samples to every bin on the other chromosomes, keeping the `refsize`
// smallest ordered by (distance, index).
//
// Kernels
//   K4 wc_center_norms_kernel   X' = X - 1 (padded, TMA-friendly) and n_i = sum_s X'[i][s]^2          (HBM bound)
//   K5 wc_dist_topk_kernel      fp64 tensor-core contraction d~ = n_i + n_j - 2 X'_i . X'_j on 128x128 tiles,
//                               operands staged by TMA (SWIZZLE_128B) through a 4-stage mbarrier ring, fused
//                               with a streaming per-row threshold filter: only entries that can still be among
//                               the row's k smallest leave the SM.  The distance matrix never exists. (FP64 bound)
//   K6 wc_finalize_kernel       per row: shortlist = entries within an error margin of the k-th smallest d~,
//                               exact re-score in the reference's operation order, sort by (d, index), remap
//                               to other-chromosome coordinates                                       (L2 bound)
//   K6b wc_exhaustive_kernel    exact brute force for rows the streaming path could not bound (massive ties)
#include "wc_common.cuh"

namespace {

constexpr int BM = 128;             // target rows per CTA tile
constexpr int BN = 128;             // candidate columns per CTA tile
constexpr int BK = 16;              // samples per pipeline stage (16 doubles = one 128-byte swizzle row)
constexpr int STAGES = 4;
constexpr int CONSUMER_WARPS = 8;   // 4 (rows) x 2 (cols) warps, each a 32 x 64 sub-tile
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
constexpr int PRODUCER_WARPS = 4;   // one warp group; only warp 0 lane 0 issues TMA, the group donates registers
constexpr int TOPK_THREADS = CONSUMER_THREADS + PRODUCER_WARPS * 32;
constexpr int TILE_BYTES = BM * BK * 8;               // 16 KiB per operand per stage
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int HIST_BINS = 256;
constexpr int FIN_THREADS = 128;
constexpr int FIN_MAX = 2048;       // candidates a finalize CTA can hold
constexpr int FIN_SHORT = 256;      // exact re-score capacity (2 passes of 128)
constexpr int EXH_THREADS = 256;

struct TopkSmem {
    alignas(1024) unsigned char tiles[STAGES][STAGE_BYTES];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    double tau[BM];      // emission threshold per row (includes the error margin); NaN = row inactive
    double nrm[BM];      // n_i
    int cs[BM];          // excluded column range [cs, ce) = the row's own chromosome
    int ce[BM];
    int cnt[BM];         // entries in the row's candidate buffer
    int flag[BM];        // 1 = buffer could not be bounded -> exhaustive fallback
    int hist[CONSUMER_WARPS][HIST_BINS];
};

struct TopkArgs {
    const double* norms;     // [Npad], NaN beyond N
    const int* row_cs;       // [N]
    const int* row_ce;       // [N]
    int N;
    int row_begin, row_end;
    int nkc;                 // k chunks of BK
    int nsteps_last;         // k4 steps in the last chunk (1..4)
    const int* rb_tile_prefix;   // [nrb+1] valid tiles before row block rb
    const int* rb_skip_lo;       // [nrb] first skipped column tile (own chromosome interior)
    const int* rb_skip_n;        // [nrb] number of skipped column tiles
    int nrb;
    int total_tiles;
    const int* cta_seg_base;     // [grid] first segment id of each CTA
    double* cand_d;              // [nseg][BM][cap]
    int* cand_j;
    int* seg_cnt;                // [nseg][BM]
    int* seg_flag;               // [nseg][BM]
    int cap;
    int k;
    double mcoef;                // margin(v) = mcoef * (n_i + |v|)
    double tau_init;
};

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int bucket_of(double d, double mn, float scale) {
    int b = (int)((float)(d - mn) * scale);
    return b < 0 ? 0 : (b > HIST_BINS - 1 ? HIST_BINS - 1 : b);
}

// Warp-collective prune of one row's candidate buffer: keep every entry <= v* + margin, where v* is the largest
// entry of the histogram bucket that holds the k-th smallest.  The kept set contains the k smallest, so v* is a
// valid upper bound of the row's final k-th smallest distance.
__device__ void prune_row(double* cd, int* cj, int n, int k, double nrm, double mcoef, int* hist, int lane,
                          double* tau_out, int* n_out) {
    double mn = INFINITY, mx = -INFINITY;
    for (int i = lane; i < n; i += 32) {
        double d = ld_cg_f64(cd + i);
        mn = fmin(mn, d);
        mx = fmax(mx, d);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    float scale = (mx > mn) ? (float)(HIST_BINS - 1) / (float)(mx - mn) : 0.0f;
    for (int b = lane; b < HIST_BINS; b += 32) hist[b] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) atomicAdd(&hist[bucket_of(ld_cg_f64(cd + i), mn, scale)], 1);
    __syncwarp();
    int c[HIST_BINS / 32], sum = 0;
#pragma unroll
    for (int t = 0; t < HIST_BINS / 32; ++t) {
        c[t] = hist[lane * (HIST_BINS / 32) + t];
        sum += c[t];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int run = incl - sum, bl = -1;
#pragma unroll
    for (int t = 0; t < HIST_BINS / 32; ++t) {
        run += c[t];
        if (bl < 0 && run >= k) bl = lane * (HIST_BINS / 32) + t;
    }
    unsigned m = __ballot_sync(0xffffffffu, bl >= 0);
    int bstar = HIST_BINS - 1;
    if (m) bstar = __shfl_sync(0xffffffffu, bl, __ffs(m) - 1);
    double vstar = -INFINITY;
    for (int i = lane; i < n; i += 32) {
        double d = ld_cg_f64(cd + i);
        if (bucket_of(d, mn, scale) <= bstar) vstar = fmax(vstar, d);
    }
    vstar = warp_max(vstar);
    double tau = vstar + mcoef * (nrm + fabs(vstar));
    int w = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        double d = 0.0;
        int j = 0;
        bool keep = false;
        if (i < n) {
            d = ld_cg_f64(cd + i);
            j = ld_cg_s32(cj + i);
            keep = d <= tau;
        }
        unsigned km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            int pos = w + __popc(km & ((1u << lane) - 1u));
            cd[pos] = d;
            cj[pos] = j;
        }
        w += __popc(km);
        __syncwarp();
    }
    *tau_out = tau;
    *n_out = w;
}

// ---------------------------------------------------------------------------------------------------------
// K5: distance tiles on the FP64 tensor cores + streaming top-k filter
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TOPK_THREADS, 1)
wc_dist_topk_kernel(const __grid_constant__ CUtensorMap tmap, const TopkArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TopkSmem& sm = *reinterpret_cast<TopkSmem*>(smem_raw);
    if (smem_u32(smem_raw) & 1023u) __trap();   // SWIZZLE_128B tiles need a 1024-byte aligned base
    const int tid = threadIdx.x;
    const int warp_all = tid >> 5, lane = tid & 31;
    const int warp = warp_all - PRODUCER_WARPS;     // consumer warp index (negative in the producer group)

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    // this CTA's contiguous range of the (row block, column tile) work list
    const long long T = a.total_tiles;
    const int lin0 = (int)(T * blockIdx.x / gridDim.x);
    const int lin1 = (int)(T * (blockIdx.x + 1) / gridDim.x);
    if (lin0 >= lin1) return;
    int rb = 0;
    {
        int lo = 0, hi = a.nrb;   // largest rb with prefix[rb] <= lin0
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (a.rb_tile_prefix[mid] <= lin0) lo = mid; else hi = mid;
        }
        rb = lo;
    }

    if (warp_all < PRODUCER_WARPS) {
        // ===== TMA producer: one elected lane streams the A (target rows) and B (candidate rows) k-slices =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp_all == 0 && lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int rbp = rb;
            for (int lin = lin0; lin < lin1; ++lin) {
                while (lin >= a.rb_tile_prefix[rbp + 1]) ++rbp;
                int q = lin - a.rb_tile_prefix[rbp];
                int t = q < a.rb_skip_lo[rbp] ? q : q + a.rb_skip_n[rbp];
                int row0 = a.row_begin + rbp * BM;
                int col0 = t * BN;
                for (int kc = 0; kc < a.nkc; ++kc) {
                    mbar_wait(&sm.empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&sm.full[stage], STAGE_BYTES);
                    tma_load_2d(sm.tiles[stage], &tmap, kc * BK, row0, &sm.full[stage]);
                    tma_load_2d(sm.tiles[stage] + TILE_BYTES, &tmap, kc * BK, col0, &sm.full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ===== consumers: 8 warps, warp tile 32 rows x 64 cols = 4 x 8 DMMA tiles =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int ctid = tid - PRODUCER_WARPS * 32;     // 0..255
    const int g = lane >> 2, q4 = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const uint32_t a_off = (uint32_t)(wm * 32 + g) * 128u;
    const uint32_t b_off = (uint32_t)TILE_BYTES + (uint32_t)(wn * 64 + g) * 128u;
    uint32_t sw[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) sw[s] = (uint32_t)((((2 * s + (q4 >> 1)) ^ g) << 4) | ((q4 & 1) << 3));

    const uint32_t tiles_u32 = smem_u32(&sm.tiles[0][0]);
    int stage = 0;
    uint32_t phase = 0;
    int seg = a.cta_seg_base[blockIdx.x];
    int cur_rb = -1;
    const size_t seg_stride = (size_t)BM * a.cap;

    for (int lin = lin0; lin < lin1; ++lin) {
        while (lin >= a.rb_tile_prefix[rb + 1]) ++rb;
        if (rb != cur_rb) {
            // segment boundary: flush the previous row block's state, load the new one's
            named_bar_sync(1, CONSUMER_THREADS);
            if (ctid < BM) {
                if (cur_rb >= 0) {
                    a.seg_cnt[(size_t)seg * BM + ctid] = sm.cnt[ctid];
                    a.seg_flag[(size_t)seg * BM + ctid] = sm.flag[ctid];
                }
                int row = a.row_begin + rb * BM + ctid;
                bool valid = row < a.row_end;
                sm.nrm[ctid] = valid ? a.norms[row] : 0.0;
                sm.cs[ctid] = valid ? a.row_cs[row] : 0;
                sm.ce[ctid] = valid ? a.row_ce[row] : 0;
                sm.tau[ctid] = valid ? a.tau_init : __longlong_as_double(0x7ff8000000000000LL);
                sm.cnt[ctid] = 0;
                sm.flag[ctid] = 0;
            }
            if (cur_rb >= 0) ++seg;
            cur_rb = rb;
            named_bar_sync(1, CONSUMER_THREADS);
        }
        const int q = lin - a.rb_tile_prefix[rb];
        const int t = q < a.rb_skip_lo[rb] ? q : q + a.rb_skip_n[rb];
        const int col0 = t * BN;

        double acc[4][8][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        for (int kc = 0; kc < a.nkc; ++kc) {
            mbar_wait(&sm.full[stage], phase);
            const uint32_t base = tiles_u32 + (uint32_t)stage * STAGE_BYTES;
            const int nsteps = (kc == a.nkc - 1) ? a.nsteps_last : 4;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (s < nsteps) {
                    double fa[4], fb[8];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
                        fa[mt] = lds_f64(base + a_off + mt * 1024 + sw[s]);
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
                        fb[nt] = lds_f64(base + b_off + nt * 1024 + sw[s]);
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) dmma_8x8x4(acc[mt][nt][0], acc[mt][nt][1], fa[mt], fb[nt]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }

        // ---- epilogue: d~ = (n_i + n_j) - 2 dot, filter against the row threshold, emit survivors ----
        double* cd = a.cand_d + (size_t)seg * seg_stride;
        int* cj = a.cand_j + (size_t)seg * seg_stride;
        double ncol[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            int j = col0 + wn * 64 + nt * 8 + q4 * 2;
            ncol[nt][0] = __ldg(a.norms + j);
            ncol[nt][1] = __ldg(a.norms + j + 1);
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int rl = wm * 32 + mt * 8 + g;
            const double tau = sm.tau[rl];
            const double nr = sm.nrm[rl];
            const int cs = sm.cs[rl];
            const unsigned clen = (unsigned)(sm.ce[rl] - cs);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int j = col0 + wn * 64 + nt * 8 + q4 * 2 + e;
                    const double d = fma(-2.0, acc[mt][nt][e], nr + ncol[nt][e]);
                    if (d <= tau && (unsigned)(j - cs) >= clen) {
                        int slot = atomicAdd(&sm.cnt[rl], 1);
                        if (slot < a.cap) {
                            cd[(size_t)rl * a.cap + slot] = d;
                            cj[(size_t)rl * a.cap + slot] = j;
                        } else {
                            sm.flag[rl] = 1;
                        }
                    }
                }
            }
        }
        named_bar_sync(1, CONSUMER_THREADS);
        // ---- prune rows whose buffer could overflow during the next tile ----
        for (int rl = warp * (BM / CONSUMER_WARPS); rl < (warp + 1) * (BM / CONSUMER_WARPS); ++rl) {
            int n = sm.cnt[rl];
            if (n > a.cap) n = a.cap;
            if (n > a.cap - BN && !sm.flag[rl]) {
                double tau;
                int kept;
                prune_row(cd + (size_t)rl * a.cap, cj + (size_t)rl * a.cap, n, a.k, sm.nrm[rl], a.mcoef,
                          sm.hist[warp], lane, &tau, &kept);
                if (lane == 0) {
                    if (kept > a.cap - BN) {       // a tie plateau wider than the buffer: exact fallback
                        sm.flag[rl] = 1;
                        sm.tau[rl] = __longlong_as_double(0x7ff8000000000000LL);
                        sm.cnt[rl] = 0;
                    } else {
                        sm.tau[rl] = tau;
                        sm.cnt[rl] = kept;
                    }
                }
                __syncwarp();
            }
        }
        named_bar_sync(1, CONSUMER_THREADS);
    }
    if (ctid < BM && cur_rb >= 0) {
        a.seg_cnt[(size_t)seg * BM + ctid] = sm.cnt[ctid] > a.cap ? a.cap : sm.cnt[ctid];
        a.seg_flag[(size_t)seg * BM + ctid] = sm.flag[ctid];
    }
}

// ---------------------------------------------------------------------------------------------------------
// K4: centre at 1.0 and row norms
// ---------------------------------------------------------------------------------------------------------
__global__ void wc_center_norms_kernel(const double* __restrict__ X, int N, int S, int ld, double* __restrict__ Xc,
                                       double* __restrict__ norms) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= N) return;
    const double* src = X + (size_t)warp * S;
    double* dst = Xc + (size_t)warp * ld;
    double acc = 0.0;
    for (int s = lane; s < ld; s += 32) {
        double v = s < S ? src[s] - 1.0 : 0.0;
        dst[s] = v;
        acc = fma(v, v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norms[warp] = acc;
}

__global__ void wc_fill_f64_kernel(double* p, size_t n, double v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------
// K6: per-row finalize
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_less(double da, int ja, double db, int jb) {
    return da < db || (da == db && ja < jb);
}

__device__ void bitonic_sort_pairs(double* key, int* val, int m, int tid, int nthreads) {
    for (int size = 2; size <= m; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (m >> 1); t += nthreads) {
                int i = ((t / stride) * stride << 1) + (t % stride);
                int j = i + stride;
                bool up = (i & size) == 0;
                double ki = key[i], kj = key[j];
                int vi = val[i], vj = val[j];
                bool swap = up ? pair_less(kj, vj, ki, vi) : pair_less(ki, vi, kj, vj);
                if (swap) {
                    key[i] = kj; key[j] = ki;
                    val[i] = vj; val[j] = vi;
                }
            }
            __syncthreads();
        }
    }
}

struct FinArgs {
    const double* X;        // original corrected data, N x S
    int N, S;
    const double* norms;
    const int* row_cs;
    const int* row_ce;
    int row_begin, row_end;
    const int* rb_seg_first;
    const int* rb_seg_count;
    const double* cand_d;
    const int* cand_j;
    const int* seg_cnt;
    const int* seg_flag;
    int cap;
    int k;
    double mcoef;
    int* idx_out;
    double* dist_out;
    int* slow_list;
    int* slow_count;
};

__global__ void __launch_bounds__(FIN_THREADS) wc_finalize_kernel(const FinArgs a) {
    extern __shared__ unsigned char fin_raw[];
    double* key = reinterpret_cast<double*>(fin_raw);                      // FIN_MAX
    int* val = reinterpret_cast<int*>(key + FIN_MAX);                      // FIN_MAX
    double* tile = reinterpret_cast<double*>(val + FIN_MAX);               // 128 x 33
    double* xi = tile + 128 * 33;                                          // 32
    double* ex_d = xi + 32;                                                // FIN_SHORT
    int* ex_j = reinterpret_cast<int*>(ex_d + FIN_SHORT);                  // FIN_SHORT
    __shared__ int s_total, s_flag, s_p;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rloc = blockIdx.x;
    const int row = a.row_begin + rloc;
    const int rb = rloc / BM, rl = rloc % BM;
    const int seg0 = a.rb_seg_first[rb], nseg = a.rb_seg_count[rb];
    int* out_i = a.idx_out + (size_t)rloc * a.k;
    double* out_d = a.dist_out + (size_t)rloc * a.k;

    if (tid == 0) {
        int tot = 0, fl = 0;
        for (int s = 0; s < nseg; ++s) {
            tot += a.seg_cnt[(size_t)(seg0 + s) * BM + rl];
            fl |= a.seg_flag[(size_t)(seg0 + s) * BM + rl];
        }
        s_total = tot;
        s_flag = fl;
    }
    __syncthreads();
    const int total = s_total;
    if (s_flag || total > FIN_MAX) {
        if (tid == 0) a.slow_list[atomicAdd(a.slow_count, 1)] = rloc;
        return;
    }
    int m = 2;
    while (m < total) m <<= 1;
    {
        int base = 0;
        for (int s = 0; s < nseg; ++s) {
            const int n = a.seg_cnt[(size_t)(seg0 + s) * BM + rl];
            const size_t off = ((size_t)(seg0 + s) * BM + rl) * a.cap;
            for (int e = tid; e < n; e += FIN_THREADS) {
                key[base + e] = a.cand_d[off + e];
                val[base + e] = a.cand_j[off + e];
            }
            base += n;
        }
        for (int e = total + tid; e < m; e += FIN_THREADS) {
            key[e] = INFINITY;
            val[e] = 0x7fffffff;
        }
    }
    __syncthreads();
    bitonic_sort_pairs(key, val, m, tid, FIN_THREADS);

    // shortlist: everything within the error margin of the k-th smallest approximate distance
    const int kk = total < a.k ? total : a.k;
    int p = 0;
    if (kk > 0) {
        const double tk = key[kk - 1];
        const double window = tk + a.mcoef * (a.norms[row] + fabs(tk));
        if (tid == 0) s_p = 0;
        __syncthreads();
        int local = 0;
        for (int e = tid; e < total; e += FIN_THREADS) local += key[e] <= window ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0) atomicAdd(&s_p, local);
        __syncthreads();
        p = s_p;
    }
    if (p > FIN_SHORT) {
        if (tid == 0) a.slow_list[atomicAdd(a.slow_count, 1)] = rloc;
        return;
    }

    // exact re-score in the reference's operation order: sequential over samples, separately rounded
    // subtract, multiply, add (wisetools.py:302 on its Fortran-ordered operands)
    const double* xrow = a.X + (size_t)row * a.S;
    for (int pass = 0; pass * 128 < p; ++pass) {
        const int c0 = pass * 128;
        const int nc = (p - c0) < 128 ? (p - c0) : 128;
        double accd = 0.0;
        for (int s0 = 0; s0 < a.S; s0 += 32) {
            const int ns = (a.S - s0) < 32 ? (a.S - s0) : 32;
            for (int c = warp; c < nc; c += FIN_THREADS / 32) {
                const double* src = a.X + (size_t)val[c0 + c] * a.S + s0;
                if (lane < ns) tile[c * 33 + lane] = src[lane];
            }
            if (warp == 0 && lane < ns) xi[lane] = xrow[s0 + lane];
            __syncthreads();
            if (tid < nc) {
                const double* tr = tile + tid * 33;
                for (int l = 0; l < ns; ++l) {
                    double v = __dsub_rn(tr[l], xi[l]);
                    accd = __dadd_rn(accd, __dmul_rn(v, v));
                }
            }
            __syncthreads();
        }
        if (tid < nc) {
            const bool ok = accd < 1e10;     // wisetools.py:312-314: strict `<` against the 1e10 start value; NaN fails
            ex_d[c0 + tid] = ok ? accd : INFINITY;
            ex_j[c0 + tid] = ok ? val[c0 + tid] : 0x7fffffff;
        }
    }
    int m2 = 2;
    while (m2 < p) m2 <<= 1;
    for (int e = p + tid; e < m2; e += FIN_THREADS) {
        ex_d[e] = INFINITY;
        ex_j[e] = 0x7fffffff;
    }
    __syncthreads();
    if (p > 0) bitonic_sort_pairs(ex_d, ex_j, m2, tid, FIN_THREADS);
    const int cs = a.row_cs[row], ce = a.row_ce[row];
    for (int e = tid; e < a.k; e += FIN_THREADS) {
        int oi = -1;
        double od = 1e10;
        if (e < p && ex_j[e] != 0x7fffffff) {
            const int j = ex_j[e];
            oi = j >= ce ? j - (ce - cs) : j;
            od = ex_d[e];
        }
        out_i[e] = oi;
        out_d[e] = od;
    }
}

// ---------------------------------------------------------------------------------------------------------
// K6b: exhaustive exact search for one row per CTA (rare)
// ---------------------------------------------------------------------------------------------------------
struct ExhArgs {
    const double* X;
    int N, S;
    const int* row_cs;
    const int* row_ce;
    int row_begin;
    const int* slow_list;
    int list_off;
    double* scratch;     // [rows in this batch][N]
    int k;
    int* idx_out;
    double* dist_out;
};

__global__ void __launch_bounds__(EXH_THREADS) wc_exhaustive_kernel(const ExhArgs a) {
    __shared__ double s_d[EXH_THREADS / 32];
    __shared__ int s_j[EXH_THREADS / 32];
    __shared__ double s_bd;
    __shared__ int s_bj;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rloc = a.slow_list[a.list_off + blockIdx.x];
    const int row = a.row_begin + rloc;
    const int cs = a.row_cs[row], ce = a.row_ce[row];
    double* dd = a.scratch + (size_t)blockIdx.x * a.N;
    const double* xrow = a.X + (size_t)row * a.S;
    for (int j = tid; j < a.N; j += EXH_THREADS) {
        double acc = INFINITY;
        if (j < cs || j >= ce) {
            const double* xj = a.X + (size_t)j * a.S;
            acc = 0.0;
            for (int s = 0; s < a.S; ++s) {
                double v = __dsub_rn(xj[s], xrow[s]);
                acc = __dadd_rn(acc, __dmul_rn(v, v));
            }
            if (!(acc < 1e10)) acc = INFINITY;
        }
        dd[j] = acc;
    }
    __syncthreads();
    double last_d = -INFINITY;
    int last_j = -1;
    int* out_i = a.idx_out + (size_t)rloc * a.k;
    double* out_d = a.dist_out + (size_t)rloc * a.k;
    for (int t = 0; t < a.k; ++t) {
        double bd = INFINITY;
        int bj = 0x7fffffff;
        for (int j = tid; j < a.N; j += EXH_THREADS) {
            double d = dd[j];
            bool after = d > last_d || (d == last_d && j > last_j);
            if (after && pair_less(d, j, bd, bj)) { bd = d; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(0xffffffffu, bd, o);
            int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (pair_less(od, oj, bd, bj)) { bd = od; bj = oj; }
        }
        if (lane == 0) { s_d[warp] = bd; s_j[warp] = bj; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < EXH_THREADS / 32; ++w)
                if (pair_less(s_d[w], s_j[w], bd, bj)) { bd = s_d[w]; bj = s_j[w]; }
            s_bd = bd;
            s_bj = bj;
            if (bd < INFINITY) {
                out_i[t] = bj >= ce ? bj - (ce - cs) : bj;
                out_d[t] = bd;
            } else {
                out_i[t] = -1;
                out_d[t] = 1e10;
            }
        }
        __syncthreads();
        last_d = s_bd;
        last_j = s_bj;
        if (!(last_d < INFINITY)) {          // everything left is a filler
            for (int e = t + 1 + tid; e < a.k; e += EXH_THREADS) { out_i[e] = -1; out_d[e] = 1e10; }
            break;
        }
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum {
    SLOT_XC = 0, SLOT_NORMS, SLOT_ROWCS, SLOT_ROWCE, SLOT_RBMETA, SLOT_CAND_D, SLOT_CAND_J, SLOT_SEGCNT,
    SLOT_SEGFLAG, SLOT_SLOW, SLOT_SCRATCH, SLOT_IO_X, SLOT_IO_IDX, SLOT_IO_DIST
};

}  // namespace

extern "C" int wc_newref_topk(wc_ctx* ctx, const double* corrected_d, int N, int S, const int* chrom_bins_h,
                              int nchrom, int row_begin, int row_end, int refsize, int32_t* idx_d, double* dist_d,
                              void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr);
    WC_CHECK_ARG(corrected_d != nullptr && chrom_bins_h != nullptr);
    WC_CHECK_ARG(N > 0 && S > 0 && nchrom > 0);
    WC_CHECK_ARG(refsize >= 1 && refsize <= 384);
    WC_CHECK_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= N);
    long long tot = 0;
    for (int c = 0; c < nchrom; ++c) {
        WC_CHECK_ARG(chrom_bins_h[c] >= 0);
        tot += chrom_bins_h[c];
    }
    WC_CHECK_ARG(tot == N);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    for (int i = 0; i < WC_NPHASE; ++i) ctx->phase_ms[i] = 0.0;
    for (int i = 0; i < WC_NCOUNTER; ++i) ctx->counter[i] = 0;
    const int rows = row_end - row_begin;
    if (rows == 0) return WC_OK;
    WC_CHECK_ARG(idx_d != nullptr && dist_d != nullptr);

    const int k = refsize;
    const int cap = k <= 128 ? 512 : 1024;
    const int ld = (S + BK - 1) / BK * BK;
    const int nkc = ld / BK;
    const int last_valid = S - (nkc - 1) * BK;            // 1..16 valid samples in the last chunk
    const int nsteps_last = (last_valid + 3) / 4;
    const size_t Npad = (size_t)(N + BN - 1) / BN * BN + BN;
    const double mcoef = 16.0 * (double)(S + 8) * 1.1102230246251565e-16;

    // ---- per-row exclusion ranges and the (row block, column tile) work list -----------------------------
    std::vector<int> row_cs(N), row_ce(N);
    {
        int pos = 0;
        for (int c = 0; c < nchrom; ++c) {
            for (int i = 0; i < chrom_bins_h[c]; ++i) {
                row_cs[pos + i] = pos;
                row_ce[pos + i] = pos + chrom_bins_h[c];
            }
            pos += chrom_bins_h[c];
        }
    }
    const int nrb = (rows + BM - 1) / BM;
    const int ntile_cols = (N + BN - 1) / BN;
    std::vector<int> prefix(nrb + 1), skip_lo(nrb), skip_n(nrb);
    prefix[0] = 0;
    for (int rb = 0; rb < nrb; ++rb) {
        int r0 = row_begin + rb * BM;
        int r1 = std::min(row_end, r0 + BM) - 1;
        int lo = ntile_cols, n = 0;
        if (row_cs[r0] == row_cs[r1]) {   // block inside one chromosome: its interior column tiles are skipped
            int first = (row_cs[r0] + BN - 1) / BN;
            int last = row_ce[r0] / BN;    // tiles [first, last) lie wholly inside [cs, ce)
            if (last > first) { lo = first; n = last - first; }
        }
        skip_lo[rb] = lo;
        skip_n[rb] = n;
        prefix[rb + 1] = prefix[rb] + ntile_cols - n;
    }
    const int total_tiles = prefix[nrb];
    int grid = ctx->sm_count;
    if (grid > (total_tiles + 7) / 8) grid = (total_tiles + 7) / 8;
    if (grid < 1) grid = 1;
    std::vector<int> cta_seg_base(grid), rb_seg_first(nrb, -1), rb_seg_count(nrb, 0);
    int nseg = 0;
    {
        int rb = 0;
        for (int c = 0; c < grid; ++c) {
            int lin0 = (int)((long long)total_tiles * c / grid);
            int lin1 = (int)((long long)total_tiles * (c + 1) / grid);
            cta_seg_base[c] = nseg;
            int cur = -1;
            for (int lin = lin0; lin < lin1;) {
                while (lin >= prefix[rb + 1]) ++rb;
                if (rb != cur) {
                    if (rb_seg_first[rb] < 0) rb_seg_first[rb] = nseg;
                    rb_seg_count[rb]++;
                    ++nseg;
                    cur = rb;
                }
                lin = std::min(lin1, prefix[rb + 1]);   // jump to the end of this row block's share
            }
        }
    }
    for (int rb = 0; rb < nrb; ++rb)
        if (rb_seg_first[rb] < 0) { rb_seg_first[rb] = 0; rb_seg_count[rb] = 0; }

    // ---- workspace ----------------------------------------------------------------------------------------
    double* Xc; double* norms; int* d_row_cs; int* d_row_ce; int* d_meta;
    double* cand_d; int* cand_j; int* seg_cnt; int* seg_flag; int* slow;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_XC, Npad * ld * sizeof(double), (void**)&Xc))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_NORMS, Npad * sizeof(double), (void**)&norms))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_ROWCS, (size_t)N * sizeof(int), (void**)&d_row_cs))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_ROWCE, (size_t)N * sizeof(int), (void**)&d_row_ce))) return rc;
    const size_t meta_ints = (size_t)(nrb + 1) + 4 * (size_t)nrb + grid;
    if ((rc = wc_reserve(ctx, SLOT_RBMETA, meta_ints * sizeof(int), (void**)&d_meta))) return rc;
    const size_t cand_n = (size_t)std::max(nseg, 1) * BM * cap;
    if ((rc = wc_reserve(ctx, SLOT_CAND_D, cand_n * sizeof(double), (void**)&cand_d))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_CAND_J, cand_n * sizeof(int), (void**)&cand_j))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_SEGCNT, (size_t)std::max(nseg, 1) * BM * sizeof(int), (void**)&seg_cnt))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_SEGFLAG, (size_t)std::max(nseg, 1) * BM * sizeof(int), (void**)&seg_flag))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_SLOW, ((size_t)rows + 1) * sizeof(int), (void**)&slow))) return rc;
    int* d_prefix = d_meta;
    int* d_skip_lo = d_prefix + (nrb + 1);
    int* d_skip_n = d_skip_lo + nrb;
    int* d_seg_first = d_skip_n + nrb;
    int* d_seg_count = d_seg_first + nrb;
    int* d_cta_seg = d_seg_count + nrb;

    WC_CUDA(cudaMemcpyAsync(d_row_cs, row_cs.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_row_ce, row_ce.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_prefix, prefix.data(), (size_t)(nrb + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_skip_lo, skip_lo.data(), (size_t)nrb * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_skip_n, skip_n.data(), (size_t)nrb * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_seg_first, rb_seg_first.data(), (size_t)nrb * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_seg_count, rb_seg_count.data(), (size_t)nrb * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_cta_seg, cta_seg_base.data(), (size_t)grid * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemsetAsync(slow, 0, sizeof(int), stream));
    WC_CUDA(cudaMemsetAsync(seg_cnt, 0, (size_t)std::max(nseg, 1) * BM * sizeof(int), stream));
    WC_CUDA(cudaMemsetAsync(seg_flag, 0, (size_t)std::max(nseg, 1) * BM * sizeof(int), stream));

    // ---- K4 -------------------------------------------------------------------------------------------------
    WC_CUDA(cudaEventRecord(ctx->ev[0], stream));
    WC_CUDA(cudaMemsetAsync(Xc + (size_t)N * ld, 0, (Npad - N) * ld * sizeof(double), stream));
    {
        size_t n = Npad - N;
        wc_fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(norms + N, n, NAN);
        int blocks = (N * 32 + 255) / 256;
        wc_center_norms_kernel<<<blocks, 256, 0, stream>>>(corrected_d, N, S, ld, Xc, norms);
    }
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[1], stream));

    // ---- K5 -------------------------------------------------------------------------------------------------
    CUtensorMap tmap;
    {
        if (!ctx->encode_tiled) {
            wc_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return WC_ERR_CUDA;
        }
        cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)Npad};
        cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
        cuuint32_t box[2] = {BK, BM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = reinterpret_cast<PFN_encodeTiled>(ctx->encode_tiled)(
            &tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, Xc, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            wc_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return WC_ERR_CUDA;
        }
    }
    TopkArgs ta;
    ta.norms = norms; ta.row_cs = d_row_cs; ta.row_ce = d_row_ce; ta.N = N;
    ta.row_begin = row_begin; ta.row_end = row_end; ta.nkc = nkc; ta.nsteps_last = nsteps_last;
    ta.rb_tile_prefix = d_prefix; ta.rb_skip_lo = d_skip_lo; ta.rb_skip_n = d_skip_n; ta.nrb = nrb;
    ta.total_tiles = total_tiles; ta.cta_seg_base = d_cta_seg;
    ta.cand_d = cand_d; ta.cand_j = cand_j; ta.seg_cnt = seg_cnt; ta.seg_flag = seg_flag;
    ta.cap = cap; ta.k = k; ta.mcoef = mcoef; ta.tau_init = 1e10 * (1.0 + 1e-6);
    const size_t topk_smem = sizeof(TopkSmem);
    WC_CUDA(cudaFuncSetAttribute(wc_dist_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topk_smem));
    WC_CUDA(cudaEventRecord(ctx->ev[2], stream));
    wc_dist_topk_kernel<<<grid, TOPK_THREADS, topk_smem, stream>>>(tmap, ta);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[3], stream));

    // ---- K6 -------------------------------------------------------------------------------------------------
    FinArgs fa;
    fa.X = corrected_d; fa.N = N; fa.S = S; fa.norms = norms; fa.row_cs = d_row_cs; fa.row_ce = d_row_ce;
    fa.row_begin = row_begin; fa.row_end = row_end; fa.rb_seg_first = d_seg_first; fa.rb_seg_count = d_seg_count;
    fa.cand_d = cand_d; fa.cand_j = cand_j; fa.seg_cnt = seg_cnt; fa.seg_flag = seg_flag; fa.cap = cap; fa.k = k;
    fa.mcoef = mcoef; fa.idx_out = idx_d; fa.dist_out = dist_d; fa.slow_list = slow + 1; fa.slow_count = slow;
    const size_t fin_smem = FIN_MAX * 12 + (128 * 33 + 32) * 8 + FIN_SHORT * 12;
    WC_CUDA(cudaFuncSetAttribute(wc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    WC_CUDA(cudaEventRecord(ctx->ev[4], stream));
    wc_finalize_kernel<<<rows, FIN_THREADS, fin_smem, stream>>>(fa);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[5], stream));

    int nslow = 0;
    WC_CUDA(cudaMemcpyAsync(&nslow, slow, sizeof(int), cudaMemcpyDeviceToHost, stream));
    WC_CUDA(cudaStreamSynchronize(stream));
    long long launches = 4;
    if (nslow > 0) {
        const int batch = 64;
        double* scratch;
        if ((rc = wc_reserve(ctx, SLOT_SCRATCH, (size_t)std::min(nslow, batch) * N * sizeof(double), (void**)&scratch)))
            return rc;
        WC_CUDA(cudaEventRecord(ctx->ev[6], stream));
        for (int off = 0; off < nslow; off += batch) {
            ExhArgs ea;
            ea.X = corrected_d; ea.N = N; ea.S = S; ea.row_cs = d_row_cs; ea.row_ce = d_row_ce; ea.row_begin = row_begin;
            ea.slow_list = slow + 1; ea.list_off = off; ea.scratch = scratch; ea.k = k; ea.idx_out = idx_d; ea.dist_out = dist_d;
            wc_exhaustive_kernel<<<std::min(batch, nslow - off), EXH_THREADS, 0, stream>>>(ea);
            ++launches;
        }
        WC_CUDA(cudaGetLastError());
        WC_CUDA(cudaEventRecord(ctx->ev[7], stream));
        WC_CUDA(cudaStreamSynchronize(stream));
    }
    float ms;
    WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->phase_ms[0] = ms;
    WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3])); ctx->phase_ms[1] = ms;
    WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5])); ctx->phase_ms[2] = ms;
    if (nslow > 0) { WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7])); ctx->phase_ms[3] = ms; }
    ctx->counter[0] = launches;
    ctx->counter[1] = nslow;
    ctx->counter[3] = total_tiles;
    ctx->counter[4] = grid;
    return WC_OK;
}

extern "C" int wc_newref_topk_host(wc_ctx* ctx, const double* corrected_h, int N, int S, const int* chrom_bins_h,
                                   int nchrom, int row_begin, int row_end, int refsize, int32_t* idx_h,
                                   double* dist_h) {
    WC_CHECK_ARG(ctx != nullptr && corrected_h != nullptr);
    WC_CHECK_ARG(N > 0 && S > 0 && refsize >= 1);
    WC_CHECK_ARG(row_begin >= 0 && row_begin <= row_end && row_end <= N);
    WC_CUDA(cudaSetDevice(ctx->device));
    const size_t rows = (size_t)(row_end - row_begin);
    double* X; int32_t* idx; double* dist;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_IO_X, (size_t)N * S * sizeof(double), (void**)&X))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_IO_IDX, std::max<size_t>(rows, 1) * refsize * sizeof(int32_t), (void**)&idx))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_IO_DIST, std::max<size_t>(rows, 1) * refsize * sizeof(double), (void**)&dist))) return rc;
    WC_CUDA(cudaMemcpyAsync(X, corrected_h, (size_t)N * S * sizeof(double), cudaMemcpyHostToDevice, 0));
    rc = wc_newref_topk(ctx, X, N, S, chrom_bins_h, nchrom, row_begin, row_end, refsize, idx, dist, nullptr);
    if (rc) return rc;
    if (rows) {
        WC_CHECK_ARG(idx_h != nullptr && dist_h != nullptr);
        WC_CUDA(cudaMemcpyAsync(idx_h, idx, rows * refsize * sizeof(int32_t), cudaMemcpyDeviceToHost, 0));
        WC_CUDA(cudaMemcpyAsync(dist_h, dist, rows * refsize * sizeof(double), cudaMemcpyDeviceToHost, 0));
        WC_CUDA(cudaStreamSynchronize(0));
    }
    return WC_OK;
}
