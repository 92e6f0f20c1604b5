// wisecondor_b200 - experimental FP16 tensor-core filter of the reference-bin search (K4h + K5h).
// Textually included by wc_search.cu inside its anonymous namespace: uses TopkArgs, TopkState, prune_row and the PTX helpers
// defined there.
#pragma once

// ---------------------------------------------------------------------------------------------------------
// K4h / K5h: the same search with an FP16 tensor-core filter (option "k5_f16"; off until measured)
// ---------------------------------------------------------------------------------------------------------
// The filter only has to produce a SUPERSET of every bin's true top-k - K6 re-scores the shortlist exactly in fp64 - so
// the contraction does not need fp64 at all.  With x' = x - 1 rounded once to fp16 (relative error 2^-11 per operand)
// and fp32 accumulation, |d~ - d| <= eps * (n_i + n_j) with eps ~ 1.1e-3 (a priori; tools/filter_precision_study.py measures the
// candidate inflation of such a margin: 109 instead of 100 candidates per bin at 600 x 250 kb).  The norms are NOT part
// of the contraction here (n/2 ~ 1 would lose all precision in fp16): d~ = (n_i + n_j) - 2 s in fp32 in the epilogue.
// Same persistent grid, TMA ring, warp-private rows, candidate buffers, prunes and symmetric column side as K5; the
// operand tile has the same bytes (128 rows x 64 halves = 128 rows x 128 B, SWIZZLE_128B), fragments come from
// ldmatrix.x4, the MMA is mma.sync.m16n8k16.f16 with fp32 accumulators (SASS HMMA.16816.F32).
constexpr int BKH = 64;             // halves per pipeline stage and operand row (128 bytes)
constexpr int F16_SCRATCH = 4096;   // per-warp parking area of the rare path (32 lanes x 32 fp32)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// threshold key -> the fp32 bound used by the per-entry compare, rounded up (conservative)
__device__ __forceinline__ float tau32_of_key(u64 key) { return __double2float_ru(dist_of_key(key)); }

// K4h: X' = X - 1 in fp16 (padded with zeros to whole 64-sample chunks), n_i in fp64 (margins) and fp32 (epilogue;
// +inf for padding rows so that they never pass), the largest finite norm and a flag for values fp16 cannot hold.
// fold (K5t): the norms ride in the contraction.  The last four columns of the last 64-sample chunk hold (h_i, l_i, 1, 1)
// with h_i + l_i = -n_i / 2 split into two halves (relative error 2^-22), and a second copy of that chunk, 64 columns further
// (row stride ldx = ldh + 64), holds (1, 1, h_i, l_i) there: with the A operand reading the first version and the B operand the
// second, the accumulator is x'_i . x'_j - n_i / 2 - n_j / 2 = -d~_ij / 2 and the epilogue needs no arithmetic before its
// compares.  Needs S + 4 <= ldh and n / 2 inside fp16's range (else the range flag: fp64 filter).
__global__ void wc_prepare_f16_kernel(const double* __restrict__ X, int N, int Npad, int S, int ldh, int ldx, int fold,
                                      __half* __restrict__ Xh, double* __restrict__ norms, float* __restrict__ n32,
                                      unsigned long long* __restrict__ stats) {
    const int row = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= Npad) return;
    __half* dst = Xh + (size_t)row * ldx;
    double acc = 0.0;
    bool big = false;
    if (row < N) {
        const double* src = X + (size_t)row * S;
        for (int s = lane; s < ldh; s += 32) {
            const double v = s < S ? src[s] - 1.0 : 0.0;
            const __half hv = __double2half(v);
            dst[s] = hv;
            if (fold && s >= ldh - BKH) dst[s + BKH] = hv;
            acc = fma(v, v, acc);
            if (fabs(v) > 60000.0 && fabs(v) < INFINITY) big = true;
        }
    } else {
        for (int s = lane; s < ldx; s += 32) dst[s] = __float2half(0.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    big = __any_sync(0xffffffffu, big);
    __syncwarp();                                   // the zeros of the spare columns are written: lane 0 overwrites four of them
    if (lane == 0) {
        norms[row] = row < N ? acc : 0.0;
        n32[row] = row < N ? (float)acc : INFINITY;
        if (fold && row < N) {
            const double nh = -0.5 * acc;
            if (!(acc * 0.5 <= 60000.0)) big = true;
            const __half h = __double2half(nh);
            const __half l = __double2half(nh - (double)__half2float(h));
            const __half one = __float2half(1.0f);
            dst[ldh - 4] = h;   dst[ldh - 3] = l;   dst[ldh - 2] = one; dst[ldh - 1] = one;      // as the A operand
            dst[ldx - 4] = one; dst[ldx - 3] = one; dst[ldx - 2] = h;   dst[ldx - 1] = l;        // as the B operand
        }
        if (row < N && acc < INFINITY) atomicMax(stats, (unsigned long long)__double_as_longlong(acc));   // acc >= 0
        if (big) atomicOr(stats + 1, 1ull);
    }
}

template <bool SYM>
__global__ void __launch_bounds__(TOPK_THREADS, 1)
wc_dist_topk_f16_kernel(const __grid_constant__ CUtensorMap tmap, const TopkArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();
    const int STAGES = a.nstages;
    unsigned char* tiles = smem_raw;
    TopkState& sm = *reinterpret_cast<TopkState*>(smem_raw + (size_t)STAGES * STAGE_BYTES);
    unsigned char* scratch = reinterpret_cast<unsigned char*>(&sm + 1);
    const int tid = threadIdx.x;
    const int warp_all = tid >> 5, lane = tid & 31;
    const int warp = warp_all - PRODUCER_WARPS;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], CONSUMER_WARPS);
        }
        for (int w = 0; w < CONSUMER_WARPS; ++w) sm.stg_cnt[w] = 0;
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    const int pb = a.cta_piece_begin[blockIdx.x], pe = a.cta_piece_begin[blockIdx.x + 1];
    if (pb >= pe) return;

    if (warp_all < PRODUCER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp_all == 0 && lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pi = pb; pi < pe; ++pi) {
                const int* pc = a.pieces + (size_t)pi * 5;
                const int rbp = pc[0], q1 = pc[2], qs = pc[3];
                const int skip_lo = a.rb_skip_lo[rbp], skip_n = a.rb_skip_n[rbp];
                const int* tl = a.tile_list ? a.tile_list + a.rb_list_off[rbp] : nullptr;
                const int row0 = a.row_begin + rbp * BM;
                for (int q = pc[1]; q < q1; q += qs) {
                    const int t = tl ? tl[q] : (q < skip_lo ? q : q + skip_n);
                    const int col0 = t * BN;
                    for (int kc = 0; kc < a.nkc; ++kc) {
                        mbar_wait(&sm.empty[stage], phase ^ 1u);
                        mbar_arrive_expect_tx(&sm.full[stage], STAGE_BYTES);
                        tma_load_2d(tiles + (size_t)stage * STAGE_BYTES, &tmap, kc * BKH, row0, &sm.full[stage]);
                        tma_load_2d(tiles + (size_t)stage * STAGE_BYTES + TILE_BYTES, &tmap, kc * BKH, col0, &sm.full[stage]);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        return;
    }

    // ===== consumers: warp w owns rows [16w, 16w+16) x 128 columns = 16 m16n8 accumulator tiles (64 fp32 per lane) =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // m16n8k16 fragments: lane (g = lane/4, q = lane%4) holds C rows g and g+8, columns 2q, 2q+1 of every n-tile.
    // ldmatrix.x4 addresses (A: matrices = rows 0-7 / 8-15 x k 0-7 / 8-15; B: n 0-7 / 8-15 x k 0-7 / 8-15 of an n-tile pair);
    // the TMA 128-byte swizzle XORs the 16-byte chunk index with (row & 7).
    const int g = lane >> 2, q4 = lane & 3;
    const uint32_t a_row = (uint32_t)(warp * WROWS + (lane & 7) + ((lane >> 3) & 1) * 8);
    const uint32_t a_off = a_row * 128u;
    const uint32_t a_kc = (uint32_t)(lane >> 4);            // which 8-sample half of the k16 step this lane addresses
    const uint32_t b_row = (uint32_t)(((lane >> 4) & 1) * 8 + (lane & 7));
    const uint32_t b_off = (uint32_t)TILE_BYTES + b_row * 128u;
    const uint32_t b_kc = (uint32_t)((lane >> 3) & 1);
    const uint32_t xr = (uint32_t)(lane & 7);               // row & 7 of both addresses

    const uint32_t tiles_u32 = smem_u32(tiles);
    const size_t scratch_per_warp = F16_SCRATCH;       // the prune works on registers: only the parking area is needed
    u64* w_sk = reinterpret_cast<u64*>(scratch + (size_t)warp * scratch_per_warp);
    int* w_sj = nullptr;
    u64* w_ct = reinterpret_cast<u64*>(scratch + (size_t)CONSUMER_WARPS * scratch_per_warp) + warp * BN;
    uint4* w_stg = reinterpret_cast<uint4*>(scratch + (size_t)CONSUMER_WARPS * scratch_per_warp +
                                            (size_t)CONSUMER_WARPS * BN * sizeof(u64)) + warp * STG;
    // fp32 side tables of the warp: norms of the tile's 128 columns, their thresholds (SYM), thresholds of the 16 rows
    float* w_cn = reinterpret_cast<float*>(scratch + (size_t)CONSUMER_WARPS * scratch_per_warp +
                                           (size_t)CONSUMER_WARPS * (BN * sizeof(u64) + STG * sizeof(uint4))) + warp * (2 * BN + 32);
    float* w_ctf = w_cn + BN;
    float* w_rt = w_ctf + BN;
    int* w_stgc = &sm.stg_cnt[warp];
    const int r0w = warp * WROWS;
    u64* w_thr = sm.thr + r0w;
    double* w_nrm = sm.nrm + r0w;
    int* w_cnt = sm.cnt + r0w;
    unsigned char* w_flag = sm.flag + r0w;
    int my_cs[2] = {0, 0}, my_ce[2] = {0, 0};
    float my_n[2] = {INFINITY, INFINITY};       // fp32 norms of this lane's two rows (g and g+8)
    int stage = 0;
    uint32_t phase = 0;
    const size_t seg_stride = (size_t)BM * a.cap;
    bool ready = false;

    long long pf_wait = 0, pf_epi = 0, pf_prune = 0, pf_nprune = 0, pf_emit = 0;
    const long long pf_t0 = clock64();
    int pi = pb;
    const int* pc = a.pieces + (size_t)pi * 5;
    int rb = pc[0], q = pc[1], q1 = pc[2], qs = pc[3], seg = pc[4];
    int skip_lo = a.rb_skip_lo[rb], skip_n = a.rb_skip_n[rb];
    const int* tl = a.tile_list ? a.tile_list + a.rb_list_off[rb] : nullptr;
    bool new_piece = true;

    auto prune_rows = [&](unsigned need, u64* ck, int* cj) {
        while (need) {
            const int rw = __ffs(need) - 1;
            need &= need - 1;
            int n = w_cnt[rw];
            if (n > a.cap) n = a.cap;
            u64 thr;
            int kept;
            ++pf_nprune;
            u64* rk = ck + (size_t)(r0w + rw) * a.cap;
            int* rj = cj + (size_t)(r0w + rw) * a.cap;
            __threadfence_block();
            if (a.cap <= 512)
                prune_row<16>(rk, rj, n, a.k, w_nrm[rw], a.mcoef, madd_of(a), lane, w_sk, w_sj, &thr, &kept);
            else
                prune_row<32>(rk, rj, n, a.k, w_nrm[rw], a.mcoef, madd_of(a), lane, w_sk, w_sj, &thr, &kept);
            if (lane == 0) {
                if (kept > a.cap - BN) {
                    w_flag[rw] = 1;
                    w_thr[rw] = KEY_NEVER;
                    w_cnt[rw] = 0;
                } else {
                    const int row = a.row_begin + rb * BM + r0w + rw;
                    const u64 other = atomicMin(a.row_thr + (row - a.row_begin), thr);
                    w_thr[rw] = other < thr ? other : thr;
                    w_cnt[rw] = kept;
                }
            }
            __syncwarp();
        }
    };
    auto flush_incoming = [&]() {
        __syncwarp();
        int n = *w_stgc;
        if (n > STG) n = STG;
        for (int e = lane; e < n; e += 32) {
            const uint4 v = w_stg[e];
            const int j = (int)v.z;
            const int w = atomicAdd(a.in_cnt + j, 1);
            if (w < a.in_cap) {
                a.in_key[(size_t)j * a.in_cap + w] = ((u64)v.y << 32) | (u64)v.x;
                a.in_j[(size_t)j * a.in_cap + w] = (int)v.w;
            }
        }
        __syncwarp();
        if (lane == 0) *w_stgc = 0;
        __syncwarp();
    };
    int tcount = 0;
    while (true) {
        if (q >= q1) {
            __syncwarp();
            if (SYM) flush_incoming();
            if (a.final_prune) {
                const unsigned need = __ballot_sync(0xffffffffu, lane < WROWS && w_cnt[lane] > a.k + 24 &&
                                                                     w_cnt[lane] <= a.cap && !w_flag[lane]);
                prune_rows(need, a.cand_key + (size_t)seg * seg_stride, a.cand_j + (size_t)seg * seg_stride);
            }
            if (lane < WROWS) {
                a.seg_cnt[(size_t)seg * BM + r0w + lane] = w_cnt[lane] > a.cap ? a.cap : w_cnt[lane];
                a.seg_flag[(size_t)seg * BM + r0w + lane] = w_flag[lane];
            }
            if (++pi >= pe) break;
            pc = a.pieces + (size_t)pi * 5;
            rb = pc[0]; q = pc[1]; q1 = pc[2]; qs = pc[3]; seg = pc[4];
            skip_lo = a.rb_skip_lo[rb]; skip_n = a.rb_skip_n[rb];
            tl = a.tile_list ? a.tile_list + a.rb_list_off[rb] : nullptr;
            new_piece = true;
            continue;
        }
        if (new_piece) {
            new_piece = false;
            __syncwarp();
            if (lane < WROWS) {
                const int row = a.row_begin + rb * BM + r0w + lane;
                const bool valid = row < a.row_end;
                w_nrm[lane] = valid ? a.norms[row] : 0.0;
                w_thr[lane] = valid ? __ldcg(a.row_thr + (row - a.row_begin)) : KEY_NEVER;
                w_cnt[lane] = 0;
                w_flag[lane] = 0;
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int row = a.row_begin + rb * BM + r0w + hh * 8 + g;
                const bool valid = row < a.row_end;
                my_cs[hh] = valid ? a.row_cs[row] : 0;
                my_ce[hh] = valid ? a.row_ce[row] : 0;
                my_n[hh] = valid ? a.n32[row] : INFINITY;
            }
            __syncwarp();
        }
        const int t = tl ? tl[q] : (q < skip_lo ? q : q + skip_n);
        const int col0 = t * BN;
        q += qs;
        // the tile's column norms (and, SYM, the column bins' thresholds): L2 -> this warp's shared copies
        __syncwarp();
        cp_async_16(w_cn + 4 * lane, a.n32 + col0 + 4 * lane);
        if (SYM) {
            cp_async_16(w_ct + 2 * lane, a.col_thr + col0 + 2 * lane);
            cp_async_16(w_ct + 64 + 2 * lane, a.col_thr + col0 + 64 + 2 * lane);
        }
        cp_async_commit();
        u64 shared_thr = ~0ull;
        if (lane < WROWS) {
            const int row = a.row_begin + rb * BM + r0w + lane;
            if (row < a.row_end) shared_thr = __ldcg(a.row_thr + (row - a.row_begin));
        }
        ++tcount;

        float acc[16][4];
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;

        for (int kc = 0; kc < a.nkc; ++kc) {
            if (!ready) {
                const long long pf_w0 = clock64();
                mbar_wait(&sm.full[stage], phase);
                pf_wait += clock64() - pf_w0;
            }
            const uint32_t base = tiles_u32 + (uint32_t)stage * STAGE_BYTES;
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == STAGES) { nstage = 0; nphase ^= 1u; }
            const bool ready_next = mbar_test_wait(&sm.full[nstage], nphase);
#pragma unroll
            for (int ks = 0; ks < BKH / 16; ++ks) {
                uint32_t fa0, fa1, fa2, fa3;
                ldsm_x4(base + a_off + ((((uint32_t)(2 * ks) + a_kc) ^ xr) << 4), fa0, fa1, fa2, fa3);
                const uint32_t bsw = (((uint32_t)(2 * ks) + b_kc) ^ xr) << 4;
#pragma unroll
                for (int np = 0; np < 8; ++np) {
                    uint32_t fb0, fb1, fb2, fb3;
                    ldsm_x4(base + b_off + (uint32_t)np * 2048u + bsw, fb0, fb1, fb2, fb3);
                    hmma_16816(acc[2 * np], fa0, fa1, fa2, fa3, fb0, fb1);
                    hmma_16816(acc[2 * np + 1], fa0, fa1, fa2, fa3, fb2, fb3);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
            stage = nstage;
            phase = nphase;
            ready = ready_next;
        }

        // ---- epilogue: d~ = (n_i + n_j) - 2 s in fp32, one compare per entry against the row's (and, SYM, the column's) bound
        const long long pf_e0 = clock64();
        if (lane < WROWS) {
            if (shared_thr < w_thr[lane]) w_thr[lane] = shared_thr;
            w_rt[lane] = tau32_of_key(w_thr[lane]);
        }
        cp_async_wait<0>();
        __syncwarp();
        if (SYM) {
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) w_ctf[lane + 32 * i] = tau32_of_key(w_ct[lane + 32 * i]);
            __syncwarp();
        }
        u64* ck = a.cand_key + (size_t)seg * seg_stride;
        int* cj = a.cand_j + (size_t)seg * seg_stride;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int rw = hh * 8 + g;
            const float taui = w_rt[rw];
            const float ni = my_n[hh];
            unsigned mask = 0, cmask = 0;
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                const float2 nj = *reinterpret_cast<const float2*>(w_cn + nt * 8 + 2 * q4);
                const float d0 = fmaf(-2.0f, acc[nt][2 * hh], ni + nj.x);
                const float d1 = fmaf(-2.0f, acc[nt][2 * hh + 1], ni + nj.y);
                if (d0 <= taui) mask |= 1u << (nt * 2);
                if (d1 <= taui) mask |= 1u << (nt * 2 + 1);
                if (SYM) {
                    const float2 tj = *reinterpret_cast<const float2*>(w_ctf + nt * 8 + 2 * q4);
                    if (d0 <= tj.x) cmask |= 1u << (nt * 2);
                    if (d1 <= tj.y) cmask |= 1u << (nt * 2 + 1);
                }
            }
            if (mask | cmask) {
                // rare path: park this row's 32 distances in the warp's scratch (lane-interleaved) and walk the set bits
                float* tmp = reinterpret_cast<float*>(w_sk) + lane;        // entry b lives at tmp[b * 32]
#pragma unroll
                for (int nt = 0; nt < 16; ++nt) {
                    const float2 nj = *reinterpret_cast<const float2*>(w_cn + nt * 8 + 2 * q4);
                    tmp[(nt * 2) * 32] = fmaf(-2.0f, acc[nt][2 * hh], ni + nj.x);
                    tmp[(nt * 2 + 1) * 32] = fmaf(-2.0f, acc[nt][2 * hh + 1], ni + nj.y);
                }
                const int cs = my_cs[hh];
                const unsigned clen = (unsigned)(my_ce[hh] - cs);
                unsigned m2 = mask;
                while (m2) {                               // drop non-finite distances and the row's own chromosome
                    const int bit = __ffs(m2) - 1;
                    m2 &= m2 - 1;
                    const int cl = (bit >> 1) * 8 + 2 * q4 + (bit & 1);
                    if (!(fabsf(tmp[bit * 32]) < INFINITY) || (unsigned)(col0 + cl - cs) < clen) mask &= ~(1u << bit);
                }
                if (mask) {
                    pf_emit += __popc(mask);
                    int w = atomicAdd(&w_cnt[rw], __popc(mask));
                    u64* rk = ck + (size_t)(r0w + rw) * a.cap;
                    int* rj = cj + (size_t)(r0w + rw) * a.cap;
                    while (mask) {
                        const int bit = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const int cl = (bit >> 1) * 8 + 2 * q4 + (bit & 1);
                        if (w < a.cap) {
                            rk[w] = (u64)__double_as_longlong(-0.5 * (double)tmp[bit * 32]);
                            rj[w] = col0 + cl;
                        } else {
                            w_flag[rw] = 1;
                        }
                        ++w;
                    }
                }
                if (SYM && cmask) {
                    const int i = a.row_begin + rb * BM + r0w + rw;
                    while (cmask) {
                        const int bit = __ffs(cmask) - 1;
                        cmask &= cmask - 1;
                        const int j = col0 + (bit >> 1) * 8 + 2 * q4 + (bit & 1);
                        const float dv = tmp[bit * 32];
                        if (!(fabsf(dv) < INFINITY) || j >= a.N || i >= a.row_end) continue;
                        if ((unsigned)(j - cs) < clen) continue;
                        const u64 key = (u64)__double_as_longlong(-0.5 * (double)dv);
                        const int pos = atomicAdd(w_stgc, 1);
                        if (pos < STG) {
                            w_stg[pos] = make_uint4((unsigned)key, (unsigned)(key >> 32), (unsigned)j, (unsigned)i);
                        } else {
                            const int w = atomicAdd(a.in_cnt + j, 1);
                            if (w < a.in_cap) {
                                a.in_key[(size_t)j * a.in_cap + w] = key;
                                a.in_j[(size_t)j * a.in_cap + w] = i;
                            }
                        }
                        ++pf_emit;
                    }
                }
            }
        }
        __syncwarp();
        if (SYM && *w_stgc >= 32) flush_incoming();
        const long long pf_p0 = clock64();
        pf_epi += pf_p0 - pf_e0;
        prune_rows(__ballot_sync(0xffffffffu, lane < WROWS && w_cnt[lane] > a.cap - BN && !w_flag[lane]), ck, cj);
        pf_prune += clock64() - pf_p0;
    }
    __syncwarp();
    if (a.prof != nullptr && warp == 0 && lane == 0) {
        long long* o = a.prof + (size_t)blockIdx.x * 8;
        o[0] = clock64() - pf_t0; o[1] = pf_wait; o[2] = pf_epi; o[3] = pf_prune;
        o[4] = tcount; o[5] = pf_nprune; o[6] = pf_emit; o[7] = 0;
    }
}

