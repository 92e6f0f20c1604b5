// wisecondor_b200 - K5t: the reference-bin search's filter on the 5th-generation tensor cores (tcgen05 + TMEM).
// Textually included by wc_search.cu inside its anonymous namespace (uses TopkArgs, prune_row, the PTX helpers and the
// K4h kernel of wc_search_f16.cuh).
#pragma once

// ---------------------------------------------------------------------------------------------------------
// K5t  wc_dist_topk_tc_kernel<SYM, DBG>
// ---------------------------------------------------------------------------------------------------------
// Same job as K5 / K5h (getRefForBins' distance row, /root/reference/wisetools.py:302, as a FILTER: K6 re-scores the
// survivors exactly in fp64): for a block of 128 target bins and one or two blocks of 128 candidate bins, the fp16 dot
// products s_ij = x'_i . x'_j with fp32 accumulation, d~ = (n_i + n_j) - 2 s, and two compares per entry - against the
// row bin's threshold and (SYM) against the column bin's.
//
// Blackwell-native structure (one persistent CTA per SM, 12 warps, no thread ever holds an MMA fragment):
//   warp 0, one lane   TMA producer: per 64-sample chunk three SWIZZLE_128B boxes (A: 128 bins, B: 2 x 128 bins, 16 KiB
//                      each) into a 3-stage mbarrier ring (SASS UTMALDG)
//   warp 1, one lane   MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 256 (or 128 for an odd last block),
//                      K = 16, operands straight from the swizzled shared-memory tiles through matrix descriptors,
//                      accumulators in TENSOR MEMORY: two buffers of 256 fp32 columns x 128 lanes (all 512 columns);
//                      tcgen05.commit releases ring stages and publishes finished accumulators (SASS UTCHMMA / UTCBAR)
//   warp 2             allocates / frees the tensor memory
//   warps 4-11         epilogue, two warpgroups: warp 4+e / 8+e own TMEM lanes [32e, 32e+32) = tile rows, ONE THREAD PER
//                      TARGET BIN, and the first / second candidate block of the item (two warps per scheduler hide each
//                      other's latencies; each half has its own candidate segment, so they share no row state).  The row's
//                      threshold, norm and exclusion range live in the thread's registers; the accumulator row arrives with
//                      tcgen05.ld.32x32b.x32 (SASS LDTM), 32 columns at a time, while the tensor core already works on the
//                      next tile in the other TMEM buffer.  Registers move from warpgroup 0 to the epilogue (setmaxnreg).
// Per 128 x 256 tile the tensor core needs 40 x 128 = 5120 cycles, the epilogue ~2000 issue slots per warp: it hides.
// What bounds the kernel is the L2 -> shared-memory operand stream (48 KiB per 64-sample chunk of a 128 x 256 tile).
constexpr int TC_THREADS = 384;                     // warpgroup 0: TMA / MMA / TMEM roles; warpgroups 1, 2: epilogue
constexpr int TC_STAGES = 3;                        // 144 KiB in flight: the L2 stream, not the latency, bounds the ring
constexpr int TC_STAGE_BYTES = 3 * TILE_BYTES;      // A, B0, B1 boxes: 128 rows x 64 halves x 2 B = 16 KiB each
constexpr int TC_NB = 2 * BN;                       // columns of a full item (two candidate blocks)
constexpr int TC_STG = 256;                         // per-warp staging entries (8 bytes each) for column-side candidates
constexpr int TC_PIVOTS = 512;                      // pivots of the pivot pass: their distances to 128 bins fill the tensor memory
constexpr int TC_EPI_WARPS = 8;                     // two per TMEM lane quarter: columns [0, 128) and [128, 256) of an item
constexpr int TC_SEGS_PER_PIECE = 2;                // each column half keeps its own candidate buffers (no shared row state)
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr int TC_PARK_LD = 36;                      // words between two rows of a parked chunk (16-byte aligned, conflict-free 128-bit stores)

struct __align__(16) TcState {
    uint64_t full[TC_STAGES];
    uint64_t empty[TC_STAGES];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
};

constexpr size_t TC_SMEM_BYTES = (size_t)TC_STAGES * TC_STAGE_BYTES                 // operand ring
                                 + 2 * 2 * TC_NB * sizeof(float)                     // column norms + thresholds, 2 buffers
                                 + (size_t)TC_EPI_WARPS * TC_STG * sizeof(uint2)     // column-side staging
                                 + (size_t)TC_EPI_WARPS * 32 * TC_PARK_LD * sizeof(float)    // per-warp parking of one 32 x 32 chunk
                                 + sizeof(TcState);

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands (K-major, SWIZZLE_128B), fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor of a K-major SWIZZLE_128B tile (rows of 128 bytes, 8-row groups 1024 bytes apart):
// start address >> 4 | leading byte offset (unused for swizzled K-major: 1) << 16 | stride byte offset (1024 >> 4) << 32 |
// descriptor version 1 << 46 | layout type SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D fp32 (1 << 4), A and B fp16 (0), both K-major (0), N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t tc_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// bits of the double -d/2 for a float distance d, integer-only (|d| below the smallest normal float counts as 0): the score
// key the candidate buffers, the prune and K6 work on (wc_search.cu: key_of_tau)
__device__ __forceinline__ u64 tc_key_of_f32(float d) {
    const unsigned u = __float_as_uint(d);
    const unsigned ex = (u >> 23) & 0xffu;
    const u64 sign = (u64)(~u & 0x80000000u) << 32;
    return ex == 0 ? sign : (sign | ((u64)(ex + 895u) << 52) | ((u64)(u & 0x7fffffu) << 29));
}
// the fp32 bound "d~ <= tau" of a threshold key (bits of the double -tau/2), rounded UP, integer-only (no FP64 pipe in the
// epilogue): exponent re-biased by +1 (x 2) and -896 (double -> float), the 29 dropped mantissa bits round away from zero
__device__ __forceinline__ float tc_tau32_of_key(u64 key) {
    if (key == 0ull) return -0.0f;                               // KEY_NEVER
    const u64 mag = key & 0x7fffffffffffffffull;
    const unsigned ex = (unsigned)(mag >> 52);
    if (ex < 897u) return 1.1754944e-38f;                        // below the float normals: the smallest normal bounds it
    if (ex >= 1149u) return 3.4028235e38f;                       // beyond the float range (tau_init and the like)
    const unsigned f = (unsigned)(mag >> 29) - (895u << 23) + ((mag & 0x1fffffffull) != 0ull ? 1u : 0u);
    return __uint_as_float(f);      // tau > 0 always (the key's sign bit is set); a mantissa carry moves into the exponent
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: plain (rows only), 1: symmetric (rows + column side), 2: pivot pass (thresholds from TMEM-resident distances)
template <int MODE, bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1)
wc_dist_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b, const TopkArgs a) {
    constexpr bool SYM = MODE == 1;
    constexpr bool PIV = MODE == 2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();   // SWIZZLE_128B tiles need a 1024-byte aligned base
    unsigned char* tiles = smem_raw;            // (no pointer arithmetic through integers: every access below stays an LDS / STS)
    float* s_nj = reinterpret_cast<float*>(smem_raw + (size_t)TC_STAGES * TC_STAGE_BYTES);    // [2][TC_NB]
    float* s_tj = s_nj + 2 * TC_NB;                                                            // [2][TC_NB]  (pivot pass: the pivots' bins, as int)
    uint2* s_stg = reinterpret_cast<uint2*>(s_tj + 2 * TC_NB);                                 // [8][TC_STG]
    float* s_park = reinterpret_cast<float*>(s_stg + TC_EPI_WARPS * TC_STG);                   // [8][32][TC_PARK_LD]
    TcState& sm = *reinterpret_cast<TcState*>(s_park + TC_EPI_WARPS * 32 * TC_PARK_LD);
    const int tid = threadIdx.x;
    const int warp_all = tid >> 5, lane = tid & 31;

    const int pb = a.cta_piece_begin[blockIdx.x], pe = a.cta_piece_begin[blockIdx.x + 1];
    if (pb >= pe) return;                       // CTA-uniform: nothing allocated yet

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&sm.tfull[b], 1);
            mbar_init(&sm.tempty[b], PIV ? 4 : TC_EPI_WARPS);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp_all == 2) {                        // one warp allocates all 512 TMEM columns (one CTA per SM: no contention)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (PIV) {                                  // the pivots' norms and bins: the same 512 columns for every row block
        for (int i = tid; i < TC_PIVOTS; i += TC_THREADS) {
            s_nj[i] = a.coln32[i];
            reinterpret_cast<int*>(s_tj)[i] = a.col_ids[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    // every role walks the same sequence of items: (row block, one or two column blocks) per piece
    auto tile_of = [&](const int* tl, int q, int skip_lo, int skip_n) {
        return tl ? tl[q] : (PIV || q < skip_lo ? q : q + skip_n);      // pivot pass: tiles of the pivot matrix
    };

    if (warp_all < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");       // warpgroup 0 hands its registers to the epilogue warpgroups
    if (warp_all == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pi = pb; pi < pe; ++pi) {
                const int* pc = a.pieces + (size_t)pi * 5;
                const int rbp = pc[0], q1 = pc[2], qs = pc[3];
                const int skip_lo = PIV ? 0 : a.rb_skip_lo[rbp], skip_n = PIV ? 0 : a.rb_skip_n[rbp];
                const int* tl = a.tile_list ? a.tile_list + a.rb_list_off[rbp] : nullptr;
                const int row0 = a.row_begin + rbp * BM;
                for (int q = pc[1]; q < q1;) {
                    const int c0 = tile_of(tl, q, skip_lo, skip_n) * BN;
                    q += qs;
                    const bool two = q < q1;
                    const int c1 = two ? tile_of(tl, q, skip_lo, skip_n) * BN : 0;
                    if (two) q += qs;
                    for (int kc = 0; kc < a.nkc; ++kc) {
                        while (!mbar_try_wait(&sm.empty[stage], phase ^ 1u)) __nanosleep(64);
                        unsigned char* st = tiles + (size_t)stage * TC_STAGE_BYTES;
                        mbar_arrive_expect_tx(&sm.full[stage], (two ? 3 : 2) * TILE_BYTES);
                        const int kb = kc == a.nkc - 1 ? a.kb_last : kc * BKH;      // folded norms: the B version of the last chunk
                        tma_load_2d(st, &tmap, kc * BKH, row0, &sm.full[stage]);
                        tma_load_2d(st + TILE_BYTES, &tmap_b, kb, c0, &sm.full[stage]);
                        if (two) tma_load_2d(st + 2 * TILE_BYTES, &tmap_b, kb, c1, &sm.full[stage]);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp_all == 1) {
        // ===== MMA issuer: one thread feeds the tensor core =====
        if (lane == 0) {
            const uint32_t idesc2 = tc_idesc(TC_NB), idesc1 = tc_idesc(BN);
            const uint32_t tiles_u32 = smem_u32(tiles);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int pi = pb; pi < pe; ++pi) {
                const int* pc = a.pieces + (size_t)pi * 5;
                const int q1 = pc[2], qs = pc[3];
                for (int q = pc[1]; q < q1; ++it) {
                    q += qs;
                    const bool two = q < q1;
                    if (two) q += qs;
                    const int buf = it & 1;
                    while (!mbar_try_wait(&sm.tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1))) __nanosleep(32);   // the epilogue has drained this buffer
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + (uint32_t)buf * TC_NB;
                    const uint32_t idesc = two ? idesc2 : idesc1;
                    for (int kc = 0; kc < a.nkc; ++kc) {
                        while (!mbar_try_wait(&sm.full[stage], phase)) __nanosleep(20);
                        tc_fence_after();
                        const uint32_t sbase = tiles_u32 + (uint32_t)stage * TC_STAGE_BYTES;
                        const uint64_t adesc = tc_smem_desc(sbase);
                        const uint64_t bdesc = tc_smem_desc(sbase + TILE_BYTES);
#pragma unroll
                        for (int ks = 0; ks < BKH / 16; ++ks)        // 16 halves = 32 bytes along K inside the swizzle atom
                            tc_mma_f16(d_addr, adesc + 2 * ks, bdesc + 2 * ks, idesc, (kc | ks) != 0 ? 1u : 0u);
                        tc_commit(&sm.empty[stage]);                 // the stage is free once these MMAs have read it
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit(&sm.tfull[buf]);                       // accumulator complete -> epilogue
                }
            }
        }
    }
    } else if (PIV) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
      if (warp_all < 8) {
        // ===== pivot pass epilogue: one thread per target bin, its 512 pivot distances stay in tensor memory =====
        // The thread bisects on the value for a cut with k <= #{valid pivots with d~ <= cut} <= k + 24, re-reading its TMEM lane
        // (16 x tcgen05.ld.32x32b.x32) per round: no candidate buffer, no prune, no global traffic but the final threshold.
        const int e = warp_all - 4;
        const int r = e * 32 + lane;
        const int* s_pid = reinterpret_cast<const int*>(s_tj);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(e * 32) << 16);
        int it = 0;
        for (int pi = pb; pi < pe; ++pi, it += 2) {
            const int rb = a.pieces[(size_t)pi * 5];
            const int row = a.row_begin + rb * BM + r;
            const bool valid = row < a.row_end;
            const float ni = valid ? a.n32[row] : INFINITY;
            const int cs = valid ? a.row_cs[row] : 0, ce = valid ? a.row_ce[row] : 0;
            // pivots are sorted by bin: those of the row's own chromosome are the index range [plo, phi)
            int plo = 0, phi = 0;
            {
                int lo = 0, hi = TC_PIVOTS;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (s_pid[m] < cs) lo = m + 1; else hi = m; }
                plo = lo;
                hi = TC_PIVOTS;
                while (lo < hi) { const int m = (lo + hi) >> 1; if (s_pid[m] < ce) lo = m + 1; else hi = m; }
                phi = lo;
            }
            mbar_wait(&sm.tfull[0], (uint32_t)((it >> 1) & 1));
            mbar_wait(&sm.tfull[1], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            // one sweep over the lane: #{valid pivots with d~ <= cut}; FIRST also tracks min / max of the valid distances
            float vmin = INFINITY, vmax = -INFINITY;
            auto sweep = [&](float cut, bool first) -> int {
                int c = 0;
                for (int ch = 0; ch < TC_PIVOTS / 32; ++ch) {
                    uint32_t v[32];
                    tc_ld32(lane_addr + (uint32_t)(ch * 32), v);
                    tc_wait_ld();
                    int lo = plo - ch * 32, hi = phi - ch * 32;
                    lo = lo < 0 ? 0 : lo;
                    hi = hi > 32 ? 32 : hi;
                    const unsigned excl = lo < hi ? ((hi == 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
                    unsigned m = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        // the accumulator holds -d~ / 2 (norms folded into the contraction, wc_prepare_f16_kernel)
                        const float d0 = -2.0f * __uint_as_float(v[4 * u + 0]);
                        const float d1 = -2.0f * __uint_as_float(v[4 * u + 1]);
                        const float d2 = -2.0f * __uint_as_float(v[4 * u + 2]);
                        const float d3 = -2.0f * __uint_as_float(v[4 * u + 3]);
                        if (d0 <= cut) m |= 1u << (4 * u + 0);
                        if (d1 <= cut) m |= 1u << (4 * u + 1);
                        if (d2 <= cut) m |= 1u << (4 * u + 2);
                        if (d3 <= cut) m |= 1u << (4 * u + 3);
                        if (first) {
                            const unsigned ok = ~excl >> (4 * u);
                            if ((ok & 1u) && fabsf(d0) < INFINITY) { vmin = fminf(vmin, d0); vmax = fmaxf(vmax, d0); }
                            if ((ok & 2u) && fabsf(d1) < INFINITY) { vmin = fminf(vmin, d1); vmax = fmaxf(vmax, d1); }
                            if ((ok & 4u) && fabsf(d2) < INFINITY) { vmin = fminf(vmin, d2); vmax = fmaxf(vmax, d2); }
                            if ((ok & 8u) && fabsf(d3) < INFINITY) { vmin = fminf(vmin, d3); vmax = fmaxf(vmax, d3); }
                        }
                    }
                    c += __popc(m & ~excl);
                }
                return c;
            };
            const int total = sweep(3.0e38f, true);              // finite valid pivot distances
            float lo = vmin, hi = vmax;                           // count(<= hi) >= k throughout (if total >= k)
            int c_hi = total;
            for (int round = 0; round < 14; ++round) {
                // the loop is warp-collective (tcgen05.ld is): lanes that are done keep sweeping with their final cut
                const bool need = total >= a.k && c_hi > a.k + 24 && hi > lo;
                if (!__any_sync(0xffffffffu, need)) break;
                const float mid = lo + 0.5f * (hi - lo);
                const int c = sweep(need ? mid : hi, false);
                if (need) {
                    if (c >= a.k) { hi = mid; c_hi = c; } else { lo = mid; }
                }
            }
            if (valid && total >= a.k) {
                const double dv = (double)hi;
                double tau = dv + a.mcoef * ((double)ni + fabs(dv)) + madd_of(a);
                if (!(tau > 1e-300)) tau = 1e-300;
                atomicMin(a.row_thr + (row - a.row_begin), key_of_tau(tau));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&sm.tempty[0]); mbar_arrive(&sm.tempty[1]); }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ===== epilogue: one thread per target bin =====
        const int e = warp_all & 3;                 // TMEM lane quarter of this warp (== warp_all % 4)
        const int g = (warp_all - 4) >> 2;          // column half of the item this warp handles (candidate block 0 / 1)
        const int we = warp_all - 4;                // epilogue warp index 0..7 (per-warp shared-memory areas)
        const int r = e * 32 + lane;                // row of the tile = TMEM lane
        uint2* w_stg = s_stg + we * TC_STG;         // (accumulator bits, bin j) of column-side candidates
        float* w_park = s_park + we * (32 * TC_PARK_LD);   // the chunk's 32 x 32 accumulator values: entry (lane l, column b) at [l * TC_PARK_LD + b]
        int cnt = 0;                                // candidates of this thread's row in ITS segment: the row has one writer
        bool flag = false;                          // the row's buffer overflowed (tie plateau): exact fallback
        const size_t seg_stride = (size_t)BM * a.cap;
        long long pf_epi = 0, pf_prune = 0, pf_nprune = 0, pf_emit = 0, pf_wait = 0, pf_col = 0, pf_flush = 0, pf_surv = 0;
        const long long pf_t0 = clock64();

        // Append the staged column-side candidates to their bins' incoming buffers: four global atomics in flight per lane
        // before the first result is needed.  rowbase = first bin of this warp's 32 rows.
        int stg_n = 0;                              // staged entries: warp-uniform, lives in a register
        auto flush_incoming = [&](int rowbase) {
            const long long pf_f0 = clock64();
            __syncwarp();
            for (int base = 0; base < stg_n; base += 128) {
                uint2 sv[4];
                int w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = base + u * 32 + lane;
                    w[u] = -1;
                    if (x < stg_n) {
                        sv[u] = w_stg[x];
                        w[u] = atomicAdd(a.in_cnt + (int)(sv[u].y & 0x7ffffffu), 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (w[u] >= 0 && w[u] < a.in_cap) {
                        const size_t o = (size_t)(sv[u].y & 0x7ffffffu) * a.in_cap + w[u];
                        a.in_key[o] = tc_key_of_f32(__uint_as_float(sv[u].x));
                        a.in_j[o] = rowbase + (int)(sv[u].y >> 27);
                    }
                }
            }
            __syncwarp();
            stg_n = 0;
            pf_flush += clock64() - pf_f0;
        };

        int it = 0;
        for (int pi = pb; pi < pe; ++pi) {
            const int* pc = a.pieces + (size_t)pi * 5;
            const int rb = pc[0], q1 = pc[2], qs = pc[3], seg = pc[4] + g;      // each column half has its own segment
            const int skip_lo = a.rb_skip_lo[rb], skip_n = a.rb_skip_n[rb];
            const int* tl = a.tile_list ? a.tile_list + a.rb_list_off[rb] : nullptr;
            // this thread's bin
            const int rowbase = a.row_begin + rb * BM + e * 32;
            const int row = rowbase + lane;
            const bool valid = row < a.row_end;
            const double nrm = valid ? a.norms[row] : 0.0;
            const int cs = valid ? a.row_cs[row] : 0;
            const unsigned clen = valid ? (unsigned)(a.row_ce[row] - cs) : 0u;
            u64 thr = valid ? __ldcg(a.row_thr + (row - a.row_begin)) : KEY_NEVER;
            u64* wk = a.cand_key + (size_t)seg * seg_stride + (size_t)(e * 32) * a.cap;       // buffers of the warp's 32 rows
            int* wj = a.cand_j + (size_t)seg * seg_stride + (size_t)(e * 32) * a.cap;
            cnt = 0;
            flag = false;

            // warp-collective prune of the rows named in `need`; the owning lane adopts the result
            auto prune_rows = [&](unsigned need) {
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    int n = __shfl_sync(0xffffffffu, cnt, src);
                    if (n > a.cap) n = a.cap;
                    const double nr = __shfl_sync(0xffffffffu, nrm, src);
                    u64 nthr;
                    int kept;
                    ++pf_nprune;
                    __threadfence_block();
                    __syncwarp();
                    if (a.cap <= 512)
                        prune_row<16>(wk + (size_t)src * a.cap, wj + (size_t)src * a.cap, n, a.k, nr, a.mcoef, madd_of(a), lane, nullptr, nullptr, &nthr, &kept);
                    else
                        prune_row<32>(wk + (size_t)src * a.cap, wj + (size_t)src * a.cap, n, a.k, nr, a.mcoef, madd_of(a), lane, nullptr, nullptr, &nthr, &kept);
                    if (lane == src) {
                        if (kept > a.cap - BN) {             // a tie plateau wider than the buffer: exact fallback
                            flag = true;
                            thr = KEY_NEVER;
                            cnt = 0;
                        } else {
                            const u64 other = atomicMin(a.row_thr + (row - a.row_begin), nthr);   // publish; adopt a tighter one
                            thr = other < nthr ? other : nthr;
                            cnt = kept;
                        }
                    }
                    __syncwarp();
                }
            };

            for (int q = pc[1]; q < q1; ++it) {
                const int c0 = tile_of(tl, q, skip_lo, skip_n) * BN;
                q += qs;
                const bool two = q < q1;
                const int c1 = two ? tile_of(tl, q, skip_lo, skip_n) * BN : 0;
                if (two) q += qs;
                const int buf = it & 1;
                const int cg = g == 0 ? c0 : c1;      // first bin of this warp's candidate block
                const bool mine = g == 0 || two;      // an odd last item has no second block: the second warpgroup only keeps step
                // the block's column thresholds (SYM) and the row's shared threshold: issued before the wait.  (Fetching them one
                // item ahead was measured slower, r03r: staler thresholds let more survivors through than the hidden latency saves.)
                u64 ctr = KEY_NEVER;
                if (SYM && mine) ctr = __ldcg(a.col_thr + cg + r);
                if (valid) {
                    const u64 shared_thr = __ldcg(a.row_thr + (row - a.row_begin));
                    if (shared_thr < thr) thr = shared_thr;
                }
                float* tj_t = s_tj + buf * TC_NB + g * BN;           // this warpgroup's half of the table: tau_j / 2
                if (SYM) tj_t[r] = 0.5f * tc_tau32_of_key(ctr);
                // The accumulator holds u = -d~ / 2 (norms folded into the contraction): "d~ <= tau" is "u + tau / 2 >= 0", the
                // SIGN BIT of one add - no compare, no select.  -inf: nothing of an invalid row passes.
                const float hti = valid ? 0.5f * tc_tau32_of_key(thr) : -INFINITY;
                if (SYM) named_bar_sync(1 + g, 128);                // table visible to the four warps of the warpgroup
                const long long pf_w0 = clock64();
                mbar_wait(&sm.tfull[buf], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                const long long pf_e0 = clock64();
                pf_wait += pf_e0 - pf_w0;

                const int nchunk = mine ? BN / 32 : 0;
                if (!mine) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.tempty[buf]);
                }
                for (int ch = 0; ch < nchunk; ++ch) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(buf * TC_NB + g * BN + ch * 32), v);
                    tc_wait_ld();
                    if (ch == nchunk - 1) {          // the accumulator row is in registers: hand the buffer back to the tensor core
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.tempty[buf]);
                    }
                    const int cb = ch * 32;          // column within the warp's block (0..127)
                    const int colbase = cg + cb;     // global bin of the chunk's first column
                    unsigned mask = 0, cmask = 0;
                    {   // Sign bits gathered by funnel shifts, most significant entry first; four independent chains per side.
                        unsigned pm[4] = {0u, 0u, 0u, 0u}, qm[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                        for (int u = 7; u >= 0; --u) {
                            float4 tj = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (SYM) tj = *reinterpret_cast<const float4*>(tj_t + cb + 4 * u);
                            const float tjv[4] = {tj.x, tj.y, tj.z, tj.w};
#pragma unroll
                            for (int w = 3; w >= 0; --w) {
                                const float uv = __uint_as_float(v[4 * u + w]);
                                pm[u >> 1] = __funnelshift_l(__float_as_uint(uv + hti), pm[u >> 1], 1);
                                if (SYM) qm[u >> 1] = __funnelshift_l(__float_as_uint(uv + tjv[w]), qm[u >> 1], 1);
                            }
                        }
                        mask = ~((pm[0] | (pm[1] << 8)) | ((pm[2] << 16) | (pm[3] << 24)));     // a set sign bit = beyond the threshold
                        if (SYM) cmask = ~((qm[0] | (qm[1] << 8)) | ((qm[2] << 16) | (qm[3] << 24)));
                    }
                    if (DBG) {
                        if (valid && a.dbg != nullptr) {
#pragma unroll
                            for (int b = 0; b < 32; ++b)
                                if (colbase + b < a.dbg_ld) a.dbg[(size_t)(row - a.row_begin) * a.dbg_ld + colbase + b] = -2.0f * __uint_as_float(v[b]);
                        }
                    }
                    // Columns beyond the last bin (zero padding: u = 0 would pass) and, on the column side, rows beyond the last
                    // target bin never count; nor do the columns of the row's own chromosome (a symmetric relation: both sides).
                    {
                        const int ncol = a.N - colbase;
                        const unsigned vm = ncol >= 32 ? 0xffffffffu : (ncol <= 0 ? 0u : (1u << ncol) - 1u);
                        mask &= vm;
                        cmask = valid ? (cmask & vm) : 0u;
                        int lo = cs - colbase, hi = lo + (int)clen;
                        lo = lo < 0 ? 0 : lo;
                        hi = hi > 32 ? 32 : hi;
                        if (lo < hi) {
                            const unsigned excl = (hi == 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u);
                            mask &= ~excl;
                            cmask &= ~excl;
                        }
                    }
                    // ---- survivors (~0.7 per thread and chunk on each side once thresholds are tight) ----
                    // Row side: a row has ONE writer, this thread - its count lives in a register, entries go straight to the row's
                    // buffer.  Column side: slots of the warp's staging area by a warp scan.  No atomics, no list indirection, no
                    // warp barrier on the way (r02: lists + shared-memory atomics + four barriers per chunk cost 2 x the MMA time).
                    const int rn = __popc(mask), cn = SYM ? __popc(cmask) : 0;
                    const long long pf_s0 = clock64();
                    if (__any_sync(0xffffffffu, (mask | cmask) != 0u)) {
                        // v[] is indexed by a run-time bit below: parked in this thread's row of the warp's patch (rows 36 words apart:
                        // the eight 128-bit stores are conflict-free)
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            *reinterpret_cast<uint4*>(w_park + lane * TC_PARK_LD + 4 * u) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                        while (mask) {                  // two survivors per round: their value -> key -> store chains overlap
                            const int bit = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int bit2 = mask ? __ffs(mask) - 1 : -1;
                            mask &= mask - 1;           // (0 stays 0)
                            const float u1 = w_park[lane * TC_PARK_LD + bit];
                            const float u2 = w_park[lane * TC_PARK_LD + (bit2 < 0 ? bit : bit2)];
                            const u64 k1 = tc_key_of_f32(-2.0f * u1), k2 = tc_key_of_f32(-2.0f * u2);
                            if (cnt < a.cap) {
                                wk[(size_t)lane * a.cap + cnt] = k1;
                                wj[(size_t)lane * a.cap + cnt] = colbase + bit;
                            } else {
                                flag = true;
                            }
                            ++cnt;
                            if (bit2 >= 0) {
                                if (cnt < a.cap) {
                                    wk[(size_t)lane * a.cap + cnt] = k2;
                                    wj[(size_t)lane * a.cap + cnt] = colbase + bit2;
                                } else {
                                    flag = true;
                                }
                                ++cnt;
                            }
                        }
                        const long long pf_c0 = clock64();
                        if (SYM) {
                            int incl = cn;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += t;
                            }
                            const int ct = __shfl_sync(0xffffffffu, incl, 31);
                            if (ct > TC_STG) {
                                // more than the staging area holds (loose thresholds at the start of a pass): straight to the bins
                                while (cmask) {
                                    const int bit = __ffs(cmask) - 1;
                                    cmask &= cmask - 1;
                                    const int j = colbase + bit;
                                    const int w = atomicAdd(a.in_cnt + j, 1);
                                    if (w < a.in_cap) {
                                        a.in_key[(size_t)j * a.in_cap + w] = tc_key_of_f32(-2.0f * w_park[lane * TC_PARK_LD + bit]);
                                        a.in_j[(size_t)j * a.in_cap + w] = row;
                                    }
                                }
                            } else if (ct != 0) {
                                if (stg_n + ct > TC_STG) flush_incoming(rowbase);
                                int q = stg_n + incl - cn;
                                while (cmask) {
                                    const int bit = __ffs(cmask) - 1;
                                    cmask &= cmask - 1;
                                    w_stg[q++] = make_uint2(__float_as_uint(-2.0f * w_park[lane * TC_PARK_LD + bit]), (unsigned)(colbase + bit) | ((unsigned)lane << 27));
                                }
                                stg_n += ct;
                            }
                        }
                        pf_emit += rn + cn;
                        pf_col += clock64() - pf_c0;
                    }
                    pf_surv += clock64() - pf_s0;
                }
                const long long pf_p0 = clock64();
                pf_epi += pf_p0 - pf_e0;
                // ---- prune rows whose buffer could overflow during the next item ----
                prune_rows(__ballot_sync(0xffffffffu, cnt > a.cap - BN && !flag));
                pf_prune += clock64() - pf_p0;
            }
            // piece finished
            __syncwarp();
            if (SYM) flush_incoming(rowbase);
            if (a.final_prune)      // every row leaves its best threshold behind (the column side of later tiles is filtered by it)
                prune_rows(__ballot_sync(0xffffffffu, cnt > a.k + 24 && cnt <= a.cap && !flag));
            __syncwarp();
            a.seg_cnt[(size_t)seg * BM + r] = cnt > a.cap ? a.cap : cnt;
            a.seg_flag[(size_t)seg * BM + r] = flag ? 1 : 0;
        }
        if (a.prof != nullptr && we == 0 && lane == 0) {
            long long* o = a.prof + (size_t)blockIdx.x * 8;
            o[0] = clock64() - pf_t0; o[1] = pf_wait; o[2] = pf_epi; o[3] = pf_prune;
            o[4] = it; o[5] = pf_col; o[6] = pf_surv; o[7] = pf_flush;
        }
    }

    // ===== teardown: every tcgen05 operation of this CTA has completed (the epilogue consumed the last accumulator) =====
    tc_fence_before();
    __syncthreads();
    if (warp_all == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// Pivot pass: tight thresholds before the first tile of the search proper
// ---------------------------------------------------------------------------------------------------------
// d(i, j) = n_i + n_j - 2 x'_i . x'_j: a bin of small norm is close to EVERY bin, so the R bins of smallest norm hold a
// large share of every bin's reference set.  One cheap pass - all target bins against those R pivots, 2 % of the tiles
// at 50 kb - therefore leaves every bin with a threshold (its k-th smallest distance among the pivots, an upper bound of
// its final k-th distance like any other threshold) that lets ~2k candidates through over the whole sweep instead of the
// ~8k of a uniform 1/8 sample.  Candidates found here are DISCARDED (the search proper meets the same pairs again and
// keeps them); only the thresholds survive, so the result cannot depend on the pivots.  (Measured on the bench matrix,
// CPU study: 512 pivots -> 233 passes per bin, 1024 -> 174, uniform 1/8 sample -> 792; final set: 100.)

// The R smallest norms (ties by bin index) -> ids[0..R), sorted by bin; one CTA.  Keys: the floats' bit patterns (norms
// are >= 0).  Radix select (four 8-bit passes over a shared-memory histogram) finds the R-th smallest key T; every thread
// then owns a contiguous range of bins, counts what it takes (keys below T, and the first R - #below keys equal to T in
// bin order), one block-wide scan of the counts places the ranges.  (r02: 34 bisection rounds plus a scan per 1024 bins,
// 0.21 ms at 57 633 bins - more than the pivot pass itself.)
__global__ void __launch_bounds__(1024) wc_pivot_select_kernel(const float* __restrict__ n32, int N, int R, int* __restrict__ ids) {
    __shared__ int s_hist[256];
    __shared__ int s_scan[2][32];
    __shared__ unsigned s_prefix;
    __shared__ int s_need;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto key_of = [&](int i) -> unsigned {
        const unsigned u = __float_as_uint(n32[i]);
        return u > 0x7f800000u ? 0xffffffffu : u;            // NaN / negative (never produced) sort last
    };
    if (tid == 0) { s_prefix = 0u; s_need = R; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix;
        const int need = s_need;
        const unsigned himask = pass == 0 ? 0u : 0xffffffffu << (shift + 8);
        for (int i = tid; i < N; i += 1024) {
            const unsigned key = key_of(i);
            if ((key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (warp == 0) {                                      // the bucket holding the need-th smallest of the matching keys
            int c[8], sum = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) { c[t] = s_hist[lane * 8 + t]; sum += c[t]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int run = incl - sum, bl = -1, below = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (bl < 0 && run + c[t] >= need) { bl = lane * 8 + t; below = run; }
                run += c[t];
            }
            const unsigned m = __ballot_sync(0xffffffffu, bl >= 0);
            const int src = m ? __ffs(m) - 1 : 31;
            const int b = __shfl_sync(0xffffffffu, bl, src), bel = __shfl_sync(0xffffffffu, below, src);
            if (lane == 0) {
                s_prefix = prefix | ((unsigned)(b < 0 ? 255 : b) << shift);
                s_need = need - (b < 0 ? 0 : bel);
            }
        }
        __syncthreads();
    }
    const unsigned T = s_prefix;                              // the R-th smallest key; s_need of the keys equal to T are taken
    const int need_eq = s_need;
    // ordered compaction: thread t owns bins [t * per, t * per + per)
    const int per = (N + 1023) / 1024;
    const int i0 = tid * per, i1 = min(N, i0 + per);
    int n_lt = 0, n_eq = 0;
    for (int i = i0; i < i1; ++i) {
        const unsigned key = key_of(i);
        n_lt += key < T ? 1 : 0;
        n_eq += key == T ? 1 : 0;
    }
    // block-wide exclusive scans of both counts
    int in_lt = n_lt, in_eq = n_eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, in_lt, o), b = __shfl_up_sync(0xffffffffu, in_eq, o);
        if (lane >= o) { in_lt += a; in_eq += b; }
    }
    if (lane == 31) { s_scan[0][warp] = in_lt; s_scan[1][warp] = in_eq; }
    __syncthreads();
    int base_lt = 0, base_eq = 0;
    for (int w = 0; w < warp; ++w) { base_lt += s_scan[0][w]; base_eq += s_scan[1][w]; }
    int ex_lt = base_lt + in_lt - n_lt, ex_eq = base_eq + in_eq - n_eq;      // elements before this thread's range
    // position of an element = (#below-T before it) + (#taken equals before it)
    for (int i = i0; i < i1; ++i) {
        const unsigned key = key_of(i);
        if (key < T) {
            const int pos = ex_lt + min(ex_eq, need_eq);
            if (pos < R) ids[pos] = i;
            ++ex_lt;
        } else if (key == T) {
            if (ex_eq < need_eq) {
                const int pos = ex_lt + ex_eq;
                if (pos < R) ids[pos] = i;
            }
            ++ex_eq;
        }
    }
}

// P[r] = the fp16 row of pivot r, its norm next to it
// (folded norms, ldx = ldh + 64: the pivots are B operands - their last chunk is the row's second version of it)
__global__ void wc_pivot_gather_kernel(const __half* __restrict__ Xh, int ldh, int ldx, const float* __restrict__ n32,
                                       const int* __restrict__ ids, __half* __restrict__ P, float* __restrict__ n32p) {
    const int r = blockIdx.x;
    const int src = ids[r];
    const uint4* s = reinterpret_cast<const uint4*>(Xh + (size_t)src * ldx);
    uint4* d = reinterpret_cast<uint4*>(P + (size_t)r * ldh);
    const int last = (ldh - BKH) / 8, shift = (ldx - ldh) / 8;
    for (int i = threadIdx.x; i < ldh / 8; i += blockDim.x) d[i] = s[i >= last ? i + shift : i];
    if (threadIdx.x == 0) n32p[r] = n32[src];
}
