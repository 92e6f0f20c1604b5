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
// Blackwell-native structure (one persistent CTA per SM, 8 warps, no thread ever holds an MMA fragment):
//   warp 0, one lane   TMA producer: per 64-sample chunk three SWIZZLE_128B boxes (A: 128 bins, B: 2 x 128 bins, 16 KiB
//                      each) into a 4-stage mbarrier ring (SASS UTMALDG)
//   warp 1, one lane   MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 256 (or 128 for an odd last block),
//                      K = 16, operands straight from the swizzled shared-memory tiles through matrix descriptors,
//                      accumulators in TENSOR MEMORY: two buffers of 256 fp32 columns x 128 lanes (all 512 columns);
//                      tcgen05.commit releases ring stages and publishes finished accumulators (SASS UTCHMMA / UTCBAR)
//   warp 2             allocates / frees the tensor memory
//   warps 4-7          epilogue: warp e owns TMEM lanes [32e, 32e+32) = tile rows, ONE THREAD PER TARGET BIN.  The row's
//                      threshold, norm, candidate count and exclusion range live in that thread's registers; the
//                      accumulator row arrives with tcgen05.ld.32x32b.x32 (SASS LDTM), 32 columns at a time, while the
//                      tensor core already works on the next tile in the other TMEM buffer.
// Per 128 x 256 tile the tensor core needs 40 x 128 = 5120 cycles, the epilogue ~2000 issue slots per warp: it hides.
// What bounds the kernel is the L2 -> shared-memory operand stream (48 KiB per 64-sample chunk of a 128 x 256 tile).
constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 4;
constexpr int TC_STAGE_BYTES = 3 * TILE_BYTES;      // A, B0, B1 boxes: 128 rows x 64 halves x 2 B = 16 KiB each
constexpr int TC_NB = 2 * BN;                       // columns of a full item (two candidate blocks)
constexpr int TC_STG = 128;                         // per-warp staging entries for column-side candidates
constexpr int TC_EPI_WARPS = 4;
constexpr uint32_t TC_TMEM_COLS = 512;

struct __align__(16) TcState {
    uint64_t full[TC_STAGES];
    uint64_t empty[TC_STAGES];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
    int stg_cnt[TC_EPI_WARPS];
};

constexpr size_t TC_SMEM_BYTES = (size_t)TC_STAGES * TC_STAGE_BYTES                 // operand ring
                                 + 2 * 2 * TC_NB * sizeof(float)                     // column norms + thresholds, 2 buffers
                                 + (size_t)TC_EPI_WARPS * TC_STG * sizeof(uint4)     // column-side staging
                                 + (size_t)32 * 128 * sizeof(float)                  // per-thread parking of one 32-column chunk
                                 + sizeof(TcState) + 1024;                           // + alignment slack

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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <bool SYM, bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1)
wc_dist_topk_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TopkArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* s_nj = reinterpret_cast<float*>(tiles + (size_t)TC_STAGES * TC_STAGE_BYTES);       // [2][TC_NB]
    float* s_tj = s_nj + 2 * TC_NB;                                                            // [2][TC_NB]
    uint4* s_stg = reinterpret_cast<uint4*>(s_tj + 2 * TC_NB);                                 // [4][TC_STG]
    float* s_park = reinterpret_cast<float*>(s_stg + TC_EPI_WARPS * TC_STG);                   // [32][128]
    TcState& sm = *reinterpret_cast<TcState*>(s_park + 32 * 128);
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
            mbar_init(&sm.tempty[b], TC_EPI_WARPS);
        }
        for (int w = 0; w < TC_EPI_WARPS; ++w) sm.stg_cnt[w] = 0;
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    if (warp_all == 2) {                        // one warp allocates all 512 TMEM columns (one CTA per SM: no contention)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(TC_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    // every role walks the same sequence of items: (row block, one or two column blocks) per piece
    auto tile_of = [&](const int* tl, int q, int skip_lo, int skip_n) { return tl ? tl[q] : (q < skip_lo ? q : q + skip_n); };

    if (warp_all == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pi = pb; pi < pe; ++pi) {
                const int* pc = a.pieces + (size_t)pi * 5;
                const int rbp = pc[0], q1 = pc[2], qs = pc[3];
                const int skip_lo = a.rb_skip_lo[rbp], skip_n = a.rb_skip_n[rbp];
                const int* tl = a.tile_list ? a.tile_list + a.rb_list_off[rbp] : nullptr;
                const int row0 = a.row_begin + rbp * BM;
                for (int q = pc[1]; q < q1;) {
                    const int c0 = tile_of(tl, q, skip_lo, skip_n) * BN;
                    q += qs;
                    const bool two = q < q1;
                    const int c1 = two ? tile_of(tl, q, skip_lo, skip_n) * BN : 0;
                    if (two) q += qs;
                    for (int kc = 0; kc < a.nkc; ++kc) {
                        mbar_wait(&sm.empty[stage], phase ^ 1u);
                        unsigned char* st = tiles + (size_t)stage * TC_STAGE_BYTES;
                        mbar_arrive_expect_tx(&sm.full[stage], (two ? 3 : 2) * TILE_BYTES);
                        tma_load_2d(st, &tmap, kc * BKH, row0, &sm.full[stage]);
                        tma_load_2d(st + TILE_BYTES, &tmap, kc * BKH, c0, &sm.full[stage]);
                        if (two) tma_load_2d(st + 2 * TILE_BYTES, &tmap, kc * BKH, c1, &sm.full[stage]);
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
                    mbar_wait(&sm.tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));      // the epilogue has drained this buffer
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + (uint32_t)buf * TC_NB;
                    const uint32_t idesc = two ? idesc2 : idesc1;
                    for (int kc = 0; kc < a.nkc; ++kc) {
                        mbar_wait(&sm.full[stage], phase);
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
    } else if (warp_all >= 4) {
        // ===== epilogue: one thread per target bin =====
        const int e = warp_all - 4;                 // TMEM lane quarter of this warp (== warp_all % 4)
        const int r = e * 32 + lane;                // row of the tile = TMEM lane
        uint4* w_stg = s_stg + e * TC_STG;
        int* w_stgc = &sm.stg_cnt[e];
        float* park = s_park + r;                   // entry b of this thread's chunk lives at park[b * 128]
        const size_t seg_stride = (size_t)BM * a.cap;
        long long pf_epi = 0, pf_prune = 0, pf_nprune = 0, pf_emit = 0, pf_wait = 0;
        const long long pf_t0 = clock64();

        auto flush_incoming = [&]() {
            __syncwarp();
            int n = *w_stgc;
            if (n > TC_STG) n = TC_STG;
            for (int x = lane; x < n; x += 32) {
                const uint4 v = w_stg[x];
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

        int it = 0;
        for (int pi = pb; pi < pe; ++pi) {
            const int* pc = a.pieces + (size_t)pi * 5;
            const int rb = pc[0], q1 = pc[2], qs = pc[3], seg = pc[4];
            const int skip_lo = a.rb_skip_lo[rb], skip_n = a.rb_skip_n[rb];
            const int* tl = a.tile_list ? a.tile_list + a.rb_list_off[rb] : nullptr;
            // this thread's bin
            const int row = a.row_begin + rb * BM + r;
            const bool valid = row < a.row_end;
            const double nrm = valid ? a.norms[row] : 0.0;
            const float ni = valid ? a.n32[row] : INFINITY;          // +inf: no entry of an invalid row ever passes
            const int cs = valid ? a.row_cs[row] : 0;
            const unsigned clen = valid ? (unsigned)(a.row_ce[row] - cs) : 0u;
            u64 thr = valid ? __ldcg(a.row_thr + (row - a.row_begin)) : KEY_NEVER;
            int cnt = 0;
            int flag = 0;
            u64* rk = a.cand_key + (size_t)seg * seg_stride + (size_t)r * a.cap;
            int* rj = a.cand_j + (size_t)seg * seg_stride + (size_t)r * a.cap;

            // warp-collective prune of the rows named in `need`; the owning lane adopts the result
            auto prune_rows = [&](unsigned need) {
                while (need) {
                    const int src = __ffs(need) - 1;
                    need &= need - 1;
                    int n = __shfl_sync(0xffffffffu, cnt, src);
                    if (n > a.cap) n = a.cap;
                    const double nr = __shfl_sync(0xffffffffu, nrm, src);
                    u64* pk = a.cand_key + (size_t)seg * seg_stride + (size_t)(e * 32 + src) * a.cap;
                    int* pj = a.cand_j + (size_t)seg * seg_stride + (size_t)(e * 32 + src) * a.cap;
                    u64 nthr;
                    int kept;
                    ++pf_nprune;
                    __threadfence_block();
                    __syncwarp();
                    if (a.cap <= 512)
                        prune_row<16>(pk, pj, n, a.k, nr, a.mcoef, a.madd, lane, nullptr, nullptr, &nthr, &kept);
                    else
                        prune_row<32>(pk, pj, n, a.k, nr, a.mcoef, a.madd, lane, nullptr, nullptr, &nthr, &kept);
                    if (lane == src) {
                        if (kept > a.cap - TC_NB) {          // a tie plateau wider than the buffer: exact fallback
                            flag = 1;
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
                // column tables of this item (norms; SYM: thresholds) and the row's shared threshold: issued before the wait
                const int tid_e = r;
                const float nj0 = a.n32[c0 + tid_e];
                const float nj1 = two ? a.n32[c1 + tid_e] : INFINITY;
                u64 ct0 = KEY_NEVER, ct1 = KEY_NEVER;
                if (SYM) {
                    ct0 = __ldcg(a.col_thr + c0 + tid_e);
                    if (two) ct1 = __ldcg(a.col_thr + c1 + tid_e);
                }
                if (valid) {
                    const u64 shared_thr = __ldcg(a.row_thr + (row - a.row_begin));
                    if (shared_thr < thr) thr = shared_thr;
                }
                float* nj_t = s_nj + buf * TC_NB;
                float* tj_t = s_tj + buf * TC_NB;
                nj_t[tid_e] = nj0;
                nj_t[BN + tid_e] = nj1;
                if (SYM) {
                    tj_t[tid_e] = tau32_of_key(ct0);
                    tj_t[BN + tid_e] = tau32_of_key(ct1);
                }
                const float taui = tau32_of_key(thr);
                named_bar_sync(1, TC_EPI_WARPS * 32);              // tables visible to the four epilogue warps
                const long long pf_w0 = clock64();
                mbar_wait(&sm.tfull[buf], (uint32_t)((it >> 1) & 1));
                tc_fence_after();
                const long long pf_e0 = clock64();
                pf_wait += pf_e0 - pf_w0;

                const int nchunk = two ? TC_NB / 32 : BN / 32;
                for (int ch = 0; ch < nchunk; ++ch) {
                    uint32_t v[32];
                    tc_ld32(tmem_base + ((uint32_t)(e * 32) << 16) + (uint32_t)(buf * TC_NB + ch * 32), v);
                    tc_wait_ld();
                    if (ch == nchunk - 1) {          // the accumulator row is in registers: hand the buffer back to the tensor core
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.tempty[buf]);
                    }
                    const int cb = ch * 32;          // column of the item (0..255)
                    const int colbase = (cb < BN ? c0 : c1 - BN) + cb;      // global bin of the chunk's first column
                    unsigned mask = 0, cmask = 0;
                    float dv[32];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 nj = *reinterpret_cast<const float4*>(nj_t + cb + 4 * u);
                        dv[4 * u + 0] = fmaf(-2.0f, __uint_as_float(v[4 * u + 0]), ni + nj.x);
                        dv[4 * u + 1] = fmaf(-2.0f, __uint_as_float(v[4 * u + 1]), ni + nj.y);
                        dv[4 * u + 2] = fmaf(-2.0f, __uint_as_float(v[4 * u + 2]), ni + nj.z);
                        dv[4 * u + 3] = fmaf(-2.0f, __uint_as_float(v[4 * u + 3]), ni + nj.w);
                    }
#pragma unroll
                    for (int b = 0; b < 32; ++b)
                        if (dv[b] <= taui) mask |= 1u << b;
                    if (SYM) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 tj = *reinterpret_cast<const float4*>(tj_t + cb + 4 * u);
                            if (dv[4 * u + 0] <= tj.x) cmask |= 1u << (4 * u + 0);
                            if (dv[4 * u + 1] <= tj.y) cmask |= 1u << (4 * u + 1);
                            if (dv[4 * u + 2] <= tj.z) cmask |= 1u << (4 * u + 2);
                            if (dv[4 * u + 3] <= tj.w) cmask |= 1u << (4 * u + 3);
                        }
                    }
                    if (DBG) {
                        if (valid && a.dbg != nullptr) {
#pragma unroll
                            for (int b = 0; b < 32; ++b)
                                if (colbase + b < a.dbg_ld) a.dbg[(size_t)(row - a.row_begin) * a.dbg_ld + colbase + b] = dv[b];
                        }
                    }
                    if (mask | cmask) {
                        // park the chunk (conflict-free: entry b of thread r at [b][r]) and walk the set bits
#pragma unroll
                        for (int b = 0; b < 32; ++b) park[b * 128] = dv[b];
                        unsigned m2 = mask;
                        while (m2) {
                            const int bit = __ffs(m2) - 1;
                            m2 &= m2 - 1;
                            const float d = park[bit * 128];
                            const int col = colbase + bit;
                            // drop non-finite distances and the row's own chromosome
                            if (!(fabsf(d) < INFINITY) || (unsigned)(col - cs) < clen) continue;
                            if (cnt < a.cap) {
                                rk[cnt] = (u64)__double_as_longlong(-0.5 * (double)d);
                                rj[cnt] = col;
                            } else {
                                flag = 1;
                            }
                            ++cnt;
                            ++pf_emit;
                        }
                        if (SYM) {
                            while (cmask) {
                                const int bit = __ffs(cmask) - 1;
                                cmask &= cmask - 1;
                                const float d = park[bit * 128];
                                const int j = colbase + bit;
                                if (!(fabsf(d) < INFINITY) || j >= a.N || !valid) continue;
                                if ((unsigned)(j - cs) < clen) continue;      // same chromosome (symmetric relation)
                                const u64 key = (u64)__double_as_longlong(-0.5 * (double)d);
                                const int pos = atomicAdd(w_stgc, 1);
                                if (pos < TC_STG) {
                                    w_stg[pos] = make_uint4((unsigned)key, (unsigned)(key >> 32), (unsigned)j, (unsigned)row);
                                } else {                                     // staging full (loose thresholds): append directly
                                    const int w = atomicAdd(a.in_cnt + j, 1);
                                    if (w < a.in_cap) {
                                        a.in_key[(size_t)j * a.in_cap + w] = key;
                                        a.in_j[(size_t)j * a.in_cap + w] = row;
                                    }
                                }
                                ++pf_emit;
                            }
                        }
                    }
                    if (SYM) {
                        __syncwarp();
                        if (*w_stgc >= 32) flush_incoming();
                    }
                }
                const long long pf_p0 = clock64();
                pf_epi += pf_p0 - pf_e0;
                // ---- prune rows whose buffer could overflow during the next item ----
                prune_rows(__ballot_sync(0xffffffffu, cnt > a.cap - TC_NB && !flag));
                pf_prune += clock64() - pf_p0;
            }
            // piece finished
            __syncwarp();
            if (SYM) flush_incoming();
            if (a.final_prune)      // every row leaves its best threshold behind (the column side of later tiles is filtered by it)
                prune_rows(__ballot_sync(0xffffffffu, cnt > a.k + 24 && cnt <= a.cap && !flag));
            a.seg_cnt[(size_t)seg * BM + r] = cnt > a.cap ? a.cap : cnt;
            a.seg_flag[(size_t)seg * BM + r] = flag;
        }
        if (a.prof != nullptr && e == 0 && lane == 0) {
            long long* o = a.prof + (size_t)blockIdx.x * 8;
            o[0] = clock64() - pf_t0; o[1] = pf_wait; o[2] = pf_epi; o[3] = pf_prune;
            o[4] = it; o[5] = pf_nprune; o[6] = pf_emit; o[7] = 0;
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
