// wisecondor_b200 - sharded symmetric search: host side of wc_newref_shard_dims / _begin / _sweep / _finish.
// Textually included by wc_search.cu after the kernels and schedule_pieces it launches.
#pragma once

// =====================================================================================================================
// Sharded symmetric search: the block pairs of the symmetric search divided over `world` ranks (one GPU each, every
// GPU holding the whole matrix).  Rank r owns the bin blocks [b0, b1) (nb / world consecutive blocks of 128 bins): it
// computes the tiles whose ROW block it owns - pass A, then pass B with the column side - and finalises its own bins.
// Three calls per rank with two collectives in between, issued by the host layer (wisecondor_b200/shard.py):
//   wc_newref_shard_begin   K4, pass A over the owned row blocks; thr[bin] = the owned bins' thresholds, the rest untouched
//     -> all-reduce(MIN) of thr over the ranks: every rank knows every bin's threshold
//   wc_newref_shard_sweep   pass B: row side into the rank's segments, column side into in_*[bin] for ALL bins
//     -> all-to-all of in_* by owner (equal splits of rows_per bins): a rank receives what every rank found for its bins
//   wc_newref_shard_finish  K6 over the owned bins: own segments + `world` incoming sources
// The column-side thresholds of bins owned elsewhere stay at their pass-A value during pass B (a rank cannot see the
// other ranks' prunes), so a bin receives ~ k * frac / 2 offers in total; in_cap leaves 4x head room per source.
// =====================================================================================================================
namespace {

struct ShardDims { int nb, bp, b0, b1, row0, row1, rows_per, in_cap, frac; size_t thr_len; };

ShardDims shard_dims(const wc_ctx* ctx, int N, int k, int world, int rank) {
    ShardDims d;
    d.frac = (ctx != nullptr && ctx->k5_sym >= 2) ? ctx->k5_sym : 8;
    d.nb = (N + BM - 1) / BM;
    d.bp = (d.nb + world - 1) / world;
    d.b0 = std::min(d.nb, rank * d.bp);
    d.b1 = std::min(d.nb, d.b0 + d.bp);
    d.row0 = std::min(N, d.b0 * BM);
    d.row1 = std::min(N, d.b1 * BM);
    d.rows_per = d.bp * BM;
    d.thr_len = (size_t)world * d.rows_per + BN;
    const int want = world == 1 ? 2 * k * d.frac : (4 * k * d.frac + world - 1) / world;
    d.in_cap = 256;
    while (d.in_cap < want) d.in_cap *= 2;
    return d;
}

// One K5 launch of a sharded symmetric search: pass 0 = threshold pass (rows only), pass 1 = symmetric pass.
// pass 2 = the pivot pass of K5t (thresholds only; seg_first names the scratch segment of every owned row block).
int shard_launch_pass(wc_ctx* ctx, int pass, unsigned long long* thr_d, unsigned long long* in_key_d, int* in_j_d,
                      int* in_cnt_d, cudaStream_t stream, const std::vector<int>* seg_first = nullptr, int pivots = 0) {
    const wc_shard_plan& pl = ctx->shard;
    if (!ctx->encode_tiled) {
        wc_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return WC_ERR_CUDA;
    }
    CUtensorMap tmap;
    {
        const int ldx = pl.f16 == 2 ? pl.ldh + BKH : pl.ldh;       // K5t: folded norms, a second version of the last chunk
        cuuint64_t dims[2] = {(cuuint64_t)(pl.f16 ? ldx : pl.ld), (cuuint64_t)pl.Npad};
        cuuint64_t strides[1] = {pl.f16 ? (cuuint64_t)ldx * sizeof(__half) : (cuuint64_t)pl.ld * sizeof(double)};
        cuuint32_t box[2] = {(cuuint32_t)(pl.f16 ? BKH : BK), BM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = reinterpret_cast<PFN_encodeTiled>(ctx->encode_tiled)(
            &tmap, pl.f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, ctx->buf[SLOT_XC].p, dims,
            strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            wc_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return WC_ERR_CUDA;
        }
    }
    const int nrb1 = std::max(pl.nrb, 1);
    int* d_meta = static_cast<int*>(ctx->buf[SLOT_RBMETA].p);
    int* d_ctaA = d_meta + 4 * nrb1;
    int* d_ctaB = d_ctaA + pl.gridA + 1;
    int* d_pieces = d_ctaB + pl.gridB + 1;
    int* d_offA = d_pieces + (size_t)std::max(pl.nseg, 1) * 5;
    int* d_offB = d_offA + pl.nrb + 1;
    int* d_listA = d_offB + pl.nrb + 1;
    int* d_listB = d_listA + pl.nlistA;
    TopkArgs ta;
    ta.norms = static_cast<double*>(ctx->buf[SLOT_NORMS].p);
    ta.row_cs = static_cast<int*>(ctx->buf[SLOT_ROWCS].p);
    ta.row_ce = static_cast<int*>(ctx->buf[SLOT_ROWCE].p);
    ta.N = pl.N; ta.row_begin = pl.row0; ta.row_end = pl.row1;
    ta.nkc = pl.nkc; ta.nd_last = pl.nd_last; ta.extra_h = pl.extra_h;
    ta.rb_skip_lo = d_meta; ta.rb_skip_n = d_meta + nrb1; ta.nrb = pl.nrb;
    ta.cta_piece_begin = pass == 0 ? d_ctaA : d_ctaB;
    ta.pieces = d_pieces;
    ta.cand_key = static_cast<u64*>(ctx->buf[SLOT_CAND_D].p); ta.cand_j = static_cast<int*>(ctx->buf[SLOT_CAND_J].p);
    ta.seg_cnt = static_cast<int*>(ctx->buf[SLOT_SEGCNT].p); ta.seg_flag = static_cast<int*>(ctx->buf[SLOT_SEGFLAG].p);
    ta.cap = pl.cap; ta.k = pl.k; ta.mcoef = pl.mcoef; ta.tau_init = pl.f16 ? 3e38 : 1e10 * (1.0 + 1e-6);
    ta.prof = nullptr; ta.trace = nullptr;
    ta.row_thr = thr_d + pl.row0;             // indexed by (bin - row_begin) on the row side ...
    ta.col_thr = thr_d;                       // ... and by global bin on the column side
    ta.lag = ctx->k5_lag; ta.nstages = pl.nstages;
    ta.tile_list = pass == 0 ? d_listA : d_listB;
    ta.rb_list_off = pass == 0 ? d_offA : d_offB;
    ta.final_prune = 1;
    ta.in_key = in_key_d; ta.in_j = in_j_d; ta.in_cnt = in_cnt_d; ta.in_cap = pl.in_cap;
    ta.madd_p = nullptr;
    ta.kb_last = pl.f16 ? (pl.f16 == 2 ? pl.ldh : pl.ldh - BKH) : 0;
    ta.madd = pl.madd; ta.n32 = pl.f16 ? static_cast<float*>(ctx->buf[SLOT_N32].p) : nullptr;
    ta.dbg = nullptr; ta.dbg_ld = 0; ta.coln32 = ta.n32; ta.col_ids = nullptr;
    const bool tc = pl.f16 == 2;
    if (pass == 2)
        return tc_pivot_pass(ctx, stream, tmap, ta, static_cast<const __half*>(ctx->buf[SLOT_XC].p), pl.ldh, pl.f16 == 2 ? pl.ldh + BKH : pl.ldh, pl.N, pl.nrb, *seg_first, pivots);
    if (tc) {
        auto kern = pass == 0 ? wc_dist_topk_tc_kernel<0, false> : wc_dist_topk_tc_kernel<1, false>;
        WC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        kern<<<pass == 0 ? pl.gridA : pl.gridB, TC_THREADS, pl.smem, stream>>>(tmap, tmap, ta);
        WC_CUDA(cudaGetLastError());
        return WC_OK;
    }
    auto kernel = pass == 0 ? (pl.f16 ? wc_dist_topk_f16_kernel<false> : wc_dist_topk_kernel<false>)
                            : (pl.f16 ? wc_dist_topk_f16_kernel<true> : wc_dist_topk_kernel<true>);
    WC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    kernel<<<pass == 0 ? pl.gridA : pl.gridB, TOPK_THREADS, pl.smem, stream>>>(tmap, ta);
    WC_CUDA(cudaGetLastError());
    return WC_OK;
}

}  // namespace

extern "C" int wc_newref_shard_dims(const wc_ctx* ctx, int N, int refsize, int world, int rank, long long* out6) {
    WC_CHECK_ARG(out6 != nullptr);                 // ctx may be NULL: sizes for the default first-pass fraction
    WC_CHECK_ARG(N > 0 && refsize >= 1 && refsize <= 384 && world >= 1 && rank >= 0 && rank < world);
    const ShardDims d = shard_dims(ctx, N, refsize, world, rank);
    out6[0] = d.rows_per; out6[1] = d.in_cap; out6[2] = (long long)d.thr_len; out6[3] = d.row0; out6[4] = d.row1; out6[5] = d.nb;
    return WC_OK;
}

extern "C" int wc_newref_shard_begin(wc_ctx* ctx, const double* corrected_d, int N, int S, const int* chrom_bins_h,
                                     int nchrom, int refsize, int rank, int world, unsigned long long* thr_d,
                                     void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && corrected_d != nullptr && chrom_bins_h != nullptr && thr_d != nullptr);
    WC_CHECK_ARG(N > 0 && S > 0 && nchrom > 0 && refsize >= 1 && refsize <= 384);
    WC_CHECK_ARG(world >= 1 && rank >= 0 && rank < world);
    long long tot = 0;
    for (int c = 0; c < nchrom; ++c) {
        WC_CHECK_ARG(chrom_bins_h[c] >= 0);
        tot += chrom_bins_h[c];
    }
    WC_CHECK_ARG(tot == N);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    wc_shard_plan& pl = ctx->shard;
    pl = wc_shard_plan();
    ctx->sched_hash = 0;                       // the schedule slots are about to be overwritten
    for (int i = 0; i < 4; ++i) ctx->phase_ms[i] = 0.0;
    ctx->phase_ms[8] = ctx->phase_ms[9] = 0.0;
    for (int i = 0; i < 5; ++i) ctx->counter[i] = 0;
    ctx->timed_mask &= ~0xfu;

    const ShardDims d = shard_dims(ctx, N, refsize, world, rank);
    const int k = refsize;
    const int cap = k <= 128 ? 512 : 1024;
    const int nblocks = (S + 7) / 8;
    const int Sx = nblocks * 8;
    const int nkc = nblocks / 2 + 1;
    const int nd_last = nblocks - 2 * (nkc - 1);
    const int ld = nkc * BK;
    const size_t Npad = (size_t)(N + BN - 1) / BN * BN + BN;
    const int nb = d.nb, nrb = d.b1 - d.b0;

    // exclusion ranges of every bin; skipped (own-chromosome interior) column tiles of every block
    std::vector<int> row_cs(N), row_ce(N), skip_lo(nb), skip_n(nb);
    {
        int pos = 0;
        for (int c = 0; c < nchrom; ++c) {
            for (int i = 0; i < chrom_bins_h[c]; ++i) { row_cs[pos + i] = pos; row_ce[pos + i] = pos + chrom_bins_h[c]; }
            pos += chrom_bins_h[c];
        }
    }
    long long tiles_plain = 0;
    for (int I = 0; I < nb; ++I) {
        const int r0 = I * BM, r1 = std::min(N, r0 + BM) - 1;
        int lo = nb, n = 0;
        if (row_cs[r0] == row_cs[r1]) {
            const int first = (row_cs[r0] + BN - 1) / BN, last = row_ce[r0] / BN;
            if (last > first) { lo = first; n = last - first; }
        }
        skip_lo[I] = lo;
        skip_n[I] = n;
        if (I >= d.b0 && I < d.b1) tiles_plain += nb - n;
    }
    std::vector<int> listA, listB, offA, offB;
    pair_lists(nb, skip_lo, skip_n, d.frac, d.b0, d.b1, listA, offA, listB, offB);
    const long long tilesA = (long long)listA.size(), tilesB = (long long)listB.size();
    const int gridA = tilesA ? (int)std::max(1ll, std::min<long long>(ctx->sm_count, (tilesA + 7) / 8)) : 0;
    const int gridB = tilesB ? (int)std::max(1ll, std::min<long long>(ctx->sm_count, (tilesB + 7) / 8)) : 0;
    std::vector<Piece> pieces;
    if (gridA) schedule_pieces(offA, nrb, gridA, 1, false, 0, pieces);
    if (gridB) {
        const double tile_bytes = (double)BM * ld * sizeof(double);
        const double matrix_bytes = (double)Npad * ld * sizeof(double);
        int G = ctx->k5_group;
        if (G <= 0) {
            G = 1;
            if (matrix_bytes > 64e6)
                while (G < 8 && (double)(gridB / G) * tile_bytes > 48e6) G *= 2;
        }
        schedule_pieces(offB, nrb, gridB, G, matrix_bytes > 64e6 || ctx->k5_group > 0, 1, pieces);
    }
    // segments numbered row-block-major; one piece table, pass A's pieces first
    const int spp = ctx->k5_f16 == 2 ? TC_SEGS_PER_PIECE : 1;        // K5t: one segment per column half of a piece
    const int npieces = (int)pieces.size();
    const int nseg = npieces * spp;
    std::vector<int> seg_first(std::max(nrb, 1), 0), seg_count(std::max(nrb, 1), 0), ctaA(gridA + 1, 0), ctaB(gridB + 1, 0);
    std::vector<int> piece_tab((size_t)std::max(npieces, 1) * 5, 0);
    {
        for (const Piece& pc : pieces) seg_count[pc.rb] += spp;
        int run = 0;
        for (int rb = 0; rb < nrb; ++rb) { seg_first[rb] = run; run += seg_count[rb]; }
        std::vector<int> next(seg_first);
        for (Piece& pc : pieces) { pc.seg = next[pc.rb]; next[pc.rb] += spp; }
        std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) {
            return x.pass != y.pass ? x.pass < y.pass : x.cta < y.cta;
        });
        int n0 = 0;
        for (int i = 0; i < npieces; ++i) {
            const Piece& pc = pieces[i];
            if (pc.pass == 0) { ctaA[pc.cta + 1]++; ++n0; } else { ctaB[pc.cta + 1]++; }
            piece_tab[(size_t)i * 5 + 0] = pc.rb;
            piece_tab[(size_t)i * 5 + 1] = pc.q0;
            piece_tab[(size_t)i * 5 + 2] = pc.q1;
            piece_tab[(size_t)i * 5 + 3] = pc.step;
            piece_tab[(size_t)i * 5 + 4] = pc.seg;
        }
        for (int c = 0; c < gridA; ++c) ctaA[c + 1] += ctaA[c];
        ctaB[0] = n0;
        for (int c = 0; c < gridB; ++c) ctaB[c + 1] += ctaB[c];
    }
    // device metadata: [zeros nrb][zeros nrb][seg_first nrb][seg_count nrb][ctaA][ctaB][pieces][offA][offB][listA][listB]
    const int nrb1 = std::max(nrb, 1);
    std::vector<int> meta;
    meta.insert(meta.end(), (size_t)2 * nrb1, 0);             // the arithmetic-mode skip tables (unused in list mode)
    meta.insert(meta.end(), seg_first.begin(), seg_first.end());
    meta.insert(meta.end(), seg_count.begin(), seg_count.end());
    meta.insert(meta.end(), ctaA.begin(), ctaA.end());
    meta.insert(meta.end(), ctaB.begin(), ctaB.end());
    meta.insert(meta.end(), piece_tab.begin(), piece_tab.end());
    meta.insert(meta.end(), offA.begin(), offA.end());
    meta.insert(meta.end(), offB.begin(), offB.end());
    meta.insert(meta.end(), listA.begin(), listA.end());
    meta.insert(meta.end(), listB.begin(), listB.end());

    double* Xc; double* norms; int* d_row_cs; int* d_row_ce; int* d_meta;
    u64* cand_key; int* cand_j; int* seg_cnt; int* seg_flag; int* slow;
    int rc;
    if ((rc = wc_reserve(ctx, SLOT_XC, std::max(Npad * ld * sizeof(double), Npad * ((size_t)(S + 4 + 2 * BKH) / BKH * BKH + BKH) * sizeof(__half)), (void**)&Xc))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_NORMS, Npad * sizeof(double), (void**)&norms))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_ROWCS, (size_t)N * sizeof(int), (void**)&d_row_cs))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_ROWCE, (size_t)N * sizeof(int), (void**)&d_row_ce))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_RBMETA, meta.size() * sizeof(int), (void**)&d_meta))) return rc;
    const size_t cand_n = (size_t)std::max(nseg, 1) * BM * cap;
    if ((rc = wc_reserve(ctx, SLOT_CAND_D, cand_n * sizeof(u64), (void**)&cand_key))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_CAND_J, cand_n * sizeof(int), (void**)&cand_j))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_SEGCNT, (size_t)std::max(nseg, 1) * BM * sizeof(int), (void**)&seg_cnt))) return rc;
    if ((rc = wc_reserve(ctx, SLOT_SEGFLAG, (size_t)std::max(nseg, 1) * BM * sizeof(int), (void**)&seg_flag))) return rc;
    const int rows = d.row1 - d.row0;
    if ((rc = wc_reserve(ctx, SLOT_SLOW, ((size_t)std::max(rows, 0) + 1) * sizeof(int), (void**)&slow))) return rc;
    WC_CUDA(cudaMemcpyAsync(d_row_cs, row_cs.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_row_ce, row_ce.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemcpyAsync(d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    WC_CUDA(cudaMemsetAsync(slow, 0, sizeof(int), stream));
    WC_CUDA(cudaMemsetAsync(seg_cnt, 0, (size_t)std::max(nseg, 1) * BM * sizeof(int), stream));
    WC_CUDA(cudaMemsetAsync(seg_flag, 0, (size_t)std::max(nseg, 1) * BM * sizeof(int), stream));

    // option k5_f16: fp16 tensor-core filter (every rank holds the whole matrix, so all ranks take the same decision)
    bool f16 = ctx->k5_f16 != 0;
    const bool tcf = ctx->k5_f16 == 2;                    // K5t: norms folded into the contraction (wc_prepare_f16_kernel)
    const int ldh = tcf ? (S + 4 + BKH - 1) / BKH * BKH : (S + BKH - 1) / BKH * BKH;
    const int ldx = tcf ? ldh + BKH : ldh;
    double nmax = 0.0;
    WC_CUDA(cudaEventRecord(ctx->ev[0], stream));
    if (f16) {
        float* n32;
        unsigned long long* stats;
        if ((rc = wc_reserve(ctx, SLOT_N32, Npad * sizeof(float), (void**)&n32))) return rc;
        if ((rc = wc_reserve(ctx, SLOT_F16STAT, 2 * sizeof(unsigned long long), (void**)&stats))) return rc;
        WC_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(unsigned long long), stream));
        wc_prepare_f16_kernel<<<(int)((Npad * 32 + 255) / 256), 256, 0, stream>>>(corrected_d, N, (int)Npad, S, ldh, ldx, tcf ? 1 : 0,
                                                                                  reinterpret_cast<__half*>(Xc), norms, n32, stats);
        WC_CUDA(cudaGetLastError());
        unsigned long long st_h[2] = {0, 0};
        WC_CUDA(cudaMemcpyAsync(st_h, stats, sizeof(st_h), cudaMemcpyDeviceToHost, stream));
        WC_CUDA(cudaStreamSynchronize(stream));
        memcpy(&nmax, &st_h[0], sizeof(double));
        if (st_h[1] != 0) f16 = false;
    }
    if (!f16)
        wc_prepare_kernel<<<(int)((Npad * 32 + 255) / 256), 256, 0, stream>>>(corrected_d, N, (int)Npad, S, ld, Sx, Xc, norms);
    WC_CUDA(cudaGetLastError());
    WC_CUDA(cudaEventRecord(ctx->ev[1], stream));
    const double eps16 = ldexp(1.0, -10) * (1.0 + ldexp(1.0, -11)) + (tcf ? 2.0 : 1.0) * (double)ldh * ldexp(1.0, -23) +
                         (tcf ? ldexp(1.0, -20) : ldexp(1.0, -21));          // see wc_newref_topk
    const double tau_init = f16 ? 3e38 : 1e10 * (1.0 + 1e-6);
    wc_fill_u64_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(thr_d, (size_t)N, host_key_of_tau(tau_init));
    wc_fill_u64_kernel<<<(unsigned)((d.thr_len - N + 255) / 256), 256, 0, stream>>>(thr_d + N, d.thr_len - (size_t)N, KEY_NEVER);
    WC_CUDA(cudaGetLastError());

    pl.N = N; pl.S = S; pl.k = k; pl.cap = cap; pl.in_cap = d.in_cap; pl.world = world; pl.rank = rank;
    pl.nb = nb; pl.bp = d.bp; pl.b0 = d.b0; pl.b1 = d.b1; pl.row0 = d.row0; pl.row1 = d.row1; pl.rows_per = d.rows_per;
    pl.nkc = nkc; pl.nd_last = nd_last; pl.extra_h = nd_last; pl.ld = ld; pl.nrb = nrb; pl.nseg = npieces;
    pl.gridA = gridA; pl.gridB = gridB; pl.Npad = Npad; pl.nlistA = listA.size();
    pl.tilesA = tilesA; pl.tilesB = tilesB; pl.tiles_plain = tiles_plain;
    pl.f16 = f16 ? ctx->k5_f16 : 0;
    pl.ldh = ldh;
    pl.mcoef = f16 ? 2.0 * eps16 : 16.0 * (double)(S + 16) * 1.1102230246251565e-16;
    pl.madd = f16 ? 2.0 * eps16 * nmax + ldexp(1.0, -20) * sqrt((double)S * nmax) : 0.0;
    if (f16) pl.nkc = ldh / BKH;
    pl.corrected = corrected_d;
    pl.nstages = f16 ? 4 : (cap <= 512 ? 4 : 3);
    pl.smem = f16 && ctx->k5_f16 == 2 ? TC_SMEM_BYTES : f16
        ? (size_t)pl.nstages * STAGE_BYTES + sizeof(TopkState) + (size_t)CONSUMER_WARPS * F16_SCRATCH +
              (size_t)CONSUMER_WARPS * (BN * sizeof(u64) + STG * sizeof(uint4) + (2 * BN + 32) * sizeof(float))
        : (size_t)pl.nstages * STAGE_BYTES + sizeof(TopkState) +
              (size_t)CONSUMER_WARPS * std::max<size_t>((size_t)cap * 12, 8192) +
              (size_t)CONSUMER_WARPS * (BN * sizeof(u64) + STG * sizeof(uint4));
    if (pl.smem > 227 * 1024) { wc_set_error("K5 shared memory %zu exceeds 227 KiB", pl.smem); return WC_ERR_INTERNAL; }

    WC_CUDA(cudaEventRecord(ctx->ev[2], stream));
    pl.pivots = (pl.f16 == 2 && ctx->k5_pivots != 0 && nrb > 0 && nb >= 24) ? pivot_count(N, k) : 0;
    if (pl.pivots > 0) {
        if ((rc = shard_launch_pass(ctx, 2, thr_d, nullptr, nullptr, nullptr, stream, &seg_first, pl.pivots))) return rc;
    }
    if (gridA > 0) {
        if ((rc = shard_launch_pass(ctx, 0, thr_d, nullptr, nullptr, nullptr, stream))) return rc;
    }
    WC_CUDA(cudaEventRecord(ctx->ev[16], stream));
    pl.valid = 1;
    pl.stage = 1;
    return WC_OK;
}

extern "C" int wc_newref_shard_sweep(wc_ctx* ctx, unsigned long long* thr_d, unsigned long long* in_key_d, int* in_j_d,
                                     int* in_cnt_d, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && thr_d != nullptr && in_key_d != nullptr && in_j_d != nullptr && in_cnt_d != nullptr);
    wc_shard_plan& pl = ctx->shard;
    if (!pl.valid || pl.stage != 1) { wc_set_error("wc_newref_shard_sweep: call wc_newref_shard_begin first"); return WC_ERR_ARG; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    pl.thr = thr_d;            // K6 drops entries beyond the bins' final thresholds
    WC_CUDA(cudaMemsetAsync(in_cnt_d, 0, (size_t)pl.world * pl.rows_per * sizeof(int), stream));
    WC_CUDA(cudaEventRecord(ctx->ev[18], stream));
    if (pl.gridB > 0) {
        int rc;
        if ((rc = shard_launch_pass(ctx, 1, thr_d, in_key_d, in_j_d, in_cnt_d, stream))) return rc;
    }
    WC_CUDA(cudaEventRecord(ctx->ev[19], stream));
    pl.stage = 2;
    return WC_OK;
}

extern "C" int wc_newref_shard_finish(wc_ctx* ctx, const unsigned long long* recv_key_d, const int* recv_j_d,
                                      const int* recv_cnt_d, int32_t* idx_d, double* dist_d, void* stream_v) {
    WC_CHECK_ARG(ctx != nullptr && recv_key_d != nullptr && recv_j_d != nullptr && recv_cnt_d != nullptr);
    wc_shard_plan& pl = ctx->shard;
    if (!pl.valid || pl.stage != 2) { wc_set_error("wc_newref_shard_finish: call wc_newref_shard_sweep first"); return WC_ERR_ARG; }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    WC_CUDA(cudaSetDevice(ctx->device));
    pl.stage = 0;
    pl.valid = 0;
    const int rows = pl.row1 - pl.row0;
    int nslow = 0;
    int* d_meta = static_cast<int*>(ctx->buf[SLOT_RBMETA].p);
    int* d_row_cs = static_cast<int*>(ctx->buf[SLOT_ROWCS].p);
    int* d_row_ce = static_cast<int*>(ctx->buf[SLOT_ROWCE].p);
    int* slow = static_cast<int*>(ctx->buf[SLOT_SLOW].p);
    long long launches = 4 + (pl.gridA > 0) + (pl.gridB > 0) + (pl.pivots > 0 ? 3 : 0);      // K4, two fills, [pivots], pass A, pass B, K6
    if (rows > 0) {
        WC_CHECK_ARG(idx_d != nullptr && dist_d != nullptr);
        const int nrb1 = std::max(pl.nrb, 1);
        FinArgs fa;
        fa.X = pl.corrected; fa.N = pl.N; fa.S = pl.S; fa.norms = static_cast<double*>(ctx->buf[SLOT_NORMS].p);
        fa.row_cs = d_row_cs; fa.row_ce = d_row_ce; fa.row_begin = pl.row0; fa.row_end = pl.row1;
        fa.rb_seg_first = d_meta + 2 * nrb1; fa.rb_seg_count = d_meta + 3 * nrb1;
        fa.cand_key = static_cast<u64*>(ctx->buf[SLOT_CAND_D].p); fa.cand_j = static_cast<int*>(ctx->buf[SLOT_CAND_J].p);
        fa.seg_cnt = static_cast<int*>(ctx->buf[SLOT_SEGCNT].p); fa.seg_flag = static_cast<int*>(ctx->buf[SLOT_SEGFLAG].p);
        fa.cap = pl.cap; fa.k = pl.k; fa.shortcap = pl.k <= (pl.f16 ? 96 : 128) ? 256 : 512; fa.mcoef = pl.mcoef;
        fa.idx_out = idx_d; fa.dist_out = dist_d; fa.slow_list = slow + 1; fa.slow_count = slow; fa.slow_bias = 0;
        fa.vec = (pl.S % 4 == 0 && (reinterpret_cast<uintptr_t>(pl.corrected) & 31) == 0) ? 4 : 1;
        fa.in_key = recv_key_d; fa.in_j = recv_j_d; fa.in_cnt = recv_cnt_d; fa.in_cap = pl.in_cap;
        fa.in_nsrc = pl.world; fa.in_src_rows = pl.rows_per; fa.madd = pl.madd; fa.madd_p = nullptr;
        fa.row_thr = pl.thr != nullptr ? pl.thr + pl.row0 : nullptr;
        WC_CUDA(cudaEventRecord(ctx->ev[4], stream));
        long long fin_launches = 0;
        int rcf;
        if ((rcf = launch_finalize(ctx, stream, fa, rows, pl.f16 != 0, &fin_launches))) return rcf;
        launches += fin_launches - 1;
        WC_CUDA(cudaEventRecord(ctx->ev[5], stream));
        int k6_stats[4] = {0, 0, 0, 0};
        WC_CUDA(cudaMemcpyAsync(&nslow, slow, sizeof(int), cudaMemcpyDeviceToHost, stream));
        if (ctx->k6_stats_d != nullptr) WC_CUDA(cudaMemcpyAsync(k6_stats, ctx->k6_stats_d, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
        WC_CUDA(cudaStreamSynchronize(stream));
        ctx->counter[8] = k6_stats[1];
        ctx->counter[9] = k6_stats[2];
    ctx->counter[10] = k6_stats[0];
    ctx->counter[11] = k6_stats[3];
        if (nslow > 0) {
            const int batch = 64;
            double* scratch;
            int rc;
            if ((rc = wc_reserve(ctx, SLOT_SCRATCH, (size_t)std::min(nslow, batch) * pl.N * sizeof(double), (void**)&scratch)))
                return rc;
            WC_CUDA(cudaEventRecord(ctx->ev[6], stream));
            for (int off = 0; off < nslow; off += batch) {
                ExhArgs ea;
                ea.X = pl.corrected; ea.N = pl.N; ea.S = pl.S; ea.row_cs = d_row_cs; ea.row_ce = d_row_ce; ea.row_begin = pl.row0;
                ea.slow_list = slow + 1; ea.list_off = off; ea.scratch = scratch; ea.k = pl.k; ea.idx_out = idx_d; ea.dist_out = dist_d;
                wc_exhaustive_kernel<<<std::min(batch, nslow - off), EXH_THREADS, 0, stream>>>(ea);
                ++launches;
            }
            WC_CUDA(cudaGetLastError());
            WC_CUDA(cudaEventRecord(ctx->ev[7], stream));
        }
    }
    WC_CUDA(cudaStreamSynchronize(stream));
    float ms, ms2;
    WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1])); ctx->phase_ms[0] = ms;
    WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[16])); ctx->phase_ms[8] = ms;
    WC_CUDA(cudaEventElapsedTime(&ms2, ctx->ev[18], ctx->ev[19])); ctx->phase_ms[9] = ms2;
    ctx->phase_ms[1] = (double)ms + (double)ms2;
    if (rows > 0) { WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5])); ctx->phase_ms[2] = ms; }
    if (nslow > 0) { WC_CUDA(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7])); ctx->phase_ms[3] = ms; }
    ctx->counter[0] = launches;
    ctx->counter[1] = nslow;
    ctx->counter[2] = pl.tiles_plain;
    ctx->counter[3] = pl.tilesA + pl.tilesB;
    ctx->counter[4] = std::max(pl.gridA, pl.gridB);
    ctx->counter[7] = pl.f16 | (pl.pivots << 4);
    return WC_OK;
}

