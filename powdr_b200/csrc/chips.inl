// Multi-chip segment under ONE transcript (the reference builds one proving context for all chips of a segment,
// /root/reference/openvm/src/trace_generation.rs:113-140, and proves it with one engine.prove call, openvm-riscv/src/lib.rs:327-332).
// Included by capi.cu inside its extern "C" block.  Per chip the kernels are the single-chip ones (LDE, LogUp, quotient, openings);
// what is shared: every challenge, three mixed-height MMCS commitments (main, permutation, quotient: Plonky3 MerkleTreeMmcs injection
// rule), and ONE FRI instance into which the reduced-opening codeword of each LDE height is added when the fold reaches that height.
// CPU restatement + independent verifier live in the test tree; transcript documented in DESIGN.md §3.

namespace {

struct McMat { const uint32_t* p; size_t width; size_t log_h; };

// mixed-height tree into `tree` (node-major layers of the tallest height back to back); h0 out
int mmcs_commit_dev(pb_ctx* ctx, const std::vector<McMat>& ms, DevBuf<uint32_t>& tree, size_t* h0_out) {
    size_t h0 = 0;
    for (const McMat& m : ms) h0 = std::max(h0, m.log_h);
    *h0_out = h0;
    const size_t H = (size_t)1 << h0;
    int rc = tree.ensure(8 * (2 * H));
    if (rc) return rc;
    size_t lowest = h0;
    for (const McMat& m : ms) lowest = std::min(lowest, m.log_h);
    // digests of every height group: the tallest goes straight into layer 0, the others into scratch until their level is reached
    std::vector<size_t> dig_off(h0 + 1, (size_t)-1);
    size_t scratch_words = 0;
    for (size_t h = lowest; h < h0; h++) {
        bool any = false;
        for (const McMat& m : ms) any = any || m.log_h == h;
        if (any) { dig_off[h] = scratch_words; scratch_words += (size_t)8 << h; }
    }
    rc = ctx->ws_mc_dig.ensure(std::max<size_t>(1, scratch_words));
    if (rc) return rc;
    for (size_t h = lowest; h <= h0; h++) {
        std::vector<const uint32_t*> cols;
        for (const McMat& m : ms)
            if (m.log_h == h)
                for (size_t c = 0; c < m.width; c++) cols.push_back(m.p + (c << h));
        if (cols.empty()) continue;
        rc = ctx->coltab.ensure(cols.size());
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->coltab.p, cols.data(), cols.size() * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
        // (pageable source: staged before the call returns; the table itself is reused in stream order)
        uint32_t* dst = h == h0 ? tree.p : ctx->ws_mc_dig.p + dig_off[h];
        const size_t rows = (size_t)1 << h;
        if (rows <= LEAF_COOP)
            p2::leaf_hash_cols_coop_kernel<<<(unsigned)((16 * rows + 255) / 256), 256, 0, ctx->stream>>>(ctx->coltab.p, (uint32_t)cols.size(), rows, dst);
        else
            p2::leaf_hash_cols_kernel<<<(unsigned)((rows + p2::LEAF_THREADS - 1) / p2::LEAF_THREADS), p2::LEAF_THREADS, 0, ctx->stream>>>(ctx->coltab.p, (uint32_t)cols.size(),
                                                                                                                                          rows, dst);
        LAUNCHED(ctx);
    }
    // levels down to the lowest injected height one by one, then the fused upper-tree launches
    uint32_t* prev = tree.p;
    size_t lh = h0;
    while (lh > lowest) {
        const size_t n = (size_t)1 << (lh - 1);
        uint32_t* cur = prev + ((size_t)16 << (lh - 1));
        p2::compress_layer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(prev), reinterpret_cast<uint4*>(cur), n);
        LAUNCHED(ctx);
        lh--;
        if (dig_off[lh] != (size_t)-1) {
            p2::inject_layer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint4*>(cur),
                                                                                         reinterpret_cast<const uint4*>(ctx->ws_mc_dig.p + dig_off[lh]), n);
            LAUNCHED(ctx);
        }
        prev = cur;
    }
    CK(cudaGetLastError());
    return merkle_upper(ctx, prev, lh);
}

}  // namespace

int pb_chips_sizes(const pb_chip_t* chips, size_t K, size_t* n_opened, size_t* words_per_query) {
    if (!chips || !K) return PB_ERR_INVALID_ARG;
    size_t hmax = 0, hperm = 0, wm = 0, wp = 0;
    for (size_t c = 0; c < K; c++) {
        if (!chips[c].air) return PB_ERR_INVALID_ARG;
        const size_t lm = chips[c].log_n + 1, p = chips[c].air->has_lu ? chips[c].air->lu.perm_width() : 0;
        hmax = std::max(hmax, lm);
        if (p) hperm = std::max(hperm, lm);
        wm += chips[c].width;
        wp += p;
    }
    if (n_opened) *n_opened = wm + 2 * wp + 8 * K;
    if (words_per_query) {
        size_t w = 1 + wm + 8 * hmax + (wp ? wp + 8 * hperm : 0) + 8 * K + 8 * hmax;
        for (size_t i = 0; i + 1 < hmax; i++) w += 8 + 8 * (hmax - 1 - i);
        *words_per_query = w;
    }
    return 0;
}

int pb_prove_chips(pb_ctx_t* ctx, const pb_chip_t* chips, size_t K, pb_chips_proof_t* proof, uint32_t* h_cumsums) {
    if (ctx) ctx->mc.valid = ctx->seg.valid = ctx->sh.valid = false;      // a failure below must not leave a stale proof state queryable; the
                                                                          // single-chip provers share the FRI workspaces with this one
    if (!ctx || !chips || !K || !proof || !h_cumsums) return PB_ERR_INVALID_ARG;
    for (size_t c = 0; c < K; c++)
        if (!chips[c].air || !chips[c].d_trace || chips[c].log_n < 1 || chips[c].log_n > 24 || chips[c].width == 0 || chips[c].width != chips[c].air->width)
            return PB_ERR_INVALID_ARG;
    int rc;
    memset(proof, 0, sizeof *proof);
    memset(h_cumsums, 0, 16 * K);
    proof->pow_bits = ctx->pow_bits;
    proof->n_queries = ctx->n_queries;
    proof->n_chips = (uint32_t)K;
    cudaStream_t st = ctx->stream;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
    std::vector<McChip>& W = ctx->mc.chips;
    if (W.size() < K) W.resize(K);
    Challenger ch;
    ch.k = &ctx->p2;
    uint32_t root_m[8];
    size_t hmax = 0, hperm = 0, hq = 0, n_max = 0;
    bool any_lu = false;
    CK(cudaEventRecord(ctx->ev[0], st));

    // ---- main commit ----
    std::vector<McMat> mats;
    for (size_t c = 0; c < K; c++) {
        McChip& w = W[c];
        w.log_n = chips[c].log_n; w.width = chips[c].width; w.air = chips[c].air; w.d_trace = chips[c].d_trace;
        w.wp = w.air->has_lu ? w.air->lu.perm_width() : 0;
        w.n_chunks = w.air->has_lu ? w.air->lu.n_chunks() : 0;
        any_lu = any_lu || w.wp;
        const size_t N = (size_t)1 << w.log_n, M = N << 1;
        n_max = std::max(n_max, N);
        RC(w.lde.ensure(w.width * M));
        RC(pb_lde_batch(ctx, w.d_trace, w.log_n, w.width, 1, bb::GEN, w.lde.p));
        mats.push_back(McMat{w.lde.p, w.width, w.log_n + 1});
    }
    RC(mmcs_commit_dev(ctx, mats, ctx->mc.tree_main, &hmax));
    RC(read_root(ctx, ctx->mc.tree_main.p, hmax, root_m));
    for (int i = 0; i < 8; i++) proof->main_root[i] = h_from_m(root_m[i]);
    ch.observe(root_m, 8);
    proof->log_max = (uint32_t)hmax;
    CK(cudaEventRecord(ctx->ev[1], st));

    // ---- LogUp: shared challenges, one permutation trace per chip with interactions, one commitment ----
    if (any_lu) {
        const bb::E4 al = ch.sample_ext(), be = ch.sample_ext();
        for (int i = 0; i < 4; i++) { proof->logup_alpha[i] = h_from_m(al.c[i]); proof->logup_beta[i] = h_from_m(be.c[i]); }
        RC(ctx->ws_rowsum.ensure(4 * n_max));
        mats.clear();
        for (size_t c = 0; c < K; c++) {
            McChip& w = W[c];
            if (!w.wp) continue;
            const size_t N = (size_t)1 << w.log_n, M = N << 1;
            RC(w.perm.ensure(w.wp * N));
            RC(w.perm_lde.ensure(w.wp * M));
            RC(upload_logup_consts(ctx, w.air, al, be));
            RC(logup::launch_perm(w.air->lujit, st, w.d_trace, N, w.air->d_kc, w.air->d_bt, w.perm.p, ctx->ws_rowsum.p));
            LAUNCHED(ctx);
            const unsigned nb = (unsigned)((N + logup::SCAN_THREADS * logup::SCAN_ITEMS - 1) / (logup::SCAN_THREADS * logup::SCAN_ITEMS));
            RC(ctx->ws_scan_tot.ensure(4 * (size_t)nb));
            uint32_t* phi = w.perm.p + 4 * w.n_chunks * N;
            logup::scan_local_kernel<<<dim3(nb, 4), logup::SCAN_THREADS, 0, st>>>(ctx->ws_rowsum.p, phi, N, ctx->ws_scan_tot.p);
            logup::scan_totals_kernel<<<1, 32, 0, st>>>(ctx->ws_scan_tot.p, nb);
            logup::scan_add_kernel<<<dim3((unsigned)((N + 255) / 256), 4), 256, 0, st>>>(phi, N, ctx->ws_scan_tot.p, nb);
            ctx->launches += 3;
            uint32_t cs[4];
            for (int l = 0; l < 4; l++) CK(cudaMemcpyAsync(&cs[l], phi + (size_t)l * N + (N - 1), 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            for (int l = 0; l < 4; l++) { w.cumsum.c[l] = cs[l]; h_cumsums[4 * c + l] = h_from_m(cs[l]); }
            RC(pb_lde_batch(ctx, w.perm.p, w.log_n, w.wp, 1, bb::GEN, w.perm_lde.p));
            mats.push_back(McMat{w.perm_lde.p, w.wp, w.log_n + 1});
        }
        RC(mmcs_commit_dev(ctx, mats, ctx->mc.tree_perm, &hperm));
        RC(read_root(ctx, ctx->mc.tree_perm.p, hperm, root_m));
        for (int i = 0; i < 8; i++) proof->perm_root[i] = h_from_m(root_m[i]);
        ch.observe(root_m, 8);
        for (size_t c = 0; c < K; c++) if (W[c].wp) ch.observe(W[c].cumsum.c, 4);
    }
    CK(cudaEventRecord(ctx->ev[2], st));
    const bb::E4 alpha = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->alpha[i] = h_from_m(alpha.c[i]);

    // ---- quotients: the same alpha for every chip, one commitment ----
    RC(ctx->ws_q.ensure(8 * n_max));
    mats.clear();
    for (size_t c = 0; c < K; c++) {
        McChip& w = W[c];
        const size_t N = (size_t)1 << w.log_n, M = N << 1, log_m = w.log_n + 1;
        RC(w.qnat.ensure(8 * N));
        RC(w.qlde.ensure(8 * M));
        if (!w.wp && w.air->n_constraints == 0) {
            CK(cudaMemsetAsync(ctx->ws_q.p, 0, 32 * N, st));          // an AIR without constraints or interactions: zero quotient
        } else if (!w.wp) {
            RC(pb_quotient(ctx, w.air, w.lde.p, w.log_n, 1, bb::GEN, proof->alpha, ctx->ws_q.p));
        } else {
            RC(ctx->ws_lu_raw.ensure(4 * M));
            RC(ctx->ws_lu_s.ensure(4 * M));
            RC(constraint_fold_m(ctx, w.air, w.lde.p, M, alpha, w.n_chunks + 3, ctx->ws_lu_raw.p));
            std::vector<bb::E4> apl(std::max<size_t>(1, w.n_chunks));
            bb::E4 cur = bb::e4_mul(alpha, alpha);
            for (size_t k = w.n_chunks; k-- > 0;) { cur = bb::e4_mul(cur, alpha); apl[k] = cur; }
            CK(cudaMemcpyAsync(w.air->d_apl, apl.data(), w.n_chunks * 16, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            RC(logup::launch_fold(w.air->lujit, st, w.lde.p, w.perm_lde.p, M, w.air->d_kc, w.air->d_bt, w.air->d_apl, ctx->ws_lu_raw.p, ctx->ws_lu_s.p));
            LAUNCHED(ctx);
            const uint32_t sn = bb::pow(h_to_m(bb::GEN), (uint64_t)1 << w.log_n);
            const uint32_t zinv0 = bb::inv(bb::sub(sn, bb::R1)), zinv1 = bb::inv(bb::sub(bb::neg(sn), bb::R1));
            logup::finish_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(ctx->ws_lu_raw.p, ctx->ws_lu_s.p, w.perm_lde.p + 4 * w.n_chunks * M, M, (int)w.log_n,
                                                                             h_to_m(bb::GEN), h_root_of_unity_m((int)log_m), bb::inv(h_root_of_unity_m((int)w.log_n)), sn,
                                                                             alpha, bb::e4_mul(alpha, alpha), w.cumsum, zinv0, zinv1, ctx->ws_q.p);
            LAUNCHED(ctx);
        }
        const size_t tot = 8 * N;
        ntt::bitrev_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ctx->ws_q.p, w.qnat.p, (int)w.log_n, 8);
        LAUNCHED(ctx);
        const uint32_t w2n_inv = h_from_m(bb::inv(h_root_of_unity_m((int)w.log_n + 1)));
        RC(pb_lde_batch(ctx, w.qnat.p, w.log_n, 4, 1, 1u, w.qlde.p));
        RC(pb_lde_batch(ctx, w.qnat.p + 4 * N, w.log_n, 4, 1, w2n_inv, w.qlde.p + 4 * M));
        mats.push_back(McMat{w.qlde.p, 8, log_m});
    }
    RC(mmcs_commit_dev(ctx, mats, ctx->mc.tree_q, &hq));
    RC(read_root(ctx, ctx->mc.tree_q.p, hq, root_m));
    for (int i = 0; i < 8; i++) proof->quotient_root[i] = h_from_m(root_m[i]);
    ch.observe(root_m, 8);
    const bb::E4 zeta = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->zeta[i] = h_from_m(zeta.c[i]);
    CK(cudaEventRecord(ctx->ev[3], st));

    // ---- openings in observation order ----
    size_t n_open = 0;
    for (size_t c = 0; c < K; c++) { W[c].e_main = n_open; n_open += W[c].width; }
    for (size_t c = 0; c < K; c++) if (W[c].wp) { W[c].e_perm = n_open; n_open += 2 * W[c].wp; }
    for (size_t c = 0; c < K; c++) { W[c].e_q = n_open; n_open += 8; }
    RC(ctx->ws_ys.ensure(4 * n_open));
    for (size_t c = 0; c < K; c++) {
        McChip& w = W[c];
        const size_t N = (size_t)1 << w.log_n;
        const bb::E4 zeta_next = bb::e4_scale(zeta, h_root_of_unity_m((int)w.log_n));
        const uint32_t g_c = bb::GEN, gw_c = h_from_m(bb::mul(h_to_m(bb::GEN), h_root_of_unity_m((int)w.log_n + 1)));
        RC(eval_at_point_m(ctx, w.d_trace, w.log_n, w.width, h_to_m(1u), zeta, ctx->ws_ys.p + 4 * w.e_main));
        if (w.wp) {
            RC(eval_at_point_m(ctx, w.perm.p, w.log_n, w.wp, h_to_m(1u), zeta, ctx->ws_ys.p + 4 * w.e_perm));
            RC(eval_at_point_m(ctx, w.perm.p, w.log_n, w.wp, h_to_m(1u), zeta_next, ctx->ws_ys.p + 4 * (w.e_perm + w.wp)));
        }
        RC(eval_at_point_m(ctx, w.qnat.p, w.log_n, 4, h_to_m(g_c), zeta, ctx->ws_ys.p + 4 * w.e_q));
        RC(eval_at_point_m(ctx, w.qnat.p + 4 * N, w.log_n, 4, h_to_m(gw_c), zeta, ctx->ws_ys.p + 4 * (w.e_q + 4)));
    }
    std::vector<uint32_t>& ys_h = ctx->mc.ys;
    ys_h.assign(4 * n_open, 0u);
    CK(cudaMemcpyAsync(ys_h.data(), ctx->ws_ys.p, 16 * n_open, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ch.observe(ys_h.data(), (int)(4 * n_open));
    const bb::E4 gamma = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->gamma[i] = h_from_m(gamma.c[i]);

    // ---- one reduced-opening codeword per LDE height ----
    bool have_h[32] = {false};
    for (size_t c = 0; c < K; c++) {
        const size_t lm = W[c].log_n + 1;
        if (!have_h[lm]) {
            RC(ctx->mc.ro[lm].ensure((size_t)4 << lm));
            CK(cudaMemsetAsync(ctx->mc.ro[lm].p, 0, (size_t)16 << lm, st));
            have_h[lm] = true;
        }
    }
    for (size_t c = 0; c < K; c++) {
        McChip& w = W[c];
        const size_t lm = w.log_n + 1, M = (size_t)1 << lm;
        const bb::E4 zeta_next = bb::e4_scale(zeta, h_root_of_unity_m((int)w.log_n));
        for (int part = 0; part < 3; part++) {
            const size_t wdt = part == 0 ? w.width : part == 1 ? 2 * w.wp : 8;
            if (!wdt) continue;
            const size_t e0 = part == 0 ? w.e_main : part == 1 ? w.e_perm : w.e_q;
            std::vector<const uint32_t*> cols(wdt);
            std::vector<uint32_t> grp(wdt, 0u);
            for (size_t j = 0; j < wdt; j++) {
                if (part == 0) cols[j] = w.lde.p + j * M;
                else if (part == 1) { cols[j] = w.perm_lde.p + (j % w.wp) * M; grp[j] = j >= w.wp; }
                else cols[j] = w.qlde.p + j * M;
            }
            std::vector<bb::E4> zs{zeta};
            if (part == 1) zs.push_back(zeta_next);
            bb::E4 gs = {{bb::R1, 0u, 0u, 0u}}, base = gamma;          // gamma^e0 by square and multiply
            for (size_t e = e0; e; e >>= 1) { if (e & 1) gs = bb::e4_mul(gs, base); base = bb::e4_mul(base, base); }
            RC(deep_quotient_groups_m(ctx, cols, grp, zs, lm, h_to_m(bb::GEN), gamma, ys_h.data() + 4 * e0, ctx->mc.ro[lm].p, 0, 0, &gs, true));
        }
    }
    CK(cudaEventRecord(ctx->ev[4], st));

    // ---- FRI: fold from the tallest codeword; the codeword of a height joins when the fold reaches it ----
    const size_t Mmax = (size_t)1 << hmax;
    RC(ctx->ws_fri_words.ensure(8 * Mmax + 64));
    RC(ctx->ws_fri_trees.ensure(8 * (2 * Mmax)));
    CK(cudaMemcpyAsync(ctx->ws_fri_words.p, ctx->mc.ro[hmax].p, (size_t)16 << hmax, cudaMemcpyDeviceToDevice, st));
    uint32_t* inject[32] = {nullptr};
    for (size_t h = 2; h < hmax; h++) if (have_h[h]) inject[h] = ctx->mc.ro[h].p;
    uint32_t layer = 0;
    uint32_t fin[8];
    RC(fri_commit_phase(ctx, ch, hmax, inject, ctx->mc.word_off, ctx->mc.tree_off, &layer, proof->fri_roots, proof->fri_betas, fin, nullptr));
    proof->n_fri_layers = layer;
    proof->final_len = 2;
    for (uint32_t i = 0; i < 2; i++)
        for (int l = 0; l < 4; l++) proof->final_poly[i][l] = h_from_m(fin[4 * i + l]);
    ch.observe(fin, 4);
    {
        uint32_t wv = 0;
        RC(grind(ctx, ch, ctx->pow_bits, &wv));
        proof->pow_witness = wv;
        const uint32_t w_m = h_to_m(wv);
        ch.observe(&w_m, 1);
        (void)ch.sample();
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaStreamSynchronize(st));
    ctx->mc.valid = true;
    ctx->mc.K = K; ctx->mc.hmax = hmax; ctx->mc.hperm = hperm; ctx->mc.n_layers = layer; ctx->mc.any_lu = any_lu;
    ctx->mc.ch = ch;
    memset(ctx->stage_ms, 0, sizeof ctx->stage_ms);
    // stage clocks of the multi-chip proof in the layout of pb_last_stage_ms: main commit -> [1]+[2] (lde slot), LogUp -> logup_commit slot,
    // quotient phase -> quotient slot, openings -> open, FRI + PoW -> fri
    cudaEventElapsedTime(&ctx->stage_ms[1], ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stage_ms[4], ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->stage_ms[5], ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->stage_ms[8], ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->stage_ms[9], ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&ctx->stage_ms[11], ctx->ev[0], ctx->ev[5]);
#undef RC
    return 0;
}

// query openings + opened values of the last pb_prove_chips (layout: include/powdr_b200.h)
int pb_query_chips(pb_ctx_t* ctx, uint32_t* h_queries, size_t cap_words, uint32_t* h_ys, size_t cap_ys_words) {
    if (!ctx || !ctx->mc.valid) return PB_ERR_INVALID_ARG;
    if (h_ys) {
        if (cap_ys_words < ctx->mc.ys.size()) return PB_ERR_INVALID_ARG;
        for (size_t i = 0; i < ctx->mc.ys.size(); i++) h_ys[i] = h_from_m(ctx->mc.ys[i]);
    }
    const size_t nq = ctx->n_queries, K = ctx->mc.K, hmax = ctx->mc.hmax, hperm = ctx->mc.hperm;
    if (!nq || !h_queries) return 0;
    std::vector<McChip>& W = ctx->mc.chips;
    size_t wm = 0, wp = 0;
    for (size_t c = 0; c < K; c++) { wm += W[c].width; wp += W[c].wp; }
    size_t wpq = 1 + wm + 8 * hmax + (wp ? wp + 8 * hperm : 0) + 8 * K + 8 * hmax;
    for (size_t i = 0; i + 1 < hmax; i++) wpq += 8 + 8 * (hmax - 1 - i);
    if (cap_words < wpq * nq) return PB_ERR_INVALID_ARG;
    int rc;
    std::vector<uint32_t> idx(nq);
    Challenger ch = ctx->mc.ch;
    for (size_t q = 0; q < nq; q++) idx[q] = h_from_m(ch.sample()) & (uint32_t)(((size_t)1 << hmax) - 1);
    if ((rc = ctx->ws_qidx.ensure(nq))) return rc;
    if ((rc = ctx->ws_qout.ensure(wpq * nq))) return rc;
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->ws_qidx.p, idx.data(), 4 * nq, cudaMemcpyHostToDevice, st));
    const unsigned G = (unsigned)nq;
    uint32_t* out = ctx->ws_qout.p;
    fri::gather_index_kernel<<<G, 32, 0, st>>>(ctx->ws_qidx.p, out, wpq);
    size_t off = 1;
    for (size_t c = 0; c < K; c++) {
        const size_t lm = W[c].log_n + 1;
        fri::gather_rows_kernel<<<G, 128, 0, st>>>(W[c].lde.p, (size_t)1 << lm, (uint32_t)W[c].width, (int)(hmax - lm), ctx->ws_qidx.p, out, wpq, off);
        off += W[c].width;
    }
    fri::gather_path_kernel<<<G, 32, 0, st>>>(ctx->mc.tree_main.p, (int)hmax, 0, ctx->ws_qidx.p, out, wpq, off);
    off += 8 * hmax;
    if (wp) {
        for (size_t c = 0; c < K; c++) {
            if (!W[c].wp) continue;
            const size_t lm = W[c].log_n + 1;
            fri::gather_rows_kernel<<<G, 128, 0, st>>>(W[c].perm_lde.p, (size_t)1 << lm, (uint32_t)W[c].wp, (int)(hmax - lm), ctx->ws_qidx.p, out, wpq, off);
            off += W[c].wp;
        }
        fri::gather_path_kernel<<<G, 32, 0, st>>>(ctx->mc.tree_perm.p, (int)hperm, (int)(hmax - hperm), ctx->ws_qidx.p, out, wpq, off);
        off += 8 * hperm;
    }
    for (size_t c = 0; c < K; c++) {
        const size_t lm = W[c].log_n + 1;
        fri::gather_rows_kernel<<<G, 32, 0, st>>>(W[c].qlde.p, (size_t)1 << lm, 8u, (int)(hmax - lm), ctx->ws_qidx.p, out, wpq, off);
        off += 8;
    }
    fri::gather_path_kernel<<<G, 32, 0, st>>>(ctx->mc.tree_q.p, (int)hmax, 0, ctx->ws_qidx.p, out, wpq, off);
    off += 8 * hmax;
    fri::FriDesc d;
    d.fri_words = ctx->ws_fri_words.p; d.fri_trees = ctx->ws_fri_trees.p; d.log_m = (int)hmax; d.n_layers = (int)ctx->mc.n_layers;
    for (int i = 0; i < 32; i++) { d.word_off[i] = ctx->mc.word_off[i]; d.tree_off[i] = ctx->mc.tree_off[i]; }
    fri::gather_fri_kernel<<<G, 32, 0, st>>>(d, ctx->ws_qidx.p, out, wpq, off);
    ctx->launches += 4 + 3 * K;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_queries, out, 4 * wpq * nq, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}
