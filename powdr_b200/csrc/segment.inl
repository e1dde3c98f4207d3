// Whole-segment prover, transcript v2 (restated on the CPU by the test oracle and checked by its independent verifier).
// Included by capi.cu inside its extern "C" block (uses its file-local helpers).
//
//   main commit -> [LogUp: alpha_lu, beta_lu; permutation trace generated, LDE'd, committed; cumulative sum observed] -> alpha ->
//   quotient (AIR constraints, then LogUp constraints, one alpha-fold) -> quotient commit -> zeta -> openings at zeta / zeta*w, every
//   value observed -> gamma -> reduced opening -> FRI commit phase -> final polynomial observed -> proof of work -> queries.

// ---------------------------------------------------------------------------------------------------------------------
int pb_ctx_set_fri_params(pb_ctx_t* ctx, uint32_t n_queries, uint32_t pow_bits) {
    if (!ctx || pow_bits > 30 || n_queries > (1u << 16)) return PB_ERR_INVALID_ARG;
    ctx->n_queries = n_queries;
    ctx->pow_bits = pow_bits;
    return 0;
}

int pb_air_perm_width(const pb_air_t* a, size_t* perm_width) {
    if (!a || !perm_width) return PB_ERR_INVALID_ARG;
    *perm_width = a->has_lu ? a->lu.perm_width() : 0;
    return 0;
}

namespace {
// validates and packs the interactions, chunks them; on success `p`, `code`, `spans`, `pool` describe the LogUp program
int prepare_logup(const uint32_t* bc, size_t n_words, const ExprSpan* arg_spans, size_t n_arg_spans, const DevInteraction* ints, size_t n_ints,
                  uint32_t width, logup::Program& p, std::vector<uint32_t>& code, std::vector<uint32_t>& pool, std::vector<air::Span>& spans) {
    // every span of every interaction becomes one packed expression (same validation as the constraints)
    std::vector<pb_expr_span_t> sp(n_arg_spans);
    for (size_t i = 0; i < n_arg_spans; i++) { sp[i].off = arg_spans[i].off; sp[i].len = arg_spans[i].len; }
    int prc = pack_program(bc, n_words, sp.data(), n_arg_spans, width, code, pool, spans);
    if (prc) return prc;
    for (size_t i = 0; i < n_ints; i++) {
        if ((size_t)ints[i].args_index_off + ints[i].num_args + 1 > n_arg_spans) return PB_ERR_BAD_BYTECODE;
        p.ints.push_back(logup::Interaction{ints[i].bus_id, ints[i].num_args, ints[i].args_index_off});
        p.max_args = std::max<size_t>(p.max_args, ints[i].num_args);
    }
    if (!logup::make_chunks(code, spans, p)) return PB_ERR_UNSUPPORTED;        // an interaction above the degree bound
    return 0;
}
}  // namespace

int pb_air_logup_compile_only(const uint32_t* bc, size_t n_words, const ExprSpan* arg_spans, size_t n_arg_spans, const DevInteraction* ints,
                              size_t n_ints, uint32_t width, size_t* cubin_bytes, size_t* perm_width) {
    if ((!bc && n_words) || (!arg_spans && n_arg_spans) || (!ints && n_ints)) return PB_ERR_INVALID_ARG;
    logup::Program p;
    std::vector<uint32_t> code, pool;
    std::vector<air::Span> spans;
    int rc = prepare_logup(bc, n_words, arg_spans, n_arg_spans, ints, n_ints, width, p, code, pool, spans);
    if (rc) return rc;
    if (perm_width) *perm_width = p.perm_width();
    size_t total = 0;
    if (n_ints && logup::build(code, spans, pool, p, nullptr, &total)) return PB_ERR_UNSUPPORTED;
    if (cubin_bytes) *cubin_bytes = total;
    return 0;
}

int pb_air_set_interactions(pb_ctx_t* ctx, pb_air_t* a, const uint32_t* bc, size_t n_words, const ExprSpan* arg_spans, size_t n_arg_spans,
                            const DevInteraction* ints, size_t n_ints) {
    if (!ctx || !a || (!bc && n_words) || (!arg_spans && n_arg_spans) || (!ints && n_ints)) return PB_ERR_INVALID_ARG;
    if (a->has_lu) return PB_ERR_INVALID_ARG;                 // set once
    if (n_ints == 0) return 0;
    logup::Program p;
    std::vector<uint32_t> code, pool;
    std::vector<air::Span> spans;
    int rc = prepare_logup(bc, n_words, arg_spans, n_arg_spans, ints, n_ints, a->width, p, code, pool, spans);
    if (rc) return rc;
    if (logup::build(code, spans, pool, p, &a->lujit)) return PB_ERR_UNSUPPORTED;   // NVRTC / driver missing: fail loudly, no fallback
    CK(cudaMalloc((void**)&a->d_kc, n_ints * sizeof(uint4)));
    CK(cudaMalloc((void**)&a->d_bt, (p.max_args + 1) * sizeof(uint4)));
    CK(cudaMalloc((void**)&a->d_apl, std::max<size_t>(1, p.n_chunks()) * sizeof(uint4)));
    a->lu = std::move(p);
    a->has_lu = true;
    return 0;
}

int pb_allgather_caps(pb_ctx_t* ctx, const pb_comm_t* comm, const uint32_t* d_local_cap, uint32_t* d_all_caps) {
    if (!ctx || !comm || !comm->all_gather || !d_local_cap || !d_all_caps) return PB_ERR_INVALID_ARG;
    if (!(comm->flags & PB_COMM_STREAM_ORDERED)) CK(cudaStreamSynchronize(ctx->stream));
    return comm->all_gather(comm->user, d_local_cap, d_all_caps, 32) ? PB_ERR_COMM : 0;
}

namespace {

// proof-of-work: smallest witness w such that observe(w); sample_bits(bits) == 0 on a copy of the challenger
int grind(pb_ctx* ctx, const Challenger& ch, uint32_t bits, uint32_t* witness) {
    p2::GrindState gs;
    memcpy(gs.s, ch.sponge, sizeof gs.s);
    for (int i = 0; i < ch.n_in; i++) gs.s[i] = ch.in_buf[i];
    const int pos = ch.n_in;                                        // n_in < 8 always (a full buffer is absorbed at once)
    const uint32_t mask = (1u << bits) - 1;
    int rc = ctx->ws_pow.ensure(1);
    if (rc) return rc;
    const uint32_t batch = 1u << 20;
    for (uint64_t base = 0; base < bb::P; base += batch) {
        CK(cudaMemsetAsync(ctx->ws_pow.p, 0xff, 4, ctx->stream));
        const uint32_t count = (uint32_t)std::min<uint64_t>(batch, bb::P - base);
        p2::grind_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(gs, pos, mask, (uint32_t)base, count, ctx->ws_pow.p);
        LAUNCHED(ctx);
        uint32_t found;
        CK(cudaMemcpyAsync(&found, ctx->ws_pow.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (found != 0xffffffffu) { *witness = found; return 0; }
    }
    return PB_ERR_UNSUPPORTED;
}

// per-proof LogUp constants: kc_i = alpha_lu + beta^k (bus+1) + sum over literal args beta^j c_j;  bt_j = beta^j
int upload_logup_consts(pb_ctx* ctx, const pb_air* a, bb::E4 al, bb::E4 be) {
    const logup::Program& p = a->lu;
    std::vector<bb::E4> bt(p.max_args + 1);
    bt[0] = bb::E4{{bb::R1, 0u, 0u, 0u}};
    for (size_t j = 1; j <= p.max_args; j++) bt[j] = bb::e4_mul(bt[j - 1], be);
    std::vector<bb::E4> kc(p.ints.size());
    for (size_t i = 0; i < p.ints.size(); i++)
        kc[i] = bb::e4_add(al, bb::e4_scale(bt[p.ints[i].num_args], h_to_m((p.ints[i].bus_id + 1) % bb::P)));
    for (const logup::LitArg& l : p.lits) kc[l.interaction] = bb::e4_add(kc[l.interaction], bb::e4_scale(bt[l.j], l.value_m));
    CK(cudaMemcpyAsync(a->d_kc, kc.data(), kc.size() * 16, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(a->d_bt, bt.data(), bt.size() * 16, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));      // host temporaries
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
int pb_prove_segment(pb_ctx_t* ctx, const pb_air_t* a, const uint32_t* trace, size_t log_n, size_t width, uint32_t flags,
                     pb_segment_proof_t* proof) {
    if (!ctx || !a || !trace || !proof) return PB_ERR_INVALID_ARG;
    if (log_n < 1 || log_n > 24 || width == 0 || width != a->width) return PB_ERR_INVALID_ARG;
    const uint32_t log_blowup = 1;
    const size_t N = (size_t)1 << log_n, M = N << log_blowup;
    const size_t log_m = log_n + log_blowup;
    const size_t n_chunks = a->has_lu ? a->lu.n_chunks() : 0, wp = a->has_lu ? a->lu.perm_width() : 0;
    int rc;
    memset(proof, 0, sizeof *proof);
    ctx->seg.valid = false;                           // a failure below must not leave a stale proof state queryable
    ctx->sh.valid = false;                            // the sharded and the multi-chip prover share workspaces with this one
    ctx->mc.valid = false;
    proof->pow_bits = ctx->pow_bits;
    proof->n_queries = ctx->n_queries;
    proof->perm_width = (uint32_t)wp;
    cudaStream_t st = ctx->stream;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
    CK(cudaEventRecord(ctx->ev[0], st));
    RC(ctx->ws_lde.ensure(width * M));
    RC(ctx->ws_layers.ensure(8 * (2 * M)));
    RC(ctx->ws_q.ensure(8 * N));
    RC(ctx->ws_qnat.ensure(8 * N));
    RC(ctx->ws_qlde.ensure(8 * M));
    RC(ctx->ws_f0.ensure(4 * M));
    RC(ctx->ws_f1.ensure(4 * (M / 2)));
    RC(ctx->ws_layers_q.ensure(8 * (2 * M)));
    RC(ctx->ws_fri_words.ensure(8 * M + 64));
    RC(ctx->ws_fri_trees.ensure(8 * (2 * M)));
    if (wp) {
        RC(ctx->ws_perm.ensure(wp * N));
        RC(ctx->ws_perm_lde.ensure(wp * M));
        RC(ctx->ws_layers_p.ensure(8 * (2 * M)));
        RC(ctx->ws_rowsum.ensure(4 * N));
        RC(ctx->ws_lu_raw.ensure(4 * M));
        RC(ctx->ws_lu_s.ensure(4 * M));
    }
    Challenger ch;
    ch.k = &ctx->p2;
    uint32_t root_m[8];
    const uint32_t* d_trace_full = trace;        // device-resident trace (the caller's buffer, or ws_trace in host mode)

    if (flags & PB_TRACE_ON_DEVICE) {
        // main trace commit: LDE then Merkle
        CK(cudaEventRecord(ctx->ev[1], st));
        RC(pb_lde_batch(ctx, trace, log_n, width, log_blowup, bb::GEN, ctx->ws_lde.p));
        CK(cudaEventRecord(ctx->ev[2], st));
        const uint32_t* mats1[1] = {ctx->ws_lde.p};
        RC(pb_merkle_commit(ctx, mats1, &width, 1, log_m, ctx->ws_layers.p, nullptr));
        CK(cudaEventRecord(ctx->ev[3], st));
    } else {
        // Host trace: software pipeline over column chunks.  The PCIe copy of chunk k+1 (copy stream, double-buffered
        // staging) overlaps the LDE and the sponge absorption of chunk k (compute stream); per-row sponge states live in
        // HBM between chunks.  Stage clocks in this mode: [1]->[2] = copy+LDE+leaf hashing overlapped, [2]->[3] = upper layers.
        size_t cw = std::max<size_t>(8, ((((size_t)256 << 20) / (4 * N)) / 8) * 8);     // ~256 MB per chunk, multiple of the sponge rate
        if (const char* e = getenv("PB_PIPE_CHUNK_COLS")) cw = std::max<size_t>(8, ((size_t)atol(e) / 8) * 8);
        cw = std::min<size_t>(width, cw);
        // The pipeline is PCIe-bound in steady state (8.5 GB at ~50 GB/s = 171 ms vs 163 ms of LDE + hashing), so what is exposed is
        // the first copy (nothing to hide behind) and the compute of the last chunk (no copy left to hide it): ramp the chunk
        // width up from 8 columns at the start and down to 8 at the end.  Every chunk but the last is a multiple of the sponge rate.
        std::vector<size_t> chunk_c0, chunk_w;
        {
            std::vector<size_t> ws;
            const size_t r8 = width % 8;
            if (width >= 4 * cw + 112) {
                for (size_t w0 : {8, 16, 32}) ws.push_back(w0);
                size_t mid = width - r8 - 112;
                while (mid > 0) { const size_t w0 = std::min(cw, mid); ws.push_back(w0); mid -= w0; }
                ws.push_back(32); ws.push_back(16); ws.push_back(8 + r8);
            } else {
                for (size_t c0 = 0, wk = 8; c0 < width; c0 += ws.back(), wk = std::min(cw, 2 * wk)) ws.push_back(std::min(wk, width - c0));
            }
            size_t c0 = 0;
            for (size_t w0 : ws) { chunk_c0.push_back(c0); chunk_w.push_back(w0); c0 += w0; }
        }
        const size_t n_chunks_h = chunk_w.size();
        RC(ctx->ws_trace.ensure(width * N));          // whole trace stays resident: it is read again by LogUp and the openings
        d_trace_full = ctx->ws_trace.p;
        RC(ctx->ws_state.ensure(16 * M));
        std::vector<const uint32_t*> cols(width);
        for (size_t c = 0; c < width; c++) cols[c] = ctx->ws_lde.p + c * M;
        RC(ctx->coltab.ensure(width));
        CK(cudaMemcpyAsync(ctx->coltab.p, cols.data(), width * sizeof(void*), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));                  // cols is a host temporary
        CK(cudaEventRecord(ctx->ev[1], st));
        CK(cudaEventRecord(ctx->ev_free[0], st));          // orders the first copies after everything already queued on st
        CK(cudaEventRecord(ctx->ev_free[1], st));
        for (size_t k = 0; k < n_chunks_h; k++) {
            const int b = (int)(k & 1);
            const size_t c0 = chunk_c0[k], wk = chunk_w[k];
            uint32_t* stage = ctx->ws_trace.p + c0 * N;
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[b], 0));
            CK(cudaMemcpyAsync(stage, trace + c0 * N, wk * N * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
            CK(cudaStreamWaitEvent(st, ctx->ev_copy[b], 0));
            RC(pb_lde_batch(ctx, stage, log_n, wk, log_blowup, bb::GEN, ctx->ws_lde.p + c0 * M));
            CK(cudaEventRecord(ctx->ev_free[b], st));
            p2::leaf_absorb_cols_kernel<<<(unsigned)((M + p2::LEAF_THREADS - 1) / p2::LEAF_THREADS), p2::LEAF_THREADS, 0, st>>>(ctx->coltab.p + c0, (uint32_t)wk, M, ctx->ws_state.p,
                                                                                    ctx->ws_layers.p, k == 0, k + 1 == n_chunks_h);
            LAUNCHED(ctx);
        }
        CK(cudaEventRecord(ctx->ev[2], st));
        RC(merkle_upper(ctx, ctx->ws_layers.p, log_m));
        CK(cudaEventRecord(ctx->ev[3], st));
    }
    RC(read_root(ctx, ctx->ws_layers.p, log_m, root_m));
    for (int i = 0; i < 8; i++) proof->trace_root[i] = h_from_m(root_m[i]);
    ch.observe(root_m, 8);

    // ---- LogUp phase: permutation trace from the bus interactions, running sum, commitment ----
    bb::E4 cumsum = {{0u, 0u, 0u, 0u}};
    if (wp) {
        const bb::E4 al = ch.sample_ext(), be = ch.sample_ext();
        for (int i = 0; i < 4; i++) { proof->logup_alpha[i] = h_from_m(al.c[i]); proof->logup_beta[i] = h_from_m(be.c[i]); }
        RC(upload_logup_consts(ctx, a, al, be));
        RC(logup::launch_perm(a->lujit, st, d_trace_full, N, a->d_kc, a->d_bt, ctx->ws_perm.p, ctx->ws_rowsum.p));
        LAUNCHED(ctx);
        // phi = inclusive prefix sums of the row sums, written as the last 4 columns of the permutation trace
        {
            const unsigned nb = (unsigned)((N + logup::SCAN_THREADS * logup::SCAN_ITEMS - 1) / (logup::SCAN_THREADS * logup::SCAN_ITEMS));
            RC(ctx->ws_scan_tot.ensure(4 * (size_t)nb));
            uint32_t* phi = ctx->ws_perm.p + 4 * n_chunks * N;
            logup::scan_local_kernel<<<dim3(nb, 4), logup::SCAN_THREADS, 0, st>>>(ctx->ws_rowsum.p, phi, N, ctx->ws_scan_tot.p);
            logup::scan_totals_kernel<<<1, 32, 0, st>>>(ctx->ws_scan_tot.p, nb);
            logup::scan_add_kernel<<<dim3((unsigned)((N + 255) / 256), 4), 256, 0, st>>>(phi, N, ctx->ws_scan_tot.p, nb);
            ctx->launches += 3;
            uint32_t cs[4];
            for (int l = 0; l < 4; l++) CK(cudaMemcpyAsync(&cs[l], phi + (size_t)l * N + (N - 1), 4, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ctx->ev[4], st));
            CK(cudaStreamSynchronize(st));
            for (int l = 0; l < 4; l++) { cumsum.c[l] = cs[l]; proof->cumulative_sum[l] = h_from_m(cs[l]); }
        }
        RC(pb_lde_batch(ctx, ctx->ws_perm.p, log_n, wp, log_blowup, bb::GEN, ctx->ws_perm_lde.p));
        const uint32_t* matsp[1] = {ctx->ws_perm_lde.p};
        RC(pb_merkle_commit(ctx, matsp, &wp, 1, log_m, ctx->ws_layers_p.p, nullptr));
        CK(cudaEventRecord(ctx->ev[5], st));
        RC(read_root(ctx, ctx->ws_layers_p.p, log_m, root_m));
        for (int i = 0; i < 8; i++) proof->perm_root[i] = h_from_m(root_m[i]);
        ch.observe(root_m, 8);
        ch.observe(cumsum.c, 4);
    } else {
        CK(cudaEventRecord(ctx->ev[4], st));
        CK(cudaEventRecord(ctx->ev[5], st));
    }
    const bb::E4 alpha = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->alpha[i] = h_from_m(alpha.c[i]);

    // ---- quotient ----
    if (!wp) {
        RC(pb_quotient(ctx, a, ctx->ws_lde.p, log_n, log_blowup, bb::GEN, proof->alpha, ctx->ws_q.p));
    } else {
        // AIR constraints folded with alpha^(K-1-k), K = C + n_chunks + 3; then the LogUp chunk constraints; then the phi constraints,
        // the division by Z_H and the chunk split
        RC(constraint_fold_m(ctx, a, ctx->ws_lde.p, M, alpha, n_chunks + 3, ctx->ws_lu_raw.p));
        {
            std::vector<bb::E4> apl(std::max<size_t>(1, n_chunks));
            bb::E4 cur = bb::e4_mul(alpha, alpha);                  // alpha^2 belongs to the first-row constraint
            for (size_t c = n_chunks; c-- > 0;) { cur = bb::e4_mul(cur, alpha); apl[c] = cur; }      // apl[c] = alpha^(n_chunks + 2 - c)
            CK(cudaMemcpyAsync(a->d_apl, apl.data(), n_chunks * 16, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
        RC(logup::launch_fold(a->lujit, st, ctx->ws_lde.p, ctx->ws_perm_lde.p, M, a->d_kc, a->d_bt, a->d_apl, ctx->ws_lu_raw.p, ctx->ws_lu_s.p));
        LAUNCHED(ctx);
        const uint32_t sn = bb::pow(h_to_m(bb::GEN), (uint64_t)1 << log_n);
        const uint32_t zinv0 = bb::inv(bb::sub(sn, bb::R1)), zinv1 = bb::inv(bb::sub(bb::neg(sn), bb::R1));
        logup::finish_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(ctx->ws_lu_raw.p, ctx->ws_lu_s.p, ctx->ws_perm_lde.p + 4 * n_chunks * M, M, (int)log_n,
                                                                         h_to_m(bb::GEN), h_root_of_unity_m((int)log_m), bb::inv(h_root_of_unity_m((int)log_n)),
                                                                         sn, alpha, bb::e4_mul(alpha, alpha), cumsum, zinv0, zinv1, ctx->ws_q.p);
        LAUNCHED(ctx);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[6], st));

    // quotient commit: chunk b = evals over g*w_{2N}^b*H (bit-reversed) -> natural -> LDE with shift g/s_b = w_{2N}^-b
    {
        const size_t tot = 8 * N;
        ntt::bitrev_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ctx->ws_q.p, ctx->ws_qnat.p, (int)log_n, 8);
        LAUNCHED(ctx);
        const uint32_t w2n_inv = h_from_m(bb::inv(h_root_of_unity_m((int)log_n + 1)));
        RC(pb_lde_batch(ctx, ctx->ws_qnat.p, log_n, 4, log_blowup, 1u, ctx->ws_qlde.p));
        RC(pb_lde_batch(ctx, ctx->ws_qnat.p + 4 * N, log_n, 4, log_blowup, w2n_inv, ctx->ws_qlde.p + 4 * M));
    }
    CK(cudaEventRecord(ctx->ev[7], st));
    const uint32_t* mats2[2] = {ctx->ws_qlde.p, ctx->ws_qlde.p + 4 * M};
    const size_t w2[2] = {4, 4};
    RC(pb_merkle_commit(ctx, mats2, w2, 2, log_m, ctx->ws_layers_q.p, nullptr));
    CK(cudaEventRecord(ctx->ev[8], st));
    RC(read_root(ctx, ctx->ws_layers_q.p, log_m, root_m));
    for (int i = 0; i < 8; i++) proof->quotient_root[i] = h_from_m(root_m[i]);
    ch.observe(root_m, 8);
    const bb::E4 zeta = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->zeta[i] = h_from_m(zeta.c[i]);
    const bb::E4 zeta_next = bb::e4_scale(zeta, h_root_of_unity_m((int)log_n));

    // ---- openings: main at zeta, perm at zeta and zeta*w, quotient chunk b over g*w_{2N}^b*H at zeta; every value is observed ----
    const size_t n_open = width + 2 * wp + 8;
    RC(ctx->ws_ys.ensure(4 * n_open));
    {
        const uint32_t g_c = bb::GEN, gw_c = h_from_m(bb::mul(h_to_m(bb::GEN), h_root_of_unity_m((int)log_n + 1)));
        RC(eval_at_point_m(ctx, d_trace_full, log_n, width, h_to_m(1u), zeta, ctx->ws_ys.p));
        if (wp) {
            RC(eval_at_point_m(ctx, ctx->ws_perm.p, log_n, wp, h_to_m(1u), zeta, ctx->ws_ys.p + 4 * width));
            RC(eval_at_point_m(ctx, ctx->ws_perm.p, log_n, wp, h_to_m(1u), zeta_next, ctx->ws_ys.p + 4 * (width + wp)));
        }
        RC(eval_at_point_m(ctx, ctx->ws_qnat.p, log_n, 4, h_to_m(g_c), zeta, ctx->ws_ys.p + 4 * (width + 2 * wp)));
        RC(eval_at_point_m(ctx, ctx->ws_qnat.p + 4 * N, log_n, 4, h_to_m(gw_c), zeta, ctx->ws_ys.p + 4 * (width + 2 * wp + 4)));
    }
    std::vector<uint32_t>& ys_h = ctx->seg.ys;
    ys_h.assign(4 * n_open, 0u);
    CK(cudaMemcpyAsync(ys_h.data(), ctx->ws_ys.p, 16 * n_open, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ch.observe(ys_h.data(), (int)(4 * n_open));
    const bb::E4 gamma = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->gamma[i] = h_from_m(gamma.c[i]);

    // ---- reduced opening over g*H': the FRI input codeword ----
    {
        std::vector<const uint32_t*> cols(n_open);
        std::vector<uint32_t> grp(n_open, 0u);
        size_t k = 0;
        for (size_t c = 0; c < width; c++) cols[k++] = ctx->ws_lde.p + c * M;
        for (size_t c = 0; c < wp; c++) cols[k++] = ctx->ws_perm_lde.p + c * M;
        for (size_t c = 0; c < wp; c++) { grp[k] = 1; cols[k++] = ctx->ws_perm_lde.p + c * M; }
        for (size_t c = 0; c < 8; c++) cols[k++] = ctx->ws_qlde.p + c * M;
        std::vector<bb::E4> zs{zeta};
        if (wp) zs.push_back(zeta_next);
        // every FRI codeword and every layer tree stays resident (back to back) for the query phase
        RC(deep_quotient_groups_m(ctx, cols, grp, zs, log_m, h_to_m(bb::GEN), gamma, ys_h.data(), ctx->ws_fri_words.p));
    }
    CK(cudaEventRecord(ctx->ev[9], st));

    // ---- FRI commit phase (device-driven, one read-back at the end) ----
    uint32_t layer = 0;
    uint32_t fin[8];
    RC(fri_commit_phase(ctx, ch, log_m, nullptr, ctx->seg.word_off, ctx->seg.tree_off, &layer, proof->fri_roots, proof->fri_betas, fin, ctx->ev[10]));
    proof->n_fri_layers = layer;
    proof->final_len = 1u << log_blowup;
    for (uint32_t i = 0; i < proof->final_len; i++)
        for (int l = 0; l < 4; l++) proof->final_poly[i][l] = h_from_m(fin[4 * i + l]);
    ch.observe(fin, 4);                              // the final polynomial is one constant

    // ---- proof of work ----
    {
        uint32_t w = 0;
        RC(grind(ctx, ch, ctx->pow_bits, &w));
        proof->pow_witness = w;
        const uint32_t w_m = h_to_m(w);
        ch.observe(&w_m, 1);
        (void)ch.sample();
    }
    CK(cudaEventRecord(ctx->ev[11], st));
    CK(cudaStreamSynchronize(st));
    ctx->seg.valid = true;
    ctx->seg.log_n = log_n; ctx->seg.log_m = log_m; ctx->seg.width = width; ctx->seg.perm_width = wp; ctx->seg.n_layers = layer;
    ctx->seg.ch = ch;
    for (int i = 0; i < 10; i++) cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
    cudaEventElapsedTime(&ctx->stage_ms[10], ctx->ev[10], ctx->ev[11]);
    cudaEventElapsedTime(&ctx->stage_ms[11], ctx->ev[0], ctx->ev[11]);
#undef RC
    return 0;
}

// query phase of the last pb_prove_segment: indices from the transcript (sample_bits(log_m) each), openings gathered on the device
static size_t query_words(size_t log_n, size_t width, size_t wp) {
    const size_t log_m = log_n + 1, layers = log_n;      // log_blowup 1, final_poly_len 1
    size_t w = 1 + width + 8 * log_m + (wp ? wp + 8 * log_m : 0) + 8 + 8 * log_m;
    for (size_t i = 0; i < layers; i++) w += 8 + 8 * (log_m - 1 - i);
    return w;
}

int pb_query_words(size_t log_n, size_t width, size_t perm_width, size_t* words_per_query) {
    if (!words_per_query || log_n < 1 || log_n > 24) return PB_ERR_INVALID_ARG;
    *words_per_query = query_words(log_n, width, perm_width);
    return 0;
}

int pb_query_segment(pb_ctx_t* ctx, uint32_t* h_out, size_t out_capacity_words) {
    if (!ctx || !h_out) return PB_ERR_INVALID_ARG;
    if (!ctx->seg.valid) return PB_ERR_INVALID_ARG;
    const size_t n_queries = ctx->n_queries;
    if (n_queries == 0) return 0;
    const size_t wpq = query_words(ctx->seg.log_n, ctx->seg.width, ctx->seg.perm_width);
    if (out_capacity_words < wpq * n_queries) return PB_ERR_INVALID_ARG;
    int rc;
    std::vector<uint32_t> idx(n_queries);
    Challenger ch = ctx->seg.ch;
    for (size_t q = 0; q < n_queries; q++) idx[q] = h_from_m(ch.sample()) & (uint32_t)(((size_t)1 << ctx->seg.log_m) - 1);
    if ((rc = ctx->ws_qidx.ensure(n_queries))) return rc;
    if ((rc = ctx->ws_qout.ensure(wpq * n_queries))) return rc;
    CK(cudaMemcpyAsync(ctx->ws_qidx.p, idx.data(), 4 * n_queries, cudaMemcpyHostToDevice, ctx->stream));
    fri::QueryDesc d;
    d.lde = ctx->ws_lde.p; d.qlde = ctx->ws_qlde.p; d.tree_t = ctx->ws_layers.p; d.tree_q = ctx->ws_layers_q.p;
    d.plde = ctx->ws_perm_lde.p; d.tree_p = ctx->ws_layers_p.p; d.perm_width = (uint32_t)ctx->seg.perm_width;
    d.fri_words = ctx->ws_fri_words.p; d.fri_trees = ctx->ws_fri_trees.p;
    d.m = (size_t)1 << ctx->seg.log_m; d.width = (uint32_t)ctx->seg.width; d.log_m = (int)ctx->seg.log_m; d.n_layers = (int)ctx->seg.n_layers;
    for (int i = 0; i < 32; i++) { d.word_off[i] = ctx->seg.word_off[i]; d.tree_off[i] = ctx->seg.tree_off[i]; }
    fri::gather_queries_kernel<<<(unsigned)n_queries, 256, 0, ctx->stream>>>(d, ctx->ws_qidx.p, ctx->ws_qout.p, wpq);
    LAUNCHED(ctx);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_out, ctx->ws_qout.p, 4 * wpq * n_queries, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// opened values of the last pb_prove_segment / pb_prove_segment_sharded, canonical, [(width + 2*perm_width + 8)][4]
int pb_last_openings(pb_ctx_t* ctx, uint32_t* h_ys, size_t capacity_words) {
    if (!ctx || !h_ys || (!ctx->seg.valid && !ctx->sh.valid) || capacity_words < ctx->seg.ys.size()) return PB_ERR_INVALID_ARG;
    for (size_t i = 0; i < ctx->seg.ys.size(); i++) h_ys[i] = h_from_m(ctx->seg.ys[i]);
    return 0;
}
