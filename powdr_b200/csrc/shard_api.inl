// One segment across G = 2^g GPUs (north_star "trace-column shards partitioned across the GPUs").
// Included by capi.cu inside its extern "C" block (uses its file-local helpers).
//
// Decomposition.  Input: the trace is COLUMN-sharded (rank r holds columns [r*per, (r+1)*per), per = ceil(W/G)).  Output
// of stage 1: the LDE is ROW-sharded: rank r holds rows [r*Ms, (r+1)*Ms) of the bit-reversed 2N-row LDE, Ms = 2N/G, for
// ALL columns -- which is what Merkle leaf hashing, constraint evaluation, the reduced opening and FRI folding need,
// because each of them is per-row (or per adjacent row pair).  A contiguous block of bit-reversed rows is a sub-coset:
// rows of block (c, h) (c = coset, h = block inside the coset, G1 = G/2 blocks per coset) are the points
//     s' * <w_N^G1>,   s' = shift * w_2N^c * w_N^k0,   k0 = bitrev_{g-1}(h)
// and a polynomial restricted to that sub-coset is the size-N/G1 polynomial  b_r = sum_q lambda^q a_{r + q N/G1},
// lambda = s'^(N/G1).  With the coefficients in bit-reversed slots (what the inverse passes leave), a_{r + q N/G1} are G1
// ADJACENT slots, so the fold is a streaming kernel.  Hence:
//   inverse NTT of my columns  ->  fold for every destination block  ->  ONE all-to-all (the distributed-FFT transpose,
//   2N*W*4 bytes in total)  ->  forward NTT of size N/G1 on my block, all columns.
// After that only digests (32 B per rank and tree), the 8 quotient columns, the opened values and the tail of the FRI
// codeword cross GPUs.  The collectives are caller-supplied callbacks (pb_comm_t): NCCL via torch.distributed in
// bench.py, a barrier + staging copy in the single-GPU thread tests.  The library itself links no communication library.
//
// The transcript, and therefore the proof, is identical to pb_prove_segment's on the same trace (tests/test_gpu_sharded.py).

namespace shard {

struct FoldParams { uint32_t lam[16][8]; };     // lam[s][q] = lambda_s^q (Montgomery), s < n_out <= 16, q < G1 <= 8

// out[s * shard_stride + col * Np + p] = sum_q lam[s][q] * coef[col * N + p * G1 + bitrev_g1(q)]
__global__ void __launch_bounds__(256) coef_fold_kernel(const uint32_t* __restrict__ coef, size_t n_cols, int n, int g1, int n_out, FoldParams P,
                                                        uint32_t* __restrict__ out, size_t shard_stride) {
    const int np = n - g1;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_cols << np)) return;
    const size_t col = i >> np, p = i & (((size_t)1 << np) - 1);
    const uint32_t* a = coef + (col << n) + (p << g1);
    uint32_t v[8];
    const int G1 = 1 << g1;
    for (int q = 0; q < G1; q++) v[q] = __ldg(a + ntt::brev((uint32_t)q, g1));
    for (int s = 0; s < n_out; s++) {
        uint32_t acc = v[0];
        for (int q = 1; q < G1; q++) acc = bb::add(acc, bb::mul(v[q], P.lam[s][q]));
        out[(size_t)s * shard_stride + i] = acc;
    }
}

// column blocks -> row blocks (the transpose that gives a rank whole trace rows for the LogUp phase):
// out[(dest * per + i) * rows + r] = cols[i * N + dest * rows + r]   for i < w_my, zero for the pad columns i >= w_my
__global__ void __launch_bounds__(256) pack_rows_kernel(const uint32_t* __restrict__ cols, size_t w_my, size_t per, size_t N, size_t rows, int G,
                                                        uint32_t* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)G * per * rows) return;
    const size_t r = idx % rows, i = (idx / rows) % per, dest = idx / (rows * per);
    out[idx] = i < w_my ? __ldg(cols + i * N + dest * rows + r) : 0u;
}
// row blocks of my columns -> whole columns:  out[i * N + src * rows + r] = in[(src * per + i) * rows + r]
__global__ void __launch_bounds__(256) unpack_cols_kernel(const uint32_t* __restrict__ in, size_t per, size_t N, size_t rows, int G, uint32_t* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)G * per * rows) return;
    const size_t r = idx % rows, i = (idx / rows) % per, src = idx / (rows * per);
    out[i * N + src * rows + r] = in[idx];
}
// a[l * n + i] += off[l]   (the running sum's offset of this rank's row block)
__global__ void __launch_bounds__(256) add_limb_offsets_kernel(uint32_t* __restrict__ a, size_t n, uint4 off) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = bb::add(a[i], off.x); a[n + i] = bb::add(a[n + i], off.y); a[2 * n + i] = bb::add(a[2 * n + i], off.z); a[3 * n + i] = bb::add(a[3 * n + i], off.w);
}
// logup::finish_kernel for a row block: the next-row values of S and phi live on another rank, so every rank's (S, phi) block is
// gathered first: sp[(blk * 8 + k) * Ms + r], k < 4: S limb k, k >= 4: phi limb k-4.  Writes the quotient VALUES of my rows, [4][Ms].
__global__ void __launch_bounds__(256) finish_block_kernel(const uint32_t* __restrict__ raw, const uint32_t* __restrict__ sp, size_t Ms, size_t row0, int log_n,
                                                           uint32_t shift_m, uint32_t omega_m_m, uint32_t w_n_inv_m, uint32_t sn_m, bb::E4 alpha, bb::E4 alpha2,
                                                           bb::E4 cumsum, uint32_t zinv, uint32_t* __restrict__ out) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Ms) return;
    const int log_m = log_n + 1;
    const size_t m = (size_t)1 << log_m, R = row0 + r, blk = row0 / Ms;
    const uint32_t i_nat = __brev((uint32_t)R) >> (32 - log_m);
    const size_t Rn = __brev((uint32_t)((i_nat + 2) & (uint32_t)(m - 1))) >> (32 - log_m);
    const size_t bn = Rn / Ms, rn = Rn % Ms;
    const uint32_t x = bb::mul(shift_m, bb::pow(omega_m_m, (uint64_t)i_nat));
    const uint32_t zh = bb::sub((i_nat & 1) ? bb::neg(sn_m) : sn_m, bb::R1);
    const uint32_t is_first = bb::mul(zh, bb::inv(bb::sub(x, bb::R1)));
    const uint32_t is_last = bb::mul(zh, bb::inv(bb::sub(x, w_n_inv_m)));
    const uint32_t is_trans = bb::sub(x, w_n_inv_m);
    bb::E4 acc, S, Sn, ph, phn;
#pragma unroll
    for (int l = 0; l < 4; l++) {
        acc.c[l] = raw[(size_t)l * Ms + r];
        S.c[l] = sp[(blk * 8 + l) * Ms + r];
        ph.c[l] = sp[(blk * 8 + 4 + l) * Ms + r];
        Sn.c[l] = sp[(bn * 8 + l) * Ms + rn];
        phn.c[l] = sp[(bn * 8 + 4 + l) * Ms + rn];
    }
    acc = bb::e4_add(acc, bb::e4_mul(alpha2, bb::e4_scale(bb::e4_sub(ph, S), is_first)));
    acc = bb::e4_add(acc, bb::e4_mul(alpha, bb::e4_scale(bb::e4_sub(bb::e4_sub(phn, ph), Sn), is_trans)));
    acc = bb::e4_add(acc, bb::e4_scale(bb::e4_sub(ph, cumsum), is_last));
#pragma unroll
    for (int l = 0; l < 4; l++) out[(size_t)l * Ms + r] = bb::mul(acc.c[l], zinv);
}

__global__ void __launch_bounds__(256) scale_kernel(uint32_t* __restrict__ a, size_t n, uint32_t k) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = bb::mul(a[i], k);
}

}  // namespace shard

namespace {

struct ShardGeom {
    int g, g1, G, G1;
    size_t N, Ms;        // Ms = rows per block = N >> g1 = 2N >> g
    int np;              // log2(Ms)
};

inline bool make_shard_geom(size_t log_n, int world, ShardGeom* s) {
    int g = 0;
    while ((1 << g) < world) g++;
    if ((1 << g) != world || g < 1 || g > 4 || log_n < (size_t)g + 3) return false;
    s->g = g; s->g1 = g - 1; s->G = world; s->G1 = world / 2;
    s->N = (size_t)1 << log_n; s->Ms = s->N >> s->g1; s->np = (int)log_n - s->g1;
    return true;
}

// s' (Montgomery) of block `blk` of the LDE with base shift `shift_m` (log_blowup 1)
inline uint32_t block_shift_m(const ShardGeom& s, size_t log_n, uint32_t shift_m, int blk) {
    const int c = blk >> s.g1, h = blk & (s.G1 - 1);
    uint32_t k0 = 0;
    for (int b = 0; b < s.g1; b++) k0 |= ((h >> b) & 1u) << (s.g1 - 1 - b);
    uint32_t sp = bb::mul(shift_m, bb::pow(h_root_of_unity_m((int)log_n + 1), (uint64_t)c));
    return bb::mul(sp, bb::pow(h_root_of_unity_m((int)log_n), (uint64_t)k0));
}

inline void fold_params_for(const ShardGeom& s, size_t log_n, uint32_t shift_m, int first_blk, int n_blk, shard::FoldParams* P) {
    memset(P, 0, sizeof *P);
    for (int i = 0; i < n_blk; i++) {
        const uint32_t lam = bb::pow(block_shift_m(s, log_n, shift_m, first_blk + i), (uint64_t)s.Ms);
        uint32_t x = bb::R1;
        for (int q = 0; q < s.G1; q++) { P->lam[i][q] = x; x = bb::mul(x, lam); }
    }
}

// inverse NTT of `w` natural-order columns (stride N), then the fold for blocks [first_blk, first_blk + n_blk):
// out[(b - first_blk) * shard_stride + col * Ms + p]
int inverse_and_fold(pb_ctx* ctx, const ShardGeom& s, const uint32_t* d_cols, size_t log_n, size_t w, uint32_t shift, int first_blk, int n_blk,
                     uint32_t* out, size_t shard_stride) {
    if (w == 0) return 0;
    const TwiddleSet* tw;
    int rc = get_twiddles(ctx, (int)log_n, 1, bb::GEN, &tw);       // only the (shift independent) inverse table is used
    if (rc) return rc;
    const NttGeom g = make_geom((int)log_n);
    const size_t batch = lde_column_batch(s.N, w);
    rc = ctx->tmp.ensure(batch * s.N);
    if (rc) return rc;
    shard::FoldParams P;
    fold_params_for(s, log_n, h_to_m(shift), first_blk, n_blk, &P);
    for (size_t c0 = 0; c0 < w; c0 += batch) {
        const unsigned nb = (unsigned)std::min(batch, w - c0);
        ntt_inverse_cols(ctx, g, tw, d_cols + c0 * s.N, nb, ctx->tmp.p);
        const size_t tot = (size_t)nb << s.np;
        shard::coef_fold_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(ctx->tmp.p, nb, (int)log_n, s.g1, n_blk, P,
                                                                                       out + c0 * s.Ms, shard_stride);
        LAUNCHED(ctx);
    }
    CK(cudaGetLastError());
    return 0;
}

// forward NTT (size Ms, one coset with shift s' of block blk) of `w` folded coefficient columns (stride Ms) -> out (stride Ms)
int forward_block(pb_ctx* ctx, const ShardGeom& s, const uint32_t* d_coef, size_t log_n, size_t w, uint32_t shift, int blk, uint32_t* out) {
    if (w == 0) return 0;
    const uint32_t sp = h_from_m(block_shift_m(s, log_n, h_to_m(shift), blk));
    const TwiddleSet* tw;
    int rc = get_twiddles(ctx, s.np, 0, sp, &tw);
    if (rc) return rc;
    const NttGeom g = make_geom(s.np);
    const size_t batch = lde_column_batch(s.Ms, w);
    rc = ctx->tmp2.ensure(batch * s.Ms);
    if (rc) return rc;
    for (size_t c0 = 0; c0 < w; c0 += batch) {
        const unsigned nb = (unsigned)std::min(batch, w - c0);
        ntt_forward_cols(ctx, g, tw, 0, d_coef + c0 * s.Ms, nb, ctx->tmp2.p, out + c0 * s.Ms, s.Ms);
    }
    CK(cudaGetLastError());
    return 0;
}

inline void host_compress(const P2Host& k, const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
    uint32_t st[16];
    memcpy(st, l, 32);
    memcpy(st + 8, r, 32);
    host_permute(st, k);
    memcpy(out, st, 32);
}

// all ranks' subtree roots (device, Montgomery) -> the root of the whole tree (host, Montgomery)
int combine_roots(pb_ctx* ctx, const pb_comm_t* comm, const uint32_t* d_my_root, uint32_t root_m[8], uint32_t (*sub_roots)[8] = nullptr) {
    int rc = ctx->ws_gather.ensure(8 * (size_t)comm->world);
    if (rc) return rc;
    if (!(comm->flags & PB_COMM_STREAM_ORDERED)) CK(cudaStreamSynchronize(ctx->stream));
    rc = comm->all_gather(comm->user, d_my_root, ctx->ws_gather.p, 32);
    if (rc) return PB_ERR_COMM;
    uint32_t nodes[16][8];
    CK(cudaMemcpyAsync(nodes, ctx->ws_gather.p, 32 * (size_t)comm->world, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (sub_roots) memcpy(sub_roots, nodes, 32 * (size_t)comm->world);        // the query phase opens paths through this top tree
    for (int n = comm->world; n > 1; n >>= 1)
        for (int i = 0; i < n / 2; i++) {
            uint32_t t[8];
            host_compress(ctx->p2, nodes[2 * i], nodes[2 * i + 1], t);
            memcpy(nodes[i], t, 32);
        }
    memcpy(root_m, nodes[0], 32);
    return 0;
}

inline const uint32_t* root_ptr(const uint32_t* d_layers, size_t log_h) { return d_layers + 8 * (((size_t)2 << log_h) - 2); }

}  // namespace

int pb_shard_columns(size_t width, int world, int rank, size_t* first, size_t* count) {
    if (world < 1 || rank < 0 || rank >= world || !first || !count) return PB_ERR_INVALID_ARG;
    const size_t per = (width + (size_t)world - 1) / (size_t)world;
    *first = std::min(width, (size_t)rank * per);
    *count = std::min(per, width - *first);
    return 0;
}

// rows [blk * Ms, (blk + 1) * Ms) of pb_lde_batch(..., log_blowup 1, shift)'s output, every column, computed without the rest:
// d_out column-major [width][Ms].  (All columns are inverse-transformed here; the multi-GPU prover shares that work instead.)
int pb_lde_shard(pb_ctx_t* ctx, const uint32_t* d_trace, size_t log_n, size_t width, uint32_t shift, int world, int blk, uint32_t* d_out) {
    if (!ctx || !d_trace || !d_out) return PB_ERR_INVALID_ARG;
    ShardGeom s;
    if (log_n > 24 || shift == 0 || shift >= bb::P || !make_shard_geom(log_n, world, &s) || blk < 0 || blk >= world) return PB_ERR_UNSUPPORTED;
    if (width == 0) return 0;
    int rc = ctx->ws_shard_coef.ensure(width * s.Ms);
    if (rc) return rc;
    rc = inverse_and_fold(ctx, s, d_trace, log_n, width, shift, blk, 1, ctx->ws_shard_coef.p, 0);
    if (rc) return rc;
    return forward_block(ctx, s, ctx->ws_shard_coef.p, log_n, width, shift, blk, d_out);
}

int pb_prove_segment_sharded(pb_ctx_t* ctx, const pb_air_t* a, const uint32_t* trace_cols, size_t log_n, size_t width, uint32_t flags,
                             const pb_comm_t* comm, pb_segment_proof_t* proof) {
    if (!ctx || !a || !proof || !comm || !comm->all_gather || !comm->all_to_all) return PB_ERR_INVALID_ARG;
    if (log_n < 1 || log_n > 24 || width == 0 || width != a->width) return PB_ERR_INVALID_ARG;
    if (a->has_lu && ((size_t)1 << log_n) % (size_t)comm->world != 0) return PB_ERR_UNSUPPORTED;
    ShardGeom s;
    if (!make_shard_geom(log_n, comm->world, &s) || comm->rank < 0 || comm->rank >= comm->world) return PB_ERR_UNSUPPORTED;
    const int G = s.G, rho = comm->rank;
    const size_t N = s.N, Ms = s.Ms, log_m = log_n + 1, log_ms = (size_t)s.np;
    const size_t per = (width + (size_t)G - 1) / (size_t)G;
    const size_t n_chunks = a->has_lu ? a->lu.n_chunks() : 0, wp = a->has_lu ? a->lu.perm_width() : 0;
    const size_t perp = (wp + (size_t)G - 1) / (size_t)G, rows_t = N / (size_t)G;      // perm columns per rank; trace rows per rank
    const size_t pmax = std::max(per, perp);
    size_t c_first, w_my;
    pb_shard_columns(width, G, rho, &c_first, &w_my);
    if (w_my && !trace_cols) return PB_ERR_INVALID_ARG;
    int rc;
    memset(proof, 0, sizeof *proof);
    proof->pow_bits = ctx->pow_bits;
    proof->n_queries = ctx->n_queries;
    proof->perm_width = (uint32_t)(a->has_lu ? a->lu.perm_width() : 0);
    cudaStream_t st = ctx->stream;
    ctx->seg.valid = false;
    ctx->sh.valid = false;
    ctx->mc.valid = false;
#define RC(x) do { rc = (x); if (rc) return rc; } while (0)
#define COMM(fn, send, recv, bytes) do { if (!(comm->flags & PB_COMM_STREAM_ORDERED)) CK(cudaStreamSynchronize(st)); if (comm->fn(comm->user, (send), (recv), (bytes))) return PB_ERR_COMM; } while (0)
    CK(cudaEventRecord(ctx->ev[0], st));
    RC(ctx->ws_shard_send.ensure((size_t)G * pmax * Ms));
    RC(ctx->ws_shard_recv.ensure((size_t)G * pmax * Ms));
    if (wp) {
        RC(ctx->ws_perm.ensure((size_t)G * perp * rows_t + 4 * rows_t));      // my trace rows x all perm columns (padded to G*perp), + row sums
        RC(ctx->ws_perm_lde.ensure(wp * Ms));
        RC(ctx->ws_layers_p.ensure(8 * (2 * Ms)));
        RC(ctx->ws_rowsum.ensure(perp * N));                                   // my perm columns, whole (column-sharded copy)
        RC(ctx->ws_lu_raw.ensure(4 * Ms));
        RC(ctx->ws_lu_s.ensure(8 * Ms + (size_t)G * 8 * Ms));
    }
    RC(ctx->ws_lde.ensure(width * Ms));
    RC(ctx->ws_layers.ensure(8 * (2 * Ms)));
    RC(ctx->ws_layers_q.ensure(8 * (2 * Ms)));
    RC(ctx->ws_q.ensure(8 * N));
    RC(ctx->ws_qnat.ensure(8 * N));
    RC(ctx->ws_qlde.ensure(8 * Ms));
    RC(ctx->ws_shard_coef.ensure(std::max((size_t)G * 4 * Ms, 4 * 2 * N)));
    RC(ctx->ws_f0.ensure(4 * 2 * N));
    RC(ctx->ws_f1.ensure(4 * 2 * N));
    Challenger ch;
    ch.k = &ctx->p2;
    uint32_t root_m[8];

    // ---- stage 0/1: (copy,) inverse NTT of my columns, fold for every block, all-to-all, forward NTT of my block ----
    const uint32_t* d_my = trace_cols;
    if (!(flags & PB_TRACE_ON_DEVICE) && w_my) {
        RC(ctx->ws_trace.ensure(w_my * N));
        CK(cudaMemcpyAsync(ctx->ws_trace.p, trace_cols, w_my * N * 4, cudaMemcpyHostToDevice, st));
        d_my = ctx->ws_trace.p;
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    if (w_my < per)      // pad columns of the last ranks: defined contents, never read back
        for (int b = 0; b < G; b++) CK(cudaMemsetAsync(ctx->ws_shard_send.p + ((size_t)b * per + w_my) * Ms, 0, (per - w_my) * Ms * 4, st));
    RC(inverse_and_fold(ctx, s, d_my, log_n, w_my, bb::GEN, 0, G, ctx->ws_shard_send.p, per * Ms));
    COMM(all_to_all, ctx->ws_shard_send.p, ctx->ws_shard_recv.p, per * Ms * 4);
    // recv = [source rank][per][Ms]: global column j = source * per + i, so the first `width` columns are the real ones
    RC(forward_block(ctx, s, ctx->ws_shard_recv.p, log_n, width, bb::GEN, rho, ctx->ws_lde.p));
    CK(cudaEventRecord(ctx->ev[2], st));

    // ---- stage 3a: Merkle subtree over my rows; the G subtree roots are the nodes of level log_ms ----
    {
        const uint32_t* mats1[1] = {ctx->ws_lde.p};
        RC(pb_merkle_commit(ctx, mats1, &width, 1, log_ms, ctx->ws_layers.p, nullptr));
    }
    CK(cudaEventRecord(ctx->ev[3], st));
    RC(combine_roots(ctx, comm, root_ptr(ctx->ws_layers.p, log_ms), root_m, ctx->sh.roots_main));
    for (int i = 0; i < 8; i++) proof->trace_root[i] = h_from_m(root_m[i]);
    ch.observe(root_m, 8);

    // ---- LogUp phase.  The permutation trace needs whole trace ROWS: transpose my column block into row blocks (all-to-all), generate
    //      the permutation columns of my rows, scan the running sum (local scan + offsets from the gathered block totals), transpose the
    //      permutation trace back into column blocks (all-to-all) and push it through the same LDE / Merkle path as the main trace ----
    bb::E4 cumsum = {{0u, 0u, 0u, 0u}};
    size_t wp_first = 0, wp_my = 0;
    if (wp) {
        pb_shard_columns(wp, G, rho, &wp_first, &wp_my);
        const bb::E4 al = ch.sample_ext(), be = ch.sample_ext();
        for (int i = 0; i < 4; i++) { proof->logup_alpha[i] = h_from_m(al.c[i]); proof->logup_beta[i] = h_from_m(be.c[i]); }
        RC(upload_logup_consts(ctx, a, al, be));
        const size_t tot_t = (size_t)G * per * rows_t;
        shard::pack_rows_kernel<<<(unsigned)((tot_t + 255) / 256), 256, 0, st>>>(d_my, w_my, per, N, rows_t, G, ctx->ws_shard_send.p);
        LAUNCHED(ctx);
        COMM(all_to_all, ctx->ws_shard_send.p, ctx->ws_shard_recv.p, per * rows_t * 4);
        // recv = [source][per][rows_t] = column-major, column j = source * per + i: my rows of every trace column
        uint32_t* perm_rows = ctx->ws_perm.p;                           // [G * perp][rows_t], the first wp columns are real
        uint32_t* rowsum = ctx->ws_perm.p + (size_t)G * perp * rows_t;  // [4][rows_t]
        if ((size_t)G * perp > wp) CK(cudaMemsetAsync(perm_rows + wp * rows_t, 0, ((size_t)G * perp - wp) * rows_t * 4, st));
        RC(logup::launch_perm(a->lujit, st, ctx->ws_shard_recv.p, rows_t, a->d_kc, a->d_bt, perm_rows, rowsum));
        LAUNCHED(ctx);
        {
            const unsigned nb = (unsigned)((rows_t + logup::SCAN_THREADS * logup::SCAN_ITEMS - 1) / (logup::SCAN_THREADS * logup::SCAN_ITEMS));
            RC(ctx->ws_scan_tot.ensure(4 * (size_t)nb + 8 + 8 * (size_t)G));
            uint32_t* phi = perm_rows + 4 * n_chunks * rows_t;
            logup::scan_local_kernel<<<dim3(nb, 4), logup::SCAN_THREADS, 0, st>>>(rowsum, phi, rows_t, ctx->ws_scan_tot.p);
            logup::scan_totals_kernel<<<1, 32, 0, st>>>(ctx->ws_scan_tot.p, nb);
            logup::scan_add_kernel<<<dim3((unsigned)((rows_t + 255) / 256), 4), 256, 0, st>>>(phi, rows_t, ctx->ws_scan_tot.p, nb);
            ctx->launches += 3;
            // block totals (last row of the local scan) of every rank -> my offset and the cumulative sum
            uint32_t* tot_my = ctx->ws_scan_tot.p + 4 * (size_t)nb;       // 8 words (4 used; 32-byte exchange unit)
            uint32_t* tot_all = tot_my + 8;
            CK(cudaMemsetAsync(tot_my, 0, 32, st));
            for (int l = 0; l < 4; l++) CK(cudaMemcpyAsync(tot_my + l, phi + (size_t)l * rows_t + (rows_t - 1), 4, cudaMemcpyDeviceToDevice, st));
            COMM(all_gather, tot_my, tot_all, 32);
            uint32_t th[16 * 8];
            CK(cudaMemcpyAsync(th, tot_all, 32 * (size_t)G, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            bb::E4 off = {{0u, 0u, 0u, 0u}};
            for (int b = 0; b < G; b++)
                for (int l = 0; l < 4; l++) {
                    if (b < rho) off.c[l] = bb::add(off.c[l], th[8 * b + l]);
                    cumsum.c[l] = bb::add(cumsum.c[l], th[8 * b + l]);
                }
            shard::add_limb_offsets_kernel<<<(unsigned)((rows_t + 255) / 256), 256, 0, st>>>(phi, rows_t, make_uint4(off.c[0], off.c[1], off.c[2], off.c[3]));
            LAUNCHED(ctx);
            for (int l = 0; l < 4; l++) proof->cumulative_sum[l] = h_from_m(cumsum.c[l]);
        }
        CK(cudaEventRecord(ctx->ev[12], st));
        // row blocks -> column blocks: destination d gets columns [d * perp, (d + 1) * perp) of my rows = a contiguous slice of perm_rows
        COMM(all_to_all, perm_rows, ctx->ws_shard_recv.p, perp * rows_t * 4);
        uint32_t* perm_cols = ctx->ws_rowsum.p;                         // [perp][N]: my permutation columns, whole
        shard::unpack_cols_kernel<<<(unsigned)(((size_t)G * perp * rows_t + 255) / 256), 256, 0, st>>>(ctx->ws_shard_recv.p, perp, N, rows_t, G, perm_cols);
        LAUNCHED(ctx);
        // LDE of the permutation trace, row-sharded like the main one
        if (wp_my < perp)
            for (int b = 0; b < G; b++) CK(cudaMemsetAsync(ctx->ws_shard_send.p + ((size_t)b * perp + wp_my) * Ms, 0, (perp - wp_my) * Ms * 4, st));
        RC(inverse_and_fold(ctx, s, perm_cols, log_n, wp_my, bb::GEN, 0, G, ctx->ws_shard_send.p, perp * Ms));
        COMM(all_to_all, ctx->ws_shard_send.p, ctx->ws_shard_recv.p, perp * Ms * 4);
        RC(forward_block(ctx, s, ctx->ws_shard_recv.p, log_n, wp, bb::GEN, rho, ctx->ws_perm_lde.p));
        {
            const uint32_t* matsp[1] = {ctx->ws_perm_lde.p};
            RC(pb_merkle_commit(ctx, matsp, &wp, 1, log_ms, ctx->ws_layers_p.p, nullptr));
        }
        RC(combine_roots(ctx, comm, root_ptr(ctx->ws_layers_p.p, log_ms), root_m, ctx->sh.roots_perm));
        for (int i = 0; i < 8; i++) proof->perm_root[i] = h_from_m(root_m[i]);
        ch.observe(root_m, 8);
        ch.observe(cumsum.c, 4);
    } else {
        CK(cudaEventRecord(ctx->ev[12], st));
    }
    CK(cudaEventRecord(ctx->ev[13], st));
    const bb::E4 alpha = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->alpha[i] = h_from_m(alpha.c[i]);

    // ---- stage 2: quotient values on my rows (all in chunk c = rho >> g1), gathered so that every rank holds both chunks ----
    {
        uint32_t* q_my = ctx->ws_shard_coef.p;                 // [4][Ms]
        uint32_t* q_all = ctx->ws_f0.p;                        // [G][4][Ms]  (ws_f0 is free until FRI)
        const uint32_t sn = bb::pow(h_to_m(bb::GEN), (uint64_t)1 << log_n);
        const uint32_t zinv = (rho >> s.g1) ? bb::inv(bb::sub(bb::neg(sn), bb::R1)) : bb::inv(bb::sub(sn, bb::R1));
        if (!wp) {
            RC(pb_constraint_fold(ctx, a, ctx->ws_lde.p, Ms, proof->alpha, q_my));
            shard::scale_kernel<<<(unsigned)((4 * Ms + 255) / 256), 256, 0, st>>>(q_my, 4 * Ms, zinv);
            LAUNCHED(ctx);
        } else {
            RC(constraint_fold_m(ctx, a, ctx->ws_lde.p, Ms, alpha, n_chunks + 3, ctx->ws_lu_raw.p));
            {
                std::vector<bb::E4> apl(std::max<size_t>(1, n_chunks));
                bb::E4 cur = bb::e4_mul(alpha, alpha);
                for (size_t c = n_chunks; c-- > 0;) { cur = bb::e4_mul(cur, alpha); apl[c] = cur; }
                CK(cudaMemcpyAsync(a->d_apl, apl.data(), n_chunks * 16, cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));
            }
            uint32_t* sp_my = ctx->ws_lu_s.p;                 // [8][Ms]: S limbs then phi limbs of my rows
            uint32_t* sp_all = ctx->ws_lu_s.p + 8 * Ms;       // [G][8][Ms]
            RC(logup::launch_fold(a->lujit, st, ctx->ws_lde.p, ctx->ws_perm_lde.p, Ms, a->d_kc, a->d_bt, a->d_apl, ctx->ws_lu_raw.p, sp_my));
            LAUNCHED(ctx);
            CK(cudaMemcpyAsync(sp_my + 4 * Ms, ctx->ws_perm_lde.p + 4 * n_chunks * Ms, 4 * Ms * 4, cudaMemcpyDeviceToDevice, st));
            COMM(all_gather, sp_my, sp_all, 8 * Ms * 4);
            shard::finish_block_kernel<<<(unsigned)((Ms + 255) / 256), 256, 0, st>>>(ctx->ws_lu_raw.p, sp_all, Ms, (size_t)rho * Ms, (int)log_n, h_to_m(bb::GEN),
                                                                                    h_root_of_unity_m((int)log_m), bb::inv(h_root_of_unity_m((int)log_n)), sn, alpha,
                                                                                    bb::e4_mul(alpha, alpha), cumsum, zinv, q_my);
            LAUNCHED(ctx);
        }
        CK(cudaEventRecord(ctx->ev[4], st));
        COMM(all_gather, q_my, q_all, 4 * Ms * 4);
        for (int b = 0; b < G; b++) {
            const size_t c = (size_t)(b >> s.g1), h = (size_t)(b & (s.G1 - 1));
            for (size_t l = 0; l < 4; l++)
                CK(cudaMemcpyAsync(ctx->ws_q.p + (c * 4 + l) * N + h * Ms, q_all + ((size_t)b * 4 + l) * Ms, Ms * 4, cudaMemcpyDeviceToDevice, st));
        }
        const size_t tot = 8 * N;
        ntt::bitrev_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(ctx->ws_q.p, ctx->ws_qnat.p, (int)log_n, 8);
        LAUNCHED(ctx);
        // quotient chunk LDEs: every rank has the 8 columns, so each computes only its own block of rows
        const uint32_t w2n_inv = h_from_m(bb::inv(h_root_of_unity_m((int)log_n + 1)));
        RC(inverse_and_fold(ctx, s, ctx->ws_qnat.p, log_n, 4, 1u, rho, 1, ctx->ws_shard_coef.p, 0));
        RC(forward_block(ctx, s, ctx->ws_shard_coef.p, log_n, 4, 1u, rho, ctx->ws_qlde.p));
        RC(inverse_and_fold(ctx, s, ctx->ws_qnat.p + 4 * N, log_n, 4, w2n_inv, rho, 1, ctx->ws_shard_coef.p, 0));
        RC(forward_block(ctx, s, ctx->ws_shard_coef.p, log_n, 4, w2n_inv, rho, ctx->ws_qlde.p + 4 * Ms));
    }
    CK(cudaEventRecord(ctx->ev[5], st));
    {
        const uint32_t* mats2[2] = {ctx->ws_qlde.p, ctx->ws_qlde.p + 4 * Ms};
        const size_t w2[2] = {4, 4};
        RC(pb_merkle_commit(ctx, mats2, w2, 2, log_ms, ctx->ws_layers_q.p, nullptr));
    }
    CK(cudaEventRecord(ctx->ev[6], st));
    RC(combine_roots(ctx, comm, root_ptr(ctx->ws_layers_q.p, log_ms), root_m, ctx->sh.roots_q));
    for (int i = 0; i < 8; i++) proof->quotient_root[i] = h_from_m(root_m[i]);
    ch.observe(root_m, 8);
    const bb::E4 zeta = ch.sample_ext();
    for (int i = 0; i < 4; i++) proof->zeta[i] = h_from_m(zeta.c[i]);

    // ---- openings at zeta: my trace columns, gathered; the 8 quotient columns on every rank; every value observed ----
    const size_t n_open = width + 2 * wp + 8;
    RC(ctx->ws_ys.ensure(4 * (n_open + (size_t)G * pmax)));
    RC(ctx->ws_gather2.ensure(4 * pmax + 4 * (size_t)G * pmax));
    {
        uint32_t* ys_my = ctx->ws_gather2.p;                    // [pmax][4]
        uint32_t* ys_all = ctx->ws_gather2.p + 4 * pmax;       // [G*pmax][4]
        CK(cudaMemsetAsync(ys_my, 0, 16 * per, st));
        if (w_my) RC(eval_at_point_m(ctx, d_my, log_n, w_my, h_to_m(1u), zeta, ys_my));
        COMM(all_gather, ys_my, ys_all, 16 * per);
        CK(cudaMemcpyAsync(ctx->ws_ys.p, ys_all, 16 * width, cudaMemcpyDeviceToDevice, st));
        if (wp) {       // my permutation columns at zeta and at zeta*w
            const bb::E4 zeta_next = bb::e4_scale(zeta, h_root_of_unity_m((int)log_n));
            for (int pt = 0; pt < 2; pt++) {
                CK(cudaMemsetAsync(ys_my, 0, 16 * perp, st));
                if (wp_my) RC(eval_at_point_m(ctx, ctx->ws_rowsum.p, log_n, wp_my, h_to_m(1u), pt ? zeta_next : zeta, ys_my));
                COMM(all_gather, ys_my, ys_all, 16 * perp);
                CK(cudaMemcpyAsync(ctx->ws_ys.p + 4 * (width + (size_t)pt * wp), ys_all, 16 * wp, cudaMemcpyDeviceToDevice, st));
            }
        }
        const uint32_t g_c = bb::GEN, gw_c = h_from_m(bb::mul(h_to_m(bb::GEN), h_root_of_unity_m((int)log_n + 1)));
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

    // ---- reduced opening on my rows: layer 0 of the FRI codewords, which all stay resident (my part of every sharded layer, the whole
    //      codeword of the replicated ones) together with the layer trees -- the query phase opens them ----
    size_t SMALL = 14;
    if (const char* e = getenv("PB_SHARD_FRI_SMALL")) SMALL = std::min<size_t>(24, std::max<size_t>(2, (size_t)atol(e)));
    {
        // offsets of every layer's codeword and tree in ws_fri_words / ws_fri_trees (a layer is sharded while log_len - g >= SMALL)
        size_t woff = 0, toff = 0, ll = log_m;
        uint32_t li = 0;
        bool shd = true;
        while (ll > 1) {
            if (shd && ll - (size_t)s.g < SMALL) shd = false;
            ctx->sh.layer_sharded[li] = shd;
            ctx->sh.word_off[li] = woff;
            ctx->sh.tree_off[li] = toff;
            const size_t log_rows = shd ? ll - 1 - (size_t)s.g : ll - 1;
            woff += (size_t)8 << log_rows;
            toff += 8 * (((size_t)2 << log_rows) - 1);
            ll--;
            li++;
        }
        ctx->sh.word_off[li] = woff;                     // the final values (2 Ext4) follow the last layer
        RC(ctx->ws_fri_words.ensure(woff + 64));
        RC(ctx->ws_fri_trees.ensure(toff + 64));
    }
    uint32_t* f = ctx->ws_fri_words.p;
    {
        std::vector<const uint32_t*> cols(n_open);
        std::vector<uint32_t> grp(n_open, 0u);
        size_t k = 0;
        for (size_t c = 0; c < width; c++) cols[k++] = ctx->ws_lde.p + c * Ms;
        for (size_t c = 0; c < wp; c++) cols[k++] = ctx->ws_perm_lde.p + c * Ms;
        for (size_t c = 0; c < wp; c++) { grp[k] = 1; cols[k++] = ctx->ws_perm_lde.p + c * Ms; }
        for (size_t c = 0; c < 8; c++) cols[k++] = ctx->ws_qlde.p + c * Ms;
        std::vector<bb::E4> zs{zeta};
        if (wp) zs.push_back(bb::e4_scale(zeta, h_root_of_unity_m((int)log_n)));
        RC(deep_quotient_groups_m(ctx, cols, grp, zs, log_m, h_to_m(bb::GEN), gamma, ys_h.data(), f, (size_t)rho * Ms, Ms));
    }
    CK(cudaEventRecord(ctx->ev[8], st));

    // ---- FRI commit phase: fold partners are adjacent, so a row block folds locally; per layer only the subtree root travels.
    //      Below 2^SMALL entries per rank the codeword is gathered once and the remaining layers run replicated: every sharded layer
    //      costs one (latency-bound) root exchange, a replicated layer of 2^k entries costs microseconds. ----
    uint32_t layer = 0;
    bool sharded = true;
    size_t log_len = log_m;
    uint32_t shift_m = h_to_m(bb::GEN);
    while (log_len > 1) {
        if (sharded && !ctx->sh.layer_sharded[layer]) {
            // my part of this layer's codeword sits where the last fold wrote it; the gathered whole codeword takes this layer's slot
            const size_t loc = (size_t)1 << (log_len - s.g);
            RC(ctx->ws_f0.ensure(4 * loc));
            CK(cudaMemcpyAsync(ctx->ws_f0.p, f, 16 * loc, cudaMemcpyDeviceToDevice, st));
            COMM(all_gather, ctx->ws_f0.p, f, 16 * loc);
            sharded = false;
        }
        uint32_t* tree = ctx->ws_fri_trees.p + ctx->sh.tree_off[layer];
        uint32_t* f_next = ctx->ws_fri_words.p + ctx->sh.word_off[layer + 1];      // (the slot after the last layer holds the final values)
        if (sharded) {
            const size_t log_rows = log_len - 1 - (size_t)s.g, loc_out = (size_t)1 << log_rows;
            RC(pb_merkle_commit_rows8(ctx, f, log_rows, tree, nullptr));
            RC(combine_roots(ctx, comm, root_ptr(tree, log_rows), root_m, ctx->sh.roots_fri[layer]));
            for (int i = 0; i < 8; i++) proof->fri_roots[layer][i] = h_from_m(root_m[i]);
            ch.observe(root_m, 8);
            const bb::E4 beta = ch.sample_ext();
            for (int i = 0; i < 4; i++) proof->fri_betas[layer][i] = h_from_m(beta.c[i]);
            RC(fri_fold_m(ctx, f, log_len, shift_m, beta, f_next, (size_t)rho * loc_out, loc_out));
        } else {
            RC(pb_merkle_commit_rows8(ctx, f, log_len - 1, tree, nullptr));
            RC(read_root(ctx, tree, log_len - 1, root_m));
            for (int i = 0; i < 8; i++) proof->fri_roots[layer][i] = h_from_m(root_m[i]);
            ch.observe(root_m, 8);
            const bb::E4 beta = ch.sample_ext();
            for (int i = 0; i < 4; i++) proof->fri_betas[layer][i] = h_from_m(beta.c[i]);
            RC(fri_fold_m(ctx, f, log_len, shift_m, beta, f_next));
        }
        f = f_next;
        shift_m = bb::mul(shift_m, shift_m);
        log_len--;
        layer++;
    }
    if (sharded) return PB_ERR_UNSUPPORTED;      // unreachable: make_shard_geom bounds log_n from below
    proof->n_fri_layers = layer;
    proof->final_len = 1u << log_len;
    uint32_t fin[8 * 4];
    CK(cudaMemcpyAsync(fin, f, 16 * proof->final_len, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[7], st));
    CK(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < proof->final_len; i++)
        for (int l = 0; l < 4; l++) proof->final_poly[i][l] = h_from_m(fin[4 * i + l]);
    ch.observe(fin, 4);
    {   // proof of work: every rank grinds the same transcript state (2^pow_bits permutations: microseconds)
        uint32_t w = 0;
        RC(grind(ctx, ch, ctx->pow_bits, &w));
        proof->pow_witness = w;
        const uint32_t w_m = h_to_m(w);
        ch.observe(&w_m, 1);
        (void)ch.sample();
    }
    CK(cudaEventRecord(ctx->ev[9], st));
    CK(cudaStreamSynchronize(st));
    ctx->sh.valid = true;
    ctx->sh.log_n = log_n; ctx->sh.width = width; ctx->sh.perm_width = wp; ctx->sh.G = G; ctx->sh.g = s.g; ctx->sh.rank = rho; ctx->sh.n_layers = layer;
    ctx->sh.ch = ch;
    // stage clocks in the layout of pb_last_stage_ms (no LogUp phase here)
    float t[8];
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&t[i], ctx->ev[i], ctx->ev[i + 1]);
    cudaEventElapsedTime(&t[6], ctx->ev[6], ctx->ev[8]);
    cudaEventElapsedTime(&t[7], ctx->ev[8], ctx->ev[7]);
    memset(ctx->stage_ms, 0, sizeof ctx->stage_ms);
    ctx->stage_ms[0] = t[0]; ctx->stage_ms[1] = t[1]; ctx->stage_ms[2] = t[2]; ctx->stage_ms[6] = t[4];
    cudaEventElapsedTime(&ctx->stage_ms[3], ctx->ev[3], ctx->ev[12]);      // LogUp: trace transpose + permutation rows + running sum
    cudaEventElapsedTime(&ctx->stage_ms[4], ctx->ev[12], ctx->ev[13]);     // transpose back + LDE + Merkle of the permutation trace
    cudaEventElapsedTime(&ctx->stage_ms[5], ctx->ev[13], ctx->ev[4]);      // quotient
    ctx->stage_ms[7] = t[5]; ctx->stage_ms[8] = t[6]; ctx->stage_ms[9] = t[7];
    cudaEventElapsedTime(&ctx->stage_ms[10], ctx->ev[7], ctx->ev[9]);
    cudaEventElapsedTime(&ctx->stage_ms[11], ctx->ev[0], ctx->ev[9]);
#undef COMM
#undef RC
    return 0;
}

// ---- query phase of the sharded prover ---------------------------------------------------------------------------------
// Same indices and the same output as pb_query_segment after pb_prove_segment on the gathered trace.  Row r of a committed matrix
// lives on rank r >> log2(Ms) together with the bottom log2(Ms) levels of its Merkle path (that rank's subtree); the top g levels run
// through the G subtree roots every rank kept from the commit.  Each rank gathers what it owns into a zeroed [n_queries][words]
// buffer, ONE all-gather brings the shares together (every word has exactly one owner, so the shares add up), the host fills in the
// top-of-tree siblings.
namespace shard {

struct QueryDesc {
    const uint32_t *lde, *plde, *qlde, *tree_t, *tree_p, *tree_q, *fri_words, *fri_trees;
    size_t ms;                       // rows per rank of the LDE matrices
    uint32_t width, perm_width;
    int log_m, log_ms, g, rank, n_layers;
    size_t word_off[32], tree_off[32];
    uint32_t sharded_mask;           // bit i: FRI layer i is sharded
};

// levels [0, n_local) of the path of leaf `idx_local` in a local subtree (layers node-major); canonical words
__device__ __forceinline__ void copy_local_path(const uint32_t* tree, int log_h, size_t idx_local, uint32_t* out) {
    for (int k = 0; k < log_h; k++) {
        const size_t level_off = ((size_t)2 << log_h) - ((size_t)2 << (log_h - k));
        const uint32_t* node = tree + 8 * (level_off + ((idx_local >> k) ^ 1));
        for (int e = threadIdx.x; e < 8; e += blockDim.x) out[8 * k + e] = bb::from_monty(node[e]);
    }
}

__global__ void __launch_bounds__(256) gather_share_kernel(QueryDesc d, const uint32_t* __restrict__ indices, uint32_t* __restrict__ out, size_t wpq) {
    const size_t r = indices[blockIdx.x];
    uint32_t* o = out + (size_t)blockIdx.x * wpq;
    const bool own = (int)(r >> d.log_ms) == d.rank;
    const size_t rl = r & (d.ms - 1);
    if (d.rank == 0 && threadIdx.x == 0) o[0] = (uint32_t)r;
    o += 1;
    if (own) {
        for (uint32_t c = threadIdx.x; c < d.width; c += blockDim.x) o[c] = bb::from_monty(d.lde[(size_t)c * d.ms + rl]);
        copy_local_path(d.tree_t, d.log_ms, rl, o + d.width);
    }
    o += d.width + 8 * d.log_m;
    if (d.perm_width) {
        if (own) {
            for (uint32_t c = threadIdx.x; c < d.perm_width; c += blockDim.x) o[c] = bb::from_monty(d.plde[(size_t)c * d.ms + rl]);
            copy_local_path(d.tree_p, d.log_ms, rl, o + d.perm_width);
        }
        o += d.perm_width + 8 * d.log_m;
    }
    if (own) {
        for (uint32_t c = threadIdx.x; c < 8; c += blockDim.x) o[c] = bb::from_monty(d.qlde[(size_t)c * d.ms + rl]);
        copy_local_path(d.tree_q, d.log_ms, rl, o + 8);
    }
    o += 8 + 8 * d.log_m;
    for (int i = 0; i < d.n_layers; i++) {
        const int log_h = d.log_m - 1 - i;                 // rows of the layer's tree (pairs)
        const size_t j = r >> (i + 1);
        if ((d.sharded_mask >> i) & 1u) {
            const int log_loc = log_h - d.g;
            if ((int)(j >> log_loc) == d.rank) {
                const size_t jl = j & (((size_t)1 << log_loc) - 1);
                const uint32_t* row = d.fri_words + d.word_off[i] + 8 * jl;
                for (uint32_t e = threadIdx.x; e < 8; e += blockDim.x) o[e] = bb::from_monty(row[e]);
                copy_local_path(d.fri_trees + d.tree_off[i], log_loc, jl, o + 8);
            }
        } else if (d.rank == 0) {
            const uint32_t* row = d.fri_words + d.word_off[i] + 8 * j;
            for (uint32_t e = threadIdx.x; e < 8; e += blockDim.x) o[e] = bb::from_monty(row[e]);
            copy_local_path(d.fri_trees + d.tree_off[i], log_h, j, o + 8);
        }
        o += 8 + 8 * log_h;
    }
}

// out[i] = sum over ranks of all[rank * n + i]  (one non-zero term per word)
__global__ void __launch_bounds__(256) sum_shares_kernel(const uint32_t* __restrict__ all, int G, size_t n, uint32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v = 0;
    for (int b = 0; b < G; b++) v += all[(size_t)b * n + i];
    out[i] = v;
}

}  // namespace shard

namespace {
// siblings along the path of leaf `blk` in the tree over the G subtree roots, bottom-up, canonical words: out[g][8]
void top_path(const P2Host& k, const uint32_t (*roots)[8], int G, int blk, uint32_t* out) {
    uint32_t nodes[16][8];
    memcpy(nodes, roots, 32 * (size_t)G);
    int lvl = 0;
    for (int n = G; n > 1; n >>= 1, lvl++) {
        const int sib = (blk >> lvl) ^ 1;
        for (int e = 0; e < 8; e++) out[8 * lvl + e] = h_from_m(nodes[sib][e]);
        for (int i = 0; i < n / 2; i++) {
            uint32_t t[8];
            host_compress(k, nodes[2 * i], nodes[2 * i + 1], t);
            memcpy(nodes[i], t, 32);
        }
    }
}
}  // namespace

int pb_query_segment_sharded(pb_ctx_t* ctx, const pb_comm_t* comm, uint32_t* h_out, size_t out_capacity_words) {
    if (!ctx || !comm || !comm->all_gather || !h_out) return PB_ERR_INVALID_ARG;
    if (!ctx->sh.valid || comm->world != ctx->sh.G || comm->rank != ctx->sh.rank) return PB_ERR_INVALID_ARG;
    const size_t nq = ctx->n_queries;
    if (nq == 0) return 0;
    const size_t log_n = ctx->sh.log_n, log_m = log_n + 1, width = ctx->sh.width, wp = ctx->sh.perm_width;
    const int G = ctx->sh.G, g = ctx->sh.g, rho = ctx->sh.rank;
    const size_t log_ms = log_m - (size_t)g, Ms = (size_t)1 << log_ms;
    const size_t wpq = query_words(log_n, width, wp);
    if (out_capacity_words < wpq * nq) return PB_ERR_INVALID_ARG;
    int rc;
    cudaStream_t st = ctx->stream;
    std::vector<uint32_t> idx(nq);
    Challenger ch = ctx->sh.ch;
    for (size_t q = 0; q < nq; q++) idx[q] = h_from_m(ch.sample()) & (uint32_t)(((size_t)1 << log_m) - 1);
    if ((rc = ctx->ws_qidx.ensure(nq))) return rc;
    if ((rc = ctx->ws_sh_q.ensure(wpq * nq))) return rc;
    if ((rc = ctx->ws_sh_qall.ensure((size_t)G * wpq * nq))) return rc;
    if ((rc = ctx->ws_qout.ensure(wpq * nq))) return rc;
    CK(cudaMemcpyAsync(ctx->ws_qidx.p, idx.data(), 4 * nq, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->ws_sh_q.p, 0, 4 * wpq * nq, st));
    shard::QueryDesc d;
    d.lde = ctx->ws_lde.p; d.plde = ctx->ws_perm_lde.p; d.qlde = ctx->ws_qlde.p;
    d.tree_t = ctx->ws_layers.p; d.tree_p = ctx->ws_layers_p.p; d.tree_q = ctx->ws_layers_q.p;
    d.fri_words = ctx->ws_fri_words.p; d.fri_trees = ctx->ws_fri_trees.p;
    d.ms = Ms; d.width = (uint32_t)width; d.perm_width = (uint32_t)wp;
    d.log_m = (int)log_m; d.log_ms = (int)log_ms; d.g = g; d.rank = rho; d.n_layers = (int)ctx->sh.n_layers;
    d.sharded_mask = 0;
    for (int i = 0; i < 32; i++) {
        d.word_off[i] = ctx->sh.word_off[i];
        d.tree_off[i] = ctx->sh.tree_off[i];
        if (i < (int)ctx->sh.n_layers && ctx->sh.layer_sharded[i]) d.sharded_mask |= 1u << i;
    }
    shard::gather_share_kernel<<<(unsigned)nq, 256, 0, st>>>(d, ctx->ws_qidx.p, ctx->ws_sh_q.p, wpq);
    LAUNCHED(ctx);
    CK(cudaGetLastError());
    if (!(comm->flags & PB_COMM_STREAM_ORDERED)) CK(cudaStreamSynchronize(st));
    if (comm->all_gather(comm->user, ctx->ws_sh_q.p, ctx->ws_sh_qall.p, 4 * wpq * nq)) return PB_ERR_COMM;
    shard::sum_shares_kernel<<<(unsigned)((wpq * nq + 255) / 256), 256, 0, st>>>(ctx->ws_sh_qall.p, G, wpq * nq, ctx->ws_qout.p);
    LAUNCHED(ctx);
    CK(cudaMemcpyAsync(h_out, ctx->ws_qout.p, 4 * wpq * nq, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // the top g levels of every path: through the subtree roots gathered at commit time
    for (size_t q = 0; q < nq; q++) {
        uint32_t* o = h_out + q * wpq + 1;
        const size_t r = idx[q];
        const int blk = (int)(r >> log_ms);
        top_path(ctx->p2, ctx->sh.roots_main, G, blk, o + width + 8 * log_ms);
        o += width + 8 * log_m;
        if (wp) {
            top_path(ctx->p2, ctx->sh.roots_perm, G, blk, o + wp + 8 * log_ms);
            o += wp + 8 * log_m;
        }
        top_path(ctx->p2, ctx->sh.roots_q, G, blk, o + 8 + 8 * log_ms);
        o += 8 + 8 * log_m;
        for (uint32_t i = 0; i < ctx->sh.n_layers; i++) {
            const size_t log_h = log_m - 1 - i, j = r >> (i + 1);
            if (ctx->sh.layer_sharded[i]) {
                const size_t log_loc = log_h - (size_t)g;
                top_path(ctx->p2, ctx->sh.roots_fri[i], G, (int)(j >> log_loc), o + 8 + 8 * log_loc);
            }
            o += 8 + 8 * log_h;
        }
    }
    return 0;
}
