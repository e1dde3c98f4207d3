/*
 * ORACLE (test infrastructure, not product code): one APC segment (one chip) through the whole path, in the order the
 * V1-shaped FRI prover runs it (metric names from /root/reference/openvm/metrics-viewer/CLAUDE.md:52-78: main_trace_commit,
 * perm_trace_commit, quotient_poly_compute, quotient_poly_commit, pcs_opening).  Call site being restated:
 * engine.prove(pk, ProvingContext) behind sdk.app_prover(exe)?.prove(stdin) (/root/reference/openvm-riscv/src/lib.rs:327-332);
 * the engine itself is un-vendored => parity unpinned (oracle.h).
 *
 * Transcript v2 (DuplexChallenger):
 *   observe(trace_root)
 *   [interactions only]  alpha_lu, beta_lu <- sample_ext x2;  observe(perm_root);  observe(cumulative_sum)
 *   alpha <- sample_ext;  observe(quotient_root);  zeta <- sample_ext
 *   observe every opened value: main at zeta | perm at zeta | perm at zeta*w | quotient chunks at zeta;  gamma <- sample_ext
 *   per FRI layer: observe(root_i), beta_i <- sample_ext;  observe(final polynomial constant)
 *   proof of work: observe(witness), sample_bits(pow_bits) == 0;  n_queries x sample_bits(log_m)
 * The heavy primitives run either on the scalar `%` restatement or (prm->fast) on the AVX-512 Montgomery ones of fast.c,
 * which are checked bit for bit against the former -- the proof is the same either way.
 */
#include "oracle.h"
#include "bb31.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* big working buffers: 2 MB aligned and advised to transparent huge pages -- the per-row kernels walk thousands of columns that are
   megabytes apart, which thrashes the TLB with 4 KB pages */
#include <sys/mman.h>
void* orc_big_alloc(size_t bytes) {
    const size_t al = (size_t)2 << 20;
    bytes = (bytes + al - 1) / al * al;
    void* p = aligned_alloc(al, bytes ? bytes : al);
#ifdef MADV_HUGEPAGE
    if (p) madvise(p, bytes ? bytes : al, MADV_HUGEPAGE);
#endif
    return p;
}
void orc_big_free(void* p) { free(p); }
#define BIG(n_words) ((uint32_t*)orc_big_alloc((size_t)(n_words) * sizeof(uint32_t)))

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct {
    void (*lde_batch)(const uint32_t*, unsigned, size_t, unsigned, uint32_t, uint32_t*);
    void (*merkle_commit)(const uint32_t* const*, const size_t*, size_t, unsigned, uint32_t*);
    void (*constraint_fold)(const uint32_t*, const orc_span_t*, size_t, const uint32_t*, size_t, const uint32_t[4], uint32_t*);
    void (*eval_at_point)(const uint32_t*, unsigned, size_t, uint32_t, const uint32_t[4], uint32_t*);
    void (*deep_groups)(const uint32_t* const*, const uint32_t*, size_t, const uint32_t*, size_t, unsigned, uint32_t, const uint32_t[4],
                        const uint32_t*, uint32_t*);
    void (*logup_perm_trace)(const uint32_t*, unsigned, const uint32_t*, const orc_span_t*, const orc_interaction_t*, size_t, const uint32_t*, size_t,
                             const uint32_t[4], const uint32_t[4], uint32_t*, uint32_t[4]);
    void (*logup_fold)(const uint32_t*, const uint32_t*, unsigned, uint32_t, const uint32_t*, const orc_span_t*, const orc_interaction_t*, size_t,
                       const uint32_t*, size_t, const uint32_t[4], const uint32_t[4], const uint32_t[4], const uint32_t[4], uint32_t*);
} ops_t;
static const ops_t OPS_SLOW = {orc_lde_batch, orc_merkle_commit, orc_constraint_fold, orc_eval_at_point, orc_deep_quotient_groups,
                               orc_logup_perm_trace, orc_logup_fold};
static const ops_t OPS_FAST = {orcf_lde_batch, orcf_merkle_commit, orcf_constraint_fold, orcf_eval_at_point, orcf_deep_quotient_groups,
                               orcf_logup_perm_trace, orcf_logup_fold};

/* Merkle tree (all layers, node-major) of a row-major matrix; returns malloc'ed [2h-1][8] */
static uint32_t* merkle_tree_rowmajor(const uint32_t* mat, size_t width, unsigned log_h) {
    size_t h = (size_t)1 << log_h;
    uint32_t* t = (uint32_t*)malloc(8 * (2 * h) * sizeof(uint32_t));
#pragma omp parallel for schedule(static) if (h >= 2048)
    for (long r = 0; r < (long)h; r++) orc_hash_row(mat + (size_t)r * width, width, t + 8 * (size_t)r);
    uint32_t* prev = t;
    for (size_t n = h >> 1; n >= 1; n >>= 1) {
        uint32_t* cur = prev + 16 * n;
#pragma omp parallel for schedule(static) if (n >= 2048)
        for (long j = 0; j < (long)n; j++) orc_compress(prev + 16 * (size_t)j, prev + 16 * (size_t)j + 8, cur + 8 * (size_t)j);
        prev = cur;
        if (n == 1) break;
    }
    return t;
}

static void copy_path(const uint32_t* tree, unsigned log_h, size_t idx, uint32_t* out) {
    for (unsigned k = 0; k < log_h; k++) {
        size_t level_off = ((size_t)2 << log_h) - ((size_t)2 << (log_h - k));
        memcpy(out + 8 * k, tree + 8 * (level_off + ((idx >> k) ^ 1)), 32);
    }
}

size_t orc_perm_width(const orc_air_t* air) {
    if (!air->n_ints) return 0;
    uint32_t* cs = (uint32_t*)malloc((air->n_ints + 1) * sizeof(uint32_t));
    int nc = orc_logup_chunks(air->ibc, air->ispans, air->ints, air->n_ints, 3, cs);
    free(cs);
    return nc < 0 ? 0 : 4 * ((size_t)nc + 1);
}
size_t orc_num_opened(size_t width, size_t perm_width) { return width + 2 * perm_width + 8; }
size_t orc_query_words(unsigned log_n, size_t width, size_t perm_width) {
    const unsigned log_m = log_n + 1;
    size_t w = 1 + width + 8 * log_m + (perm_width ? perm_width + 8 * log_m : 0) + 8 + 8 * log_m;
    for (unsigned i = 0; i < log_n; i++) w += 8 + 8 * (log_m - 1 - i);
    return w;
}

uint32_t orc_grind(const orc_challenger_t* c, unsigned bits) {
    const uint32_t mask = bits >= 31 ? 0x7fffffffu : ((1u << bits) - 1);
    for (uint32_t base = 0;; base += 4096) {
        uint32_t found = 0xffffffffu;
#pragma omp parallel for schedule(static)
        for (long i = 0; i < 4096; i++) {
            const uint32_t w = base + (uint32_t)i;
            if (w >= BB_P) continue;
            orc_challenger_t t = *c;
            orc_challenger_observe(&t, &w, 1);
            if ((orc_challenger_sample(&t) & mask) == 0) {
#pragma omp critical
                if (w < found) found = w;
            }
        }
        if (found != 0xffffffffu) return found;
    }
}

void orc_prove_segment(const uint32_t* trace, unsigned log_n, size_t width, const orc_air_t* air, const orc_params_t* prm,
                       orc_segment_proof_t* proof, double st[10], uint32_t* ys_out, uint32_t* queries_out) {
    const ops_t* ops = (prm->fast && orcf_available()) ? &OPS_FAST : &OPS_SLOW;
    const unsigned log_blowup = 1;
    const size_t n = (size_t)1 << log_n, m = n << log_blowup;
    const unsigned log_m = log_n + log_blowup;
    memset(proof, 0, sizeof *proof);
    memset(st, 0, 10 * sizeof(double));
    proof->pow_bits = prm->pow_bits;
    proof->n_queries = prm->n_queries;
    orc_challenger_t ch;
    orc_challenger_init(&ch);

    /* stage 1 + 3a: main trace commit */
    double t0 = now_s();
    uint32_t* lde = BIG(width * m);
    ops->lde_batch(trace, log_n, width, log_blowup, BB_GENERATOR, lde);
    double t1 = now_s();
    st[0] = t1 - t0;
    uint32_t* layers = BIG((2 * m) * 8);
    const uint32_t* mats1[1] = {lde};
    ops->merkle_commit(mats1, &width, 1, log_m, layers);
    memcpy(proof->trace_root, layers + 8 * (2 * m - 2), 32);
    orc_challenger_observe(&ch, proof->trace_root, 8);
    double t2 = now_s();
    st[1] = t2 - t1;

    /* LogUp: permutation trace, its commitment, the exposed cumulative sum */
    size_t wp = 0, n_chunks = 0;
    uint32_t* chunk_start = NULL;
    uint32_t *perm = NULL, *perm_lde = NULL, *layers_p = NULL;
    if (air->n_ints) {
        chunk_start = (uint32_t*)malloc((air->n_ints + 1) * sizeof(uint32_t));
        int nc = orc_logup_chunks(air->ibc, air->ispans, air->ints, air->n_ints, 3, chunk_start);
        if (nc < 0) abort();
        n_chunks = (size_t)nc;
        wp = 4 * (n_chunks + 1);
        orc_challenger_sample_ext(&ch, proof->logup_alpha);
        orc_challenger_sample_ext(&ch, proof->logup_beta);
        perm = BIG(wp * n);
        ops->logup_perm_trace(trace, log_n, air->ibc, air->ispans, air->ints, air->n_ints, chunk_start, n_chunks, proof->logup_alpha,
                             proof->logup_beta, perm, proof->cumulative_sum);
        double t3 = now_s();
        st[2] = t3 - t2;
        perm_lde = BIG(wp * m);
        ops->lde_batch(perm, log_n, wp, log_blowup, BB_GENERATOR, perm_lde);
        layers_p = BIG((2 * m) * 8);
        const uint32_t* matsp[1] = {perm_lde};
        ops->merkle_commit(matsp, &wp, 1, log_m, layers_p);
        memcpy(proof->perm_root, layers_p + 8 * (2 * m - 2), 32);
        orc_challenger_observe(&ch, proof->perm_root, 8);
        orc_challenger_observe(&ch, proof->cumulative_sum, 4);
        st[3] = now_s() - t3;
    }
    proof->perm_width = (uint32_t)wp;
    orc_challenger_sample_ext(&ch, proof->alpha);

    /* stage 2: quotient = (Horner fold of the AIR's constraints, then of the LogUp constraints) / Z_H, split by natural-index parity */
    double t4 = now_s();
    uint32_t* q = (uint32_t*)malloc(8 * n * sizeof(uint32_t));
    {
        uint32_t* acc4 = BIG(4 * m);
        ops->constraint_fold(air->bc, air->spans, air->n_constraints, lde, m, proof->alpha, acc4);
        if (air->n_ints)
            ops->logup_fold(lde, perm_lde, log_n, BB_GENERATOR, air->ibc, air->ispans, air->ints, air->n_ints, chunk_start, n_chunks,
                           proof->logup_alpha, proof->logup_beta, proof->cumulative_sum, proof->alpha, acc4);
        uint32_t sn = bb_pow(BB_GENERATOR, n);
        uint32_t zinv[2] = {bb_inv(bb_sub(sn, 1)), bb_inv(bb_sub(bb_neg(sn), 1))};
#pragma omp parallel for schedule(static)
        for (long r = 0; r < (long)m; r++) {
            size_t chunk = (size_t)r >> log_n, j = (size_t)r & (n - 1);
            for (int l = 0; l < 4; l++) q[(chunk * 4 + l) * n + j] = bb_mul(acc4[(size_t)l * m + r], zinv[chunk]);
        }
        free(acc4);
    }
    double t5 = now_s();
    st[4] = t5 - t4;

    /* quotient commit: chunk b holds evals over g*w_{2N}^b*H in bit-reversed order -> natural, LDE with shift g/s_b */
    uint32_t* qlde = BIG(8 * m);
    uint32_t w2n = bb_root_of_unity(log_n + 1);
    uint32_t* nat = (uint32_t*)malloc(8 * n * sizeof(uint32_t));
    for (int b = 0; b < 2; b++) {
        for (int l = 0; l < 4; l++)
            for (size_t j = 0; j < n; j++) nat[((size_t)b * 4 + l) * n + bitrev32((uint32_t)j, log_n)] = q[((size_t)b * 4 + l) * n + j];
        uint32_t shift = b ? bb_inv(w2n) : 1;
        ops->lde_batch(nat + (size_t)b * 4 * n, log_n, 4, log_blowup, shift, qlde + (size_t)b * 4 * m);
    }
    free(q);
    const uint32_t* mats2[2] = {qlde, qlde + 4 * m};
    size_t w2[2] = {4, 4};
    uint32_t* layers_q = BIG((2 * m) * 8);
    ops->merkle_commit(mats2, w2, 2, log_m, layers_q);
    memcpy(proof->quotient_root, layers_q + 8 * (2 * m - 2), 32);
    orc_challenger_observe(&ch, proof->quotient_root, 8);
    orc_challenger_sample_ext(&ch, proof->zeta);
    double t6 = now_s();
    st[5] = t6 - t5;

    /* openings: main at zeta, perm at zeta and zeta*w (the transition constraint reads the next row), quotient chunks at zeta */
    const size_t n_open = orc_num_opened(width, wp);
    uint32_t* ys = (uint32_t*)calloc(n_open * 4, sizeof(uint32_t));
    uint32_t zeta_next[4];
    {
        bb4_t z = {{proof->zeta[0], proof->zeta[1], proof->zeta[2], proof->zeta[3]}};
        z = bb4_scale(z, bb_root_of_unity(log_n));
        memcpy(zeta_next, z.c, 16);
    }
    ops->eval_at_point(trace, log_n, width, 1, proof->zeta, ys);
    if (wp) {
        ops->eval_at_point(perm, log_n, wp, 1, proof->zeta, ys + 4 * width);
        ops->eval_at_point(perm, log_n, wp, 1, zeta_next, ys + 4 * (width + wp));
    }
    ops->eval_at_point(nat, log_n, 4, BB_GENERATOR, proof->zeta, ys + 4 * (width + 2 * wp));
    ops->eval_at_point(nat + 4 * n, log_n, 4, bb_mul(BB_GENERATOR, w2n), proof->zeta, ys + 4 * (width + 2 * wp + 4));
    free(nat);
    if (prm->cheat_opening) ys[0] = bb_add(ys[0], 1);
    orc_challenger_observe(&ch, ys, 4 * n_open);
    orc_challenger_sample_ext(&ch, proof->gamma);

    /* stage 3b: FRI commit phase on the reduced opening over g*H' */
    uint32_t* f = BIG(4 * m);
    {
        const uint32_t** cols = (const uint32_t**)malloc(n_open * sizeof(*cols));
        uint32_t* grp = (uint32_t*)malloc(n_open * sizeof(uint32_t));
        size_t k = 0;
        for (size_t c = 0; c < width; c++) { cols[k] = lde + c * m; grp[k++] = 0; }
        for (size_t c = 0; c < wp; c++) { cols[k] = perm_lde + c * m; grp[k++] = 0; }
        for (size_t c = 0; c < wp; c++) { cols[k] = perm_lde + c * m; grp[k++] = 1; }
        for (size_t c = 0; c < 8; c++) { cols[k] = qlde + c * m; grp[k++] = 0; }
        uint32_t zs[8];
        memcpy(zs, proof->zeta, 16);
        memcpy(zs + 4, zeta_next, 16);
        ops->deep_groups(cols, grp, n_open, zs, wp ? 2 : 1, log_m, BB_GENERATOR, proof->gamma, ys, f);
        free(cols); free(grp);
    }
    if (ys_out) memcpy(ys_out, ys, 16 * n_open);
    free(ys);
    double t7 = now_s();
    st[6] = t7 - t6;
    unsigned log_len = log_m;
    uint32_t shift = BB_GENERATOR;
    uint32_t layer_i = 0;
    uint32_t* words[32];
    uint32_t* trees[32];
    while (log_len > log_blowup) {                       /* final_poly_len = 1 */
        words[layer_i] = f;
        trees[layer_i] = merkle_tree_rowmajor(f, 8, log_len - 1);
        memcpy(proof->fri_roots[layer_i], trees[layer_i] + 8 * (((size_t)2 << (log_len - 1)) - 2), 32);
        orc_challenger_observe(&ch, proof->fri_roots[layer_i], 8);
        orc_challenger_sample_ext(&ch, proof->fri_betas[layer_i]);
        uint32_t* nf = (uint32_t*)malloc(4 * ((size_t)1 << (log_len - 1)) * sizeof(uint32_t));
        orc_fri_fold(f, log_len, shift, proof->fri_betas[layer_i], nf);
        f = nf;
        shift = bb_mul(shift, shift);
        log_len--;
        layer_i++;
    }
    proof->n_fri_layers = layer_i;
    proof->final_len = 1u << log_len;
    memcpy(proof->final_poly, f, 16 * proof->final_len);
    free(f);
    orc_challenger_observe(&ch, proof->final_poly[0], 4);
    double t8 = now_s();
    st[7] = t8 - t7;

    /* proof of work, then the query phase */
    proof->pow_witness = orc_grind(&ch, prm->pow_bits);
    orc_challenger_observe(&ch, &proof->pow_witness, 1);
    (void)orc_challenger_sample(&ch);
    double t9 = now_s();
    st[8] = t9 - t8;
    if (prm->n_queries && queries_out) {
        const size_t wpq = orc_query_words(log_n, width, wp);
        for (size_t qi = 0; qi < prm->n_queries; qi++) {
            uint32_t* o = queries_out + qi * wpq;
            size_t r = orc_challenger_sample(&ch) & (((size_t)1 << log_m) - 1);
            *o++ = (uint32_t)r;
            for (size_t c = 0; c < width; c++) o[c] = lde[c * m + r];
            o += width;
            copy_path(layers, log_m, r, o);
            o += 8 * log_m;
            if (wp) {
                for (size_t c = 0; c < wp; c++) o[c] = perm_lde[c * m + r];
                o += wp;
                copy_path(layers_p, log_m, r, o);
                o += 8 * log_m;
            }
            for (size_t c = 0; c < 8; c++) o[c] = qlde[c * m + r];
            o += 8;
            copy_path(layers_q, log_m, r, o);
            o += 8 * log_m;
            for (unsigned i = 0; i < layer_i; i++) {
                unsigned lh = log_m - 1 - i;
                size_t j = r >> (i + 1);
                memcpy(o, words[i] + 8 * j, 32);
                o += 8;
                copy_path(trees[i], lh, j, o);
                o += 8 * lh;
            }
        }
    }
    st[9] = now_s() - t9;
    for (unsigned i = 0; i < layer_i; i++) { free(words[i]); free(trees[i]); }
    free(lde); free(qlde); free(layers); free(layers_q);
    free(perm); free(perm_lde); free(layers_p); free(chunk_start);
}
