/*
 * ORACLE (test infrastructure, not product code): one APC segment through the three north-star stages,
 * in the order the V1-shaped prover runs them (metric names from
 * /root/reference/openvm/metrics-viewer/CLAUDE.md:52-78: main_trace_commit, quotient_poly_compute,
 * quotient_poly_commit, FRI commit phase of pcs_opening).  Call site being restated:
 * engine.prove(pk, ProvingContext) behind sdk.app_prover(exe)?.prove(stdin)
 * (/root/reference/openvm-riscv/src/lib.rs:327-332); the engine itself is un-vendored => parity unpinned.
 *
 * Transcript (simplified, documented in DESIGN.md): observe(trace_root) -> alpha; observe(quotient_root) -> zeta;
 * open every trace column and quotient-chunk column at zeta, observe(Merkle root of the opened values) -> gamma;
 * FRI input = reduced opening sum_j gamma^j (f_j(x) - f_j(zeta))/(x - zeta); per layer observe(root_i) -> beta_i.
 */
#include "oracle.h"
#include "bb31.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* Merkle root of a row-major matrix (FRI layer: rows = (lo, hi) pairs, 8 base elements) */
static void merkle_root_rowmajor(const uint32_t* mat, size_t width, unsigned log_h, uint32_t root[8]) {
    size_t h = (size_t)1 << log_h;
    uint32_t* layer = (uint32_t*)malloc(8 * h * sizeof(uint32_t));
#pragma omp parallel for schedule(static) if (h >= 2048)
    for (long r = 0; r < (long)h; r++) orc_hash_row(mat + (size_t)r * width, width, layer + 8 * (size_t)r);
    for (size_t n = h >> 1; n >= 1; n >>= 1) {
        uint32_t* next = (uint32_t*)malloc(8 * n * sizeof(uint32_t));
#pragma omp parallel for schedule(static) if (n >= 2048)
        for (long j = 0; j < (long)n; j++) orc_compress(layer + 16 * (size_t)j, layer + 16 * (size_t)j + 8, next + 8 * (size_t)j);
        free(layer);
        layer = next;
    }
    memcpy(root, layer, 32);
    free(layer);
}

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

void orc_prove_segment(const uint32_t* trace, unsigned log_n, size_t width, const uint32_t* bc, const orc_span_t* spans,
                       size_t n_constraints, orc_segment_proof_t* proof, double st[8]) {
    orc_prove_segment_q(trace, log_n, width, bc, spans, n_constraints, proof, st, NULL, 0, NULL);
}

/* same, and additionally the opened values (ys_out, [(width+8)][4]) and n_queries query openings in the layout of
   pb_query_segment (see include/powdr_b200.h) */
void orc_prove_segment_q(const uint32_t* trace, unsigned log_n, size_t width, const uint32_t* bc, const orc_span_t* spans,
                         size_t n_constraints, orc_segment_proof_t* proof, double st[8], uint32_t* ys_out, size_t n_queries,
                         uint32_t* queries_out) {
    const unsigned log_blowup = 1;
    size_t n = (size_t)1 << log_n, m = n << log_blowup;
    unsigned log_m = log_n + log_blowup;
    memset(proof, 0, sizeof *proof);
    memset(st, 0, 8 * sizeof(double));
    orc_challenger_t ch;
    orc_challenger_init(&ch);

    /* stage 1 + 3a: main trace commit */
    double t0 = now_s();
    uint32_t* lde = (uint32_t*)malloc(width * m * sizeof(uint32_t));
    orc_lde_batch(trace, log_n, width, log_blowup, BB_GENERATOR, lde);
    double t1 = now_s();
    uint32_t* layers = (uint32_t*)malloc((2 * m) * 8 * sizeof(uint32_t));
    const uint32_t* mats1[1] = {lde};
    orc_merkle_commit(mats1, &width, 1, log_m, layers);
    memcpy(proof->trace_root, layers + 8 * (2 * m - 2), 32);
    double t2 = now_s();
    orc_challenger_observe(&ch, proof->trace_root, 8);
    orc_challenger_sample_ext(&ch, proof->alpha);

    /* stage 2: quotient */
    uint32_t* q = (uint32_t*)malloc(8 * n * sizeof(uint32_t));
    orc_quotient(bc, spans, n_constraints, lde, log_n, log_blowup, BB_GENERATOR, proof->alpha, q);
    double t3 = now_s();

    /* quotient commit: chunk b holds evals over g*w_{2N}^b*H in bit-reversed order -> natural, LDE with shift g/s_b */
    uint32_t* qlde = (uint32_t*)malloc(8 * m * sizeof(uint32_t));
    uint32_t w2n = bb_root_of_unity(log_n + 1);
    uint32_t* nat = (uint32_t*)malloc(8 * n * sizeof(uint32_t));
    for (int b = 0; b < 2; b++) {
        for (int l = 0; l < 4; l++)
            for (size_t j = 0; j < n; j++) nat[((size_t)b * 4 + l) * n + bitrev32((uint32_t)j, log_n)] = q[((size_t)b * 4 + l) * n + j];
        uint32_t shift = b ? bb_inv(w2n) : 1;
        orc_lde_batch(nat + (size_t)b * 4 * n, log_n, 4, log_blowup, shift, qlde + (size_t)b * 4 * m);
    }
    free(q);
    double t4 = now_s();
    const uint32_t* mats2[2] = {qlde, qlde + 4 * m};
    size_t w2[2] = {4, 4};
    uint32_t* layers_q = (uint32_t*)malloc((2 * m) * 8 * sizeof(uint32_t));
    orc_merkle_commit(mats2, w2, 2, log_m, layers_q);
    memcpy(proof->quotient_root, layers_q + 8 * (2 * m - 2), 32);
    double t5 = now_s();
    orc_challenger_observe(&ch, proof->quotient_root, 8);
    orc_challenger_sample_ext(&ch, proof->zeta);

    /* openings at zeta: trace columns over H, quotient chunk b over g*w_{2N}^b*H; committed as rows of 8 and observed */
    size_t n_open = width + 8;
    size_t open_words = 4 * n_open, open_rows = 1;
    while (open_rows * 8 < open_words) open_rows <<= 1;
    unsigned log_open_rows = 0;
    while (((size_t)1 << log_open_rows) < open_rows) log_open_rows++;
    uint32_t* ys = (uint32_t*)calloc(open_rows * 8, sizeof(uint32_t));
    orc_eval_at_point(trace, log_n, width, 1, proof->zeta, ys);
    orc_eval_at_point(nat, log_n, 4, BB_GENERATOR, proof->zeta, ys + 4 * width);
    orc_eval_at_point(nat + 4 * n, log_n, 4, bb_mul(BB_GENERATOR, w2n), proof->zeta, ys + 4 * (width + 4));
    free(nat);
    merkle_root_rowmajor(ys, 8, log_open_rows, proof->openings_root);
    orc_challenger_observe(&ch, proof->openings_root, 8);
    orc_challenger_sample_ext(&ch, proof->gamma);

    /* stage 3b: FRI commit phase on the reduced opening ro(x) = sum_j gamma^j (f_j(x) - f_j(zeta)) / (x - zeta) over g*H' */
    uint32_t* f = (uint32_t*)malloc(4 * m * sizeof(uint32_t));
    {
        const uint32_t* mats3[3] = {lde, qlde, qlde + 4 * m};
        size_t w3[3] = {width, 4, 4};
        orc_deep_quotient(mats3, w3, 3, log_m, BB_GENERATOR, proof->zeta, proof->gamma, ys, f);
    }
    if (ys_out) memcpy(ys_out, ys, 16 * n_open);
    free(ys);
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
    /* query phase */
    if (n_queries && queries_out) {
        size_t wpq = 1 + width + 8 * log_m + 8 + 8 * log_m;
        for (unsigned i = 0; i < layer_i; i++) wpq += 8 + 8 * (log_m - 1 - i);
        for (size_t qi = 0; qi < n_queries; qi++) {
            uint32_t* o = queries_out + qi * wpq;
            size_t r = orc_challenger_sample(&ch) & (((size_t)1 << log_m) - 1);
            *o++ = (uint32_t)r;
            for (size_t c = 0; c < width; c++) o[c] = lde[c * m + r];
            o += width;
            copy_path(layers, log_m, r, o);
            o += 8 * log_m;
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
    for (unsigned i = 0; i < layer_i; i++) { free(words[i]); free(trees[i]); }
    free(lde);
    free(qlde);
    free(layers);
    free(layers_q);
    double t6 = now_s();
    st[0] = t1 - t0; st[1] = t2 - t1; st[2] = t3 - t2; st[3] = t4 - t3; st[4] = t5 - t4; st[5] = t6 - t5;
}
