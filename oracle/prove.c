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

void orc_prove_segment(const uint32_t* trace, unsigned log_n, size_t width, const uint32_t* bc, const orc_span_t* spans,
                       size_t n_constraints, orc_segment_proof_t* proof, double st[8]) {
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
    orc_merkle_commit(mats2, w2, 2, log_m, layers);
    memcpy(proof->quotient_root, layers + 8 * (2 * m - 2), 32);
    free(layers);
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
    free(ys);
    free(lde);
    free(qlde);
    unsigned log_len = log_m;
    uint32_t shift = BB_GENERATOR;
    uint32_t layer_i = 0;
    while (log_len > log_blowup) {                       /* final_poly_len = 1 */
        merkle_root_rowmajor(f, 8, log_len - 1, proof->fri_roots[layer_i]);
        orc_challenger_observe(&ch, proof->fri_roots[layer_i], 8);
        orc_challenger_sample_ext(&ch, proof->fri_betas[layer_i]);
        uint32_t* nf = (uint32_t*)malloc(4 * ((size_t)1 << (log_len - 1)) * sizeof(uint32_t));
        orc_fri_fold(f, log_len, shift, proof->fri_betas[layer_i], nf);
        free(f);
        f = nf;
        shift = bb_mul(shift, shift);
        log_len--;
        layer_i++;
    }
    proof->n_fri_layers = layer_i;
    proof->final_len = 1u << log_len;
    memcpy(proof->final_poly, f, 16 * proof->final_len);
    free(f);
    double t6 = now_s();
    st[0] = t1 - t0; st[1] = t2 - t1; st[2] = t3 - t2; st[3] = t4 - t3; st[4] = t5 - t4; st[5] = t6 - t5;
}
