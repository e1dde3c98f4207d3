/*
 * ORACLE (test infrastructure, not product code): an independent VERIFIER for the segment proof object.
 * The reference's own tests pin the prover only by validity -- a proof produced by engine E must verify under the CPU
 * engine (verify_app_proof::<BabyBearPoseidon2CpuEngine>, /root/reference/openvm-riscv/src/lib.rs:337-341).
 * This file plays that role for the transcript of DESIGN.md §3: it re-derives every challenge, recomputes the reduced
 * opening from the opened rows, checks every Merkle path, every FRI fold and the final polynomial, and (optionally, for
 * satisfying traces) the constraint identity  sum_k alpha^(C-1-k) c_k(T(zeta)) = Z_H(zeta) * Q(zeta).
 * It shares no state with the prover: inputs are the proof struct, the opened values and the query openings.
 */
#include "oracle.h"
#include "bb31.h"
#include <string.h>
#include <stdlib.h>

static bb4_t ld4(const uint32_t* p) { bb4_t r; memcpy(r.c, p, 16); return r; }
static int eq4(bb4_t a, bb4_t b) { return memcmp(a.c, b.c, 16) == 0; }

static int check_path(const uint32_t leaf[8], size_t idx, const uint32_t* path, unsigned log_h, const uint32_t root[8]) {
    uint32_t node[8], next[8];
    memcpy(node, leaf, 32);
    for (unsigned k = 0; k < log_h; k++) {
        const uint32_t* sib = path + 8 * k;
        if ((idx >> k) & 1) orc_compress(sib, node, next); else orc_compress(node, sib, next);
        memcpy(node, next, 32);
    }
    return memcmp(node, root, 32) == 0;
}

/* bytecode over Ext4 values (column j -> vals[j]); same opcodes as orc_eval_expr */
static bb4_t eval_ext(const uint32_t* bc, uint32_t len, const uint32_t* vals) {
    bb4_t st[16];
    int sp = 0;
    for (uint32_t ip = 0; ip < len;) {
        uint32_t op = bc[ip++];
        switch (op) {
        case 0: st[sp++] = ld4(vals + 4 * bc[ip++]); break;
        case 1: st[sp++] = bb4_from_base(bc[ip++] % BB_P); break;
        case 2: sp--; st[sp - 1] = bb4_add(st[sp - 1], st[sp]); break;
        case 3: sp--; st[sp - 1] = bb4_sub(st[sp - 1], st[sp]); break;
        case 4: sp--; st[sp - 1] = bb4_mul(st[sp - 1], st[sp]); break;
        case 5: st[sp - 1] = bb4_sub(bb4_from_base(0), st[sp - 1]); break;
        default: { bb4_t z = bb4_from_base(0); st[sp - 1] = eq4(st[sp - 1], z) ? z : bb4_inv(st[sp - 1]); break; }
        }
    }
    return st[0];
}

int orc_verify_segment(const uint32_t* bc, const orc_span_t* spans, size_t n_constraints, unsigned log_n, size_t width,
                       const orc_segment_proof_t* proof, const uint32_t* ys, const uint32_t* queries, size_t n_queries,
                       int check_constraints) {
    const unsigned log_m = log_n + 1;
    const size_t n = (size_t)1 << log_n, n_open = width + 8;
    if (proof->n_fri_layers != log_n || proof->final_len != 2) return 20;

    /* 1. transcript */
    orc_challenger_t ch;
    orc_challenger_init(&ch);
    uint32_t t4[4];
    orc_challenger_observe(&ch, proof->trace_root, 8);
    orc_challenger_sample_ext(&ch, t4);
    if (memcmp(t4, proof->alpha, 16)) return 1;
    orc_challenger_observe(&ch, proof->quotient_root, 8);
    orc_challenger_sample_ext(&ch, t4);
    if (memcmp(t4, proof->zeta, 16)) return 2;
    {
        size_t rows = 1;
        unsigned lr = 0;
        while (rows * 8 < 4 * n_open) { rows <<= 1; lr++; }
        uint32_t* buf = (uint32_t*)calloc(rows * 8, 4);
        memcpy(buf, ys, 16 * n_open);
        uint32_t* layer = (uint32_t*)malloc(8 * rows * 4);
        for (size_t r = 0; r < rows; r++) orc_hash_row(buf + 8 * r, 8, layer + 8 * r);
        for (size_t m = rows >> 1; m >= 1; m >>= 1) {
            for (size_t j = 0; j < m; j++) { uint32_t o[8]; orc_compress(layer + 16 * j, layer + 16 * j + 8, o); memcpy(layer + 8 * j, o, 32); }
            if (m == 1) break;
        }
        int ok = memcmp(layer, proof->openings_root, 32) == 0;
        free(buf);
        free(layer);
        if (!ok) return 3;
    }
    orc_challenger_observe(&ch, proof->openings_root, 8);
    orc_challenger_sample_ext(&ch, t4);
    if (memcmp(t4, proof->gamma, 16)) return 3;
    for (uint32_t i = 0; i < proof->n_fri_layers; i++) {
        orc_challenger_observe(&ch, proof->fri_roots[i], 8);
        orc_challenger_sample_ext(&ch, t4);
        if (memcmp(t4, proof->fri_betas[i], 16)) return 4;
    }

    const bb4_t zeta = ld4(proof->zeta), gamma = ld4(proof->gamma), alpha = ld4(proof->alpha);

    /* 2. constraint identity at zeta (only meaningful for a satisfying trace) */
    if (check_constraints) {
        bb4_t acc = bb4_from_base(0);
        for (size_t k = 0; k < n_constraints; k++) {
            acc = bb4_mul(acc, alpha);
            acc = bb4_add(acc, eval_ext(bc + spans[k].off, spans[k].len, ys));
        }
        bb4_t zn = bb4_pow(zeta, n);
        uint32_t gn = bb_pow(BB_GENERATOR, n);
        /* Q(zeta) = sum_b (zeta^N - g^N (-1)^(1-b)) / (2 g^N (-1)^b) * Q_b(zeta),  Q_b = sum_l x^l-basis limb l of chunk b */
        bb4_t q = bb4_from_base(0);
        for (int b = 0; b < 2; b++) {
            uint32_t sgn_b = b ? BB_P - 1 : 1, sgn_nb = b ? 1 : BB_P - 1;
            bb4_t num = zn;
            num.c[0] = bb_sub(num.c[0], bb_mul(gn, sgn_nb));
            uint32_t den = bb_inv(bb_mul(2, bb_mul(gn, sgn_b)));
            bb4_t qb = bb4_from_base(0);
            for (int l = 0; l < 4; l++) {
                bb4_t e = bb4_from_base(0);
                e.c[l] = 1;                                             /* basis element x^l of Ext4 */
                qb = bb4_add(qb, bb4_mul(e, ld4(ys + 4 * (width + 4 * b + l))));
            }
            q = bb4_add(q, bb4_mul(bb4_scale(num, den), qb));
        }
        bb4_t zh = zn;
        zh.c[0] = bb_sub(zh.c[0], 1);
        if (!eq4(acc, bb4_mul(zh, q))) return 12;
    }

    /* 3. queries */
    size_t wpq = 1 + width + 8 * log_m + 8 + 8 * log_m;
    for (unsigned i = 0; i < log_n; i++) wpq += 8 + 8 * (log_m - 1 - i);
    bb4_t* gp = (bb4_t*)malloc(n_open * sizeof(bb4_t));
    bb4_t cur = bb4_from_base(1), ysum = bb4_from_base(0);
    for (size_t j = 0; j < n_open; j++) {
        gp[j] = cur;
        ysum = bb4_add(ysum, bb4_mul(cur, ld4(ys + 4 * j)));
        cur = bb4_mul(cur, gamma);
    }
    const uint32_t two_inv = bb_inv(2);
    int rc = 0;
    for (size_t qi = 0; qi < n_queries && !rc; qi++) {
        const uint32_t* o = queries + qi * wpq;
        const size_t r = o[0];
        if (r != (orc_challenger_sample(&ch) & (((size_t)1 << log_m) - 1))) { rc = 5; break; }
        const uint32_t* trow = o + 1;
        const uint32_t* tpath = trow + width;
        const uint32_t* qrow = tpath + 8 * log_m;
        const uint32_t* qpath = qrow + 8;
        const uint32_t* fr = qpath + 8 * log_m;
        uint32_t leaf[8];
        orc_hash_row(trow, width, leaf);
        if (!check_path(leaf, r, tpath, log_m, proof->trace_root)) { rc = 6; break; }
        orc_hash_row(qrow, 8, leaf);
        if (!check_path(leaf, r, qpath, log_m, proof->quotient_root)) { rc = 7; break; }
        bb4_t acc = bb4_from_base(0);
        for (size_t j = 0; j < width; j++) acc = bb4_add(acc, bb4_scale(gp[j], trow[j]));
        for (size_t j = 0; j < 8; j++) acc = bb4_add(acc, bb4_scale(gp[width + j], qrow[j]));
        acc = bb4_sub(acc, ysum);
        uint32_t x = bb_mul(BB_GENERATOR, bb_pow(bb_root_of_unity(log_m), bitrev32((uint32_t)r, log_m)));
        bb4_t d = bb4_from_base(x);
        d = bb4_sub(d, zeta);
        bb4_t val = bb4_mul(acc, bb4_inv(d));
        size_t idx = r;
        uint32_t shift = BB_GENERATOR;
        for (unsigned i = 0; i < log_n; i++) {
            const unsigned log_len = log_m - i, log_h = log_len - 1;
            const uint32_t* pair = fr;
            const uint32_t* path = fr + 8;
            fr += 8 + 8 * log_h;
            bb4_t lo = ld4(pair), hi = ld4(pair + 4);
            if (!eq4((idx & 1) ? hi : lo, val)) { rc = i == 0 ? 8 : 10; break; }
            const size_t j = idx >> 1;
            orc_hash_row(pair, 8, leaf);
            if (!check_path(leaf, j, path, log_h, proof->fri_roots[i])) { rc = 9; break; }
            uint32_t xj = bb_mul(shift, bb_pow(bb_root_of_unity(log_len), bitrev32((uint32_t)j, log_h)));
            bb4_t beta = ld4(proof->fri_betas[i]);
            bb4_t s = bb4_scale(bb4_add(lo, hi), two_inv);
            bb4_t df = bb4_scale(bb4_sub(lo, hi), bb_mul(two_inv, bb_inv(xj)));
            val = bb4_add(s, bb4_mul(beta, df));
            idx = j;
            shift = bb_mul(shift, shift);
        }
        if (!rc && !eq4(val, ld4(proof->final_poly[idx]))) rc = 11;
    }
    free(gp);
    return rc;
}
