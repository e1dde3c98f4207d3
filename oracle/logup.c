/*
 * ORACLE (test infrastructure, not product code): the LogUp / bus-interaction argument for one APC chip.
 *
 * What the reference pins: PowdrAir::eval pushes EVERY bus interaction of the machine on row_slice(0) with
 *   builder.push_interaction(bus_id, args, mult, count_weight = 1)
 * (/root/reference/openvm/src/powdr_extension/chip.rs:117-128) -- one signed multiplicity expression `mult` and a tuple of
 * argument expressions per interaction, bus index = SymbolicBusInteraction::id
 * (/root/reference/autoprecompiles/src/symbolic_machine.rs:70-75); value-level semantics of the three periphery buses at
 * /root/reference/openvm/src/powdr_extension/trace_generator/cpu/periphery.rs:179-236; degree bound of bus expressions = 2
 * (DegreeBound{identities:3,bus_interactions:2}, /root/reference/openvm/src/lib.rs:97-101).
 * What it does NOT hold: the argument itself (openvm-stark-backend, un-vendored).  Restated here is the FRI-LogUp phase of
 * stark-backend v1 (interaction/fri_log_up.rs), the prover generation the north-star's stage list describes -- PARITY UNPINNED:
 *   challenges   alpha_lu, beta_lu in Ext4, sampled after the main-trace commitment
 *   denominator  d_i(row) = alpha_lu + sum_{j<k} beta_lu^j * arg_j(row) + beta_lu^k * (bus_id + 1),  k = number of args
 *   chunks       consecutive interactions are grouped greedily while the chunk constraint below keeps degree <= 3
 *   perm trace   one Ext4 column per chunk  perm_c(row) = sum_{i in c} mult_i(row) / d_i(row),  plus the running sum
 *                phi(r) = sum_{r' <= r} sum_c perm_c(r');  width 4*(n_chunks + 1) base columns; cumulative_sum = phi(N-1) is exposed
 *   constraints  (appended to the alpha-fold after the AIR's own), Ext4-valued:
 *                  per chunk   perm_c * prod_i d_i - sum_i mult_i prod_{j != i} d_j = 0
 *                  first row   is_first * (phi - sum_c perm_c) = 0
 *                  transition  is_transition * (phi' - phi - sum_c perm_c') = 0        (' = next row)
 *                  last row    is_last * (phi - cumulative_sum) = 0
 *                with Plonky3's unnormalised coset selectors  is_first = Z_H(x)/(x-1), is_last = Z_H(x)/(x - w^-1),
 *                is_transition = x - w^-1.
 * Bytecode of the interactions: the reference's compile_bus_to_gpu layout (cuda/mod.rs:143-177: per interaction the spans
 * [mult, arg_0 .. arg_{k-1}] starting at args_index_off) with the PUSH_APC operand being the COLUMN INDEX.
 */
#include "oracle.h"
#include "bb31.h"
#include <assert.h>
#include <stdlib.h>
#include <string.h>

enum { OP_PUSH_APC = 0, OP_PUSH_CONST = 1, OP_ADD = 2, OP_SUB = 3, OP_MUL = 4, OP_NEG = 5, OP_INV_OR_ZERO = 6 };

static unsigned expr_degree(const uint32_t* bc, uint32_t len) {
    unsigned st[16];
    int sp = 0;
    for (uint32_t ip = 0; ip < len;) {
        switch (bc[ip++]) {
        case OP_PUSH_APC: ip++; st[sp++] = 1; break;
        case OP_PUSH_CONST: ip++; st[sp++] = 0; break;
        case OP_ADD: case OP_SUB: sp--; st[sp - 1] = st[sp - 1] > st[sp] ? st[sp - 1] : st[sp]; break;
        case OP_MUL: sp--; st[sp - 1] += st[sp]; break;
        case OP_NEG: break;
        default: st[sp - 1] = 99; break;       /* INV_OR_ZERO is not a polynomial: not allowed in an interaction */
        }
    }
    return st[0];
}

/* greedy chunking in declaration order; chunk_start must hold n_ints + 1 entries.  Returns n_chunks, or -1 when a single
   interaction already exceeds the degree bound. */
int orc_logup_chunks(const uint32_t* ibc, const orc_span_t* isp, const orc_interaction_t* ints, size_t n_ints, unsigned max_degree,
                     uint32_t* chunk_start) {
    int n_chunks = 0;
    size_t i = 0;
    while (i < n_ints) {
        chunk_start[n_chunks++] = (uint32_t)i;
        unsigned sum_dd = 0, max_extra = 0;      /* chunk degree = max(1 + sum_dd, max_i (md_i - dd_i) + sum_dd) */
        size_t cnt = 0;
        for (; i < n_ints; i++, cnt++) {
            const orc_span_t* s = isp + ints[i].args_index_off;
            unsigned md = expr_degree(ibc + s[0].off, s[0].len), dd = 0;
            for (uint32_t j = 0; j < ints[i].num_args; j++) {
                unsigned d = expr_degree(ibc + s[1 + j].off, s[1 + j].len);
                if (d > dd) dd = d;
            }
            const unsigned extra = md > dd ? md - dd : 0;
            const unsigned nsum = sum_dd + dd, nextra = extra > max_extra ? extra : max_extra;
            const unsigned deg = (1 > nextra ? 1 : nextra) + nsum;
            if (deg > max_degree) {
                if (cnt == 0) return -1;
                break;
            }
            sum_dd = nsum;
            max_extra = nextra;
        }
    }
    chunk_start[n_chunks] = (uint32_t)n_ints;
    return n_chunks;
}

/* (mult_i, d_i) of every interaction on one row of a column-major matrix (stride = its height) */
static void eval_interactions(const uint32_t* ibc, const orc_span_t* isp, const orc_interaction_t* ints, size_t n_ints, const uint32_t* mat,
                              size_t height, size_t r, bb4_t alpha, const bb4_t* betas, uint32_t* mult, bb4_t* den) {
    for (size_t i = 0; i < n_ints; i++) {
        const orc_span_t* s = isp + ints[i].args_index_off;
        /* PUSH_APC operand is a column index: evaluate with absolute offset col*height + r */
        uint32_t st[16];
        for (uint32_t e = 0; e <= ints[i].num_args; e++) {
            const uint32_t* bc = ibc + s[e].off;
            int sp = 0;
            for (uint32_t ip = 0; ip < s[e].len;) {
                uint32_t op = bc[ip++];
                switch (op) {
                case OP_PUSH_APC: st[sp++] = mat[(size_t)bc[ip++] * height + r]; break;
                case OP_PUSH_CONST: st[sp++] = bc[ip++] % BB_P; break;
                case OP_ADD: sp--; st[sp - 1] = bb_add(st[sp - 1], st[sp]); break;
                case OP_SUB: sp--; st[sp - 1] = bb_sub(st[sp - 1], st[sp]); break;
                case OP_MUL: sp--; st[sp - 1] = bb_mul(st[sp - 1], st[sp]); break;
                case OP_NEG: st[sp - 1] = bb_neg(st[sp - 1]); break;
                default: st[sp - 1] = bb_inv(st[sp - 1]); break;
                }
            }
            if (e == 0) { mult[i] = st[0]; den[i] = alpha; }
            else den[i] = bb4_add(den[i], bb4_scale(betas[e - 1], st[0]));
        }
        den[i] = bb4_add(den[i], bb4_scale(betas[ints[i].num_args], (ints[i].bus_id + 1) % BB_P));
    }
}

static bb4_t* beta_powers(const uint32_t beta[4], const orc_interaction_t* ints, size_t n_ints) {
    size_t maxk = 0;
    for (size_t i = 0; i < n_ints; i++) if (ints[i].num_args > maxk) maxk = ints[i].num_args;
    bb4_t* b = (bb4_t*)malloc((maxk + 1) * sizeof(bb4_t));
    bb4_t be = {{beta[0], beta[1], beta[2], beta[3]}};
    b[0] = bb4_from_base(1);
    for (size_t j = 1; j <= maxk; j++) b[j] = bb4_mul(b[j - 1], be);
    return b;
}

/* perm: [4*(n_chunks+1)][N] column-major (chunk c limb l at column 4c+l, phi at 4*n_chunks + l); cumsum: phi(N-1) */
void orc_logup_perm_trace(const uint32_t* trace, unsigned log_n, const uint32_t* ibc, const orc_span_t* isp, const orc_interaction_t* ints,
                          size_t n_ints, const uint32_t* chunk_start, size_t n_chunks, const uint32_t alpha_lu[4], const uint32_t beta_lu[4],
                          uint32_t* perm, uint32_t cumsum[4]) {
    const size_t n = (size_t)1 << log_n;
    bb4_t al = {{alpha_lu[0], alpha_lu[1], alpha_lu[2], alpha_lu[3]}};
    bb4_t* betas = beta_powers(beta_lu, ints, n_ints);
    bb4_t* rowsum = (bb4_t*)malloc(n * sizeof(bb4_t));
#pragma omp parallel
    {
        uint32_t* mult = (uint32_t*)malloc((n_ints ? n_ints : 1) * sizeof(uint32_t));
        bb4_t* den = (bb4_t*)malloc((n_ints ? n_ints : 1) * sizeof(bb4_t));
#pragma omp for schedule(static)
        for (long r = 0; r < (long)n; r++) {
            eval_interactions(ibc, isp, ints, n_ints, trace, n, (size_t)r, al, betas, mult, den);
            bb4_t rs = bb4_from_base(0);
            for (size_t c = 0; c < n_chunks; c++) {
                bb4_t v = bb4_from_base(0);
                for (uint32_t i = chunk_start[c]; i < chunk_start[c + 1]; i++) v = bb4_add(v, bb4_scale(bb4_inv(den[i]), mult[i]));
                for (int l = 0; l < 4; l++) perm[(4 * c + l) * n + (size_t)r] = v.c[l];
                rs = bb4_add(rs, v);
            }
            rowsum[r] = rs;
        }
        free(mult); free(den);
    }
    bb4_t phi = bb4_from_base(0);
    for (size_t r = 0; r < n; r++) {
        phi = bb4_add(phi, rowsum[r]);
        for (int l = 0; l < 4; l++) perm[(4 * n_chunks + l) * n + r] = phi.c[l];
    }
    memcpy(cumsum, phi.c, 16);
    free(rowsum); free(betas);
}

/* Horner continuation of the constraint fold over the LDE domain shift*H' (bit-reversed rows, log_blowup 1):
   acc[4][m] (in/out) <- ((acc*alpha + L_0)*alpha + ... )  for the n_chunks + 3 LogUp constraints, per row. */
void orc_logup_fold(const uint32_t* lde, const uint32_t* perm_lde, unsigned log_n, uint32_t shift, const uint32_t* ibc, const orc_span_t* isp,
                    const orc_interaction_t* ints, size_t n_ints, const uint32_t* chunk_start, size_t n_chunks, const uint32_t alpha_lu[4],
                    const uint32_t beta_lu[4], const uint32_t cumsum[4], const uint32_t alpha[4], uint32_t* acc4) {
    const size_t n = (size_t)1 << log_n, m = n << 1;
    const unsigned log_m = log_n + 1;
    bb4_t al = {{alpha_lu[0], alpha_lu[1], alpha_lu[2], alpha_lu[3]}}, a = {{alpha[0], alpha[1], alpha[2], alpha[3]}};
    bb4_t cs = {{cumsum[0], cumsum[1], cumsum[2], cumsum[3]}};
    bb4_t* betas = beta_powers(beta_lu, ints, n_ints);
    const uint32_t w_m = bb_root_of_unity(log_m), w_n_inv = bb_inv(bb_root_of_unity(log_n));
    const uint32_t sn = bb_pow(shift, n);
#pragma omp parallel
    {
        uint32_t* mult = (uint32_t*)malloc((n_ints ? n_ints : 1) * sizeof(uint32_t));
        bb4_t* den = (bb4_t*)malloc((n_ints ? n_ints : 1) * sizeof(bb4_t));
#pragma omp for schedule(static)
        for (long r = 0; r < (long)m; r++) {
            const uint32_t i_nat = bitrev32((uint32_t)r, log_m);
            const size_t r_next = bitrev32((uint32_t)((i_nat + 2) & (m - 1)), log_m);        /* next trace row = 2 LDE points on */
            const uint32_t x = bb_mul(shift, bb_pow(w_m, i_nat));
            const uint32_t zh = bb_sub((i_nat & 1) ? bb_neg(sn) : sn, 1);                      /* x^N - 1 */
            const uint32_t is_first = bb_mul(zh, bb_inv(bb_sub(x, 1)));
            const uint32_t is_last = bb_mul(zh, bb_inv(bb_sub(x, w_n_inv)));
            const uint32_t is_trans = bb_sub(x, w_n_inv);
            eval_interactions(ibc, isp, ints, n_ints, lde, m, (size_t)r, al, betas, mult, den);
            bb4_t acc, sum = bb4_from_base(0), sum_next = bb4_from_base(0);
            for (int l = 0; l < 4; l++) acc.c[l] = acc4[(size_t)l * m + (size_t)r];
            for (size_t c = 0; c < n_chunks; c++) {
                bb4_t pc, pn;
                for (int l = 0; l < 4; l++) { pc.c[l] = perm_lde[(4 * c + l) * m + (size_t)r]; pn.c[l] = perm_lde[(4 * c + l) * m + r_next]; }
                sum = bb4_add(sum, pc);
                sum_next = bb4_add(sum_next, pn);
                bb4_t prod = bb4_from_base(1), num = bb4_from_base(0);
                for (uint32_t i = chunk_start[c]; i < chunk_start[c + 1]; i++) {
                    bb4_t others = bb4_from_base(1);
                    for (uint32_t j = chunk_start[c]; j < chunk_start[c + 1]; j++) if (j != i) others = bb4_mul(others, den[j]);
                    num = bb4_add(num, bb4_scale(others, mult[i]));
                    prod = bb4_mul(prod, den[i]);
                }
                acc = bb4_add(bb4_mul(acc, a), bb4_sub(bb4_mul(pc, prod), num));
            }
            bb4_t phi, phi_next;
            for (int l = 0; l < 4; l++) { phi.c[l] = perm_lde[(4 * n_chunks + l) * m + (size_t)r]; phi_next.c[l] = perm_lde[(4 * n_chunks + l) * m + r_next]; }
            acc = bb4_add(bb4_mul(acc, a), bb4_scale(bb4_sub(phi, sum), is_first));
            acc = bb4_add(bb4_mul(acc, a), bb4_scale(bb4_sub(bb4_sub(phi_next, phi), sum_next), is_trans));
            acc = bb4_add(bb4_mul(acc, a), bb4_scale(bb4_sub(phi, cs), is_last));
            for (int l = 0; l < 4; l++) acc4[(size_t)l * m + (size_t)r] = acc.c[l];
        }
        free(mult); free(den);
    }
    free(betas);
}

/* the same constraints at an out-of-domain point (verifier side): main/perm openings at zeta, perm openings at zeta*w.
   Continues the Horner fold `acc` and returns it. */
bb4_t orc_logup_fold_at_point(bb4_t acc, bb4_t alpha, const uint32_t* main_ys /*[W][4]*/, const uint32_t* perm_ys, const uint32_t* perm_next_ys,
                              unsigned log_n, bb4_t zeta, const uint32_t* ibc, const orc_span_t* isp, const orc_interaction_t* ints, size_t n_ints,
                              const uint32_t* chunk_start, size_t n_chunks, bb4_t al, const uint32_t beta_lu[4], bb4_t cumsum) {
    const size_t n = (size_t)1 << log_n;
    bb4_t* betas = beta_powers(beta_lu, ints, n_ints);
    bb4_t* den = (bb4_t*)malloc((n_ints ? n_ints : 1) * sizeof(bb4_t));
    bb4_t* mult = (bb4_t*)malloc((n_ints ? n_ints : 1) * sizeof(bb4_t));
    for (size_t i = 0; i < n_ints; i++) {
        const orc_span_t* s = isp + ints[i].args_index_off;
        for (uint32_t e = 0; e <= ints[i].num_args; e++) {
            bb4_t st[16];
            int sp = 0;
            const uint32_t* bc = ibc + s[e].off;
            for (uint32_t ip = 0; ip < s[e].len;) {
                uint32_t op = bc[ip++];
                switch (op) {
                case OP_PUSH_APC: memcpy(st[sp++].c, main_ys + 4 * bc[ip++], 16); break;
                case OP_PUSH_CONST: st[sp++] = bb4_from_base(bc[ip++] % BB_P); break;
                case OP_ADD: sp--; st[sp - 1] = bb4_add(st[sp - 1], st[sp]); break;
                case OP_SUB: sp--; st[sp - 1] = bb4_sub(st[sp - 1], st[sp]); break;
                case OP_MUL: sp--; st[sp - 1] = bb4_mul(st[sp - 1], st[sp]); break;
                default: st[sp - 1] = bb4_sub(bb4_from_base(0), st[sp - 1]); break;
                }
            }
            if (e == 0) { mult[i] = st[0]; den[i] = al; }
            else den[i] = bb4_add(den[i], bb4_mul(betas[e - 1], st[0]));
        }
        den[i] = bb4_add(den[i], bb4_scale(betas[ints[i].num_args], (ints[i].bus_id + 1) % BB_P));
    }
    bb4_t zn = bb4_pow(zeta, n), zh = zn;
    zh.c[0] = bb_sub(zh.c[0], 1);
    const uint32_t w_inv = bb_inv(bb_root_of_unity(log_n));
    bb4_t zm1 = zeta, zmw = zeta;
    zm1.c[0] = bb_sub(zm1.c[0], 1);
    zmw.c[0] = bb_sub(zmw.c[0], w_inv);
    const bb4_t is_first = bb4_mul(zh, bb4_inv(zm1)), is_last = bb4_mul(zh, bb4_inv(zmw)), is_trans = zmw;
    bb4_t sum = bb4_from_base(0), sum_next = bb4_from_base(0);
    for (size_t c = 0; c < n_chunks; c++) {
        bb4_t pc, pn;
        /* an Ext4 column is committed as 4 base columns; its opening is sum_l X^l * y_l with y_l in Ext4 */
        pc = bb4_from_base(0); pn = bb4_from_base(0);
        for (int l = 0; l < 4; l++) {
            bb4_t e = bb4_from_base(0), y, yn;
            e.c[l] = 1;
            memcpy(y.c, perm_ys + 4 * (4 * c + l), 16);
            memcpy(yn.c, perm_next_ys + 4 * (4 * c + l), 16);
            pc = bb4_add(pc, bb4_mul(e, y));
            pn = bb4_add(pn, bb4_mul(e, yn));
        }
        sum = bb4_add(sum, pc);
        sum_next = bb4_add(sum_next, pn);
        bb4_t prod = bb4_from_base(1), num = bb4_from_base(0);
        for (uint32_t i = chunk_start[c]; i < chunk_start[c + 1]; i++) {
            bb4_t others = bb4_from_base(1);
            for (uint32_t j = chunk_start[c]; j < chunk_start[c + 1]; j++) if (j != i) others = bb4_mul(others, den[j]);
            num = bb4_add(num, bb4_mul(others, mult[i]));
            prod = bb4_mul(prod, den[i]);
        }
        acc = bb4_add(bb4_mul(acc, alpha), bb4_sub(bb4_mul(pc, prod), num));
    }
    bb4_t phi = bb4_from_base(0), phi_next = bb4_from_base(0);
    for (int l = 0; l < 4; l++) {
        bb4_t e = bb4_from_base(0), y, yn;
        e.c[l] = 1;
        memcpy(y.c, perm_ys + 4 * (4 * n_chunks + l), 16);
        memcpy(yn.c, perm_next_ys + 4 * (4 * n_chunks + l), 16);
        phi = bb4_add(phi, bb4_mul(e, y));
        phi_next = bb4_add(phi_next, bb4_mul(e, yn));
    }
    acc = bb4_add(bb4_mul(acc, alpha), bb4_mul(bb4_sub(phi, sum), is_first));
    acc = bb4_add(bb4_mul(acc, alpha), bb4_mul(bb4_sub(bb4_sub(phi_next, phi), sum_next), is_trans));
    acc = bb4_add(bb4_mul(acc, alpha), bb4_mul(bb4_sub(phi, cumsum), is_last));
    free(betas); free(den); free(mult);
    return acc;
}
