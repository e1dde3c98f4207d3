/*
 * ORACLE (test infrastructure, not product code): radix-2 NTT and coset low-degree extension.
 * Restates the published two-adic coset LDE used by Plonky3's TwoAdicFriPcs::commit
 * (coset_lde_batch then bit_reverse_rows).  The reference tree has no source for
 * this stage => parity unpinned; self-checked against the O(n^2) DFT below.
 */
#include "oracle.h"
#include "bb31.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_dft_naive(const uint32_t* in, uint32_t* out, unsigned log_n, uint32_t shift) {
    size_t n = (size_t)1 << log_n;
    uint32_t w = bb_root_of_unity(log_n);
    for (size_t k = 0; k < n; k++) {
        uint32_t x = bb_mul(shift, bb_pow(w, k)), acc = 0, xp = 1;
        for (size_t i = 0; i < n; i++) { acc = bb_add(acc, bb_mul(in[i], xp)); xp = bb_mul(xp, x); }
        out[k] = acc;
    }
}

static void bit_reverse_inplace(uint32_t* a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev32((uint32_t)i, log_n); if (i < j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; } }
}

/* textbook iterative Cooley-Tukey: bit-reverse, then spans 1,2,4,... with twiddle w_{2s}^j */
static void ntt_core(uint32_t* a, unsigned log_n, uint32_t root) {
    size_t n = (size_t)1 << log_n;
    bit_reverse_inplace(a, log_n);
    for (unsigned s = 0; s < log_n; s++) {
        size_t half = (size_t)1 << s;
        uint32_t wm = bb_pow(root, n >> (s + 1));
        for (size_t blk = 0; blk < n; blk += 2 * half) {
            uint32_t w = 1;
            for (size_t j = 0; j < half; j++) {
                uint32_t u = a[blk + j], v = bb_mul(a[blk + j + half], w);
                a[blk + j] = bb_add(u, v);
                a[blk + j + half] = bb_sub(u, v);
                w = bb_mul(w, wm);
            }
        }
    }
}

void orc_ntt(uint32_t* a, unsigned log_n) { ntt_core(a, log_n, bb_root_of_unity(log_n)); }

void orc_intt(uint32_t* a, unsigned log_n) {
    size_t n = (size_t)1 << log_n;
    ntt_core(a, log_n, bb_inv(bb_root_of_unity(log_n)));
    uint32_t ninv = bb_inv((uint32_t)(n % BB_P));
    for (size_t i = 0; i < n; i++) a[i] = bb_mul(a[i], ninv);
}

void orc_lde_batch(const uint32_t* trace, unsigned log_n, size_t width, unsigned log_blowup, uint32_t shift, uint32_t* lde) {
    size_t n = (size_t)1 << log_n, m = n << log_blowup;
    unsigned log_m = log_n + log_blowup;
#pragma omp parallel
    {
        uint32_t* buf = (uint32_t*)malloc(m * sizeof(uint32_t));
#pragma omp for schedule(dynamic, 1)
        for (long c = 0; c < (long)width; c++) {
            memcpy(buf, trace + (size_t)c * n, n * sizeof(uint32_t));
            orc_intt(buf, log_n);                        /* coefficients */
            uint32_t sp = 1;
            for (size_t i = 0; i < n; i++) { buf[i] = bb_mul(buf[i], sp); sp = bb_mul(sp, shift); }   /* f(shift*x) */
            memset(buf + n, 0, (m - n) * sizeof(uint32_t));
            orc_ntt(buf, log_m);                          /* natural-order evals on shift*H' */
            uint32_t* out = lde + (size_t)c * m;
            for (size_t i = 0; i < m; i++) out[bitrev32((uint32_t)i, log_m)] = buf[i];   /* bit_reverse_rows */
        }
        free(buf);
    }
}

/* out[j] = (lo+hi)/2 + beta*(lo-hi)/(2 x_j), x_j = shift * w_len^{bitrev(j)}   */
void orc_fri_fold(const uint32_t* in, unsigned log_len, uint32_t shift, const uint32_t beta[4], uint32_t* out) {
    size_t half = (size_t)1 << (log_len - 1);
    uint32_t w = bb_root_of_unity(log_len);
    uint32_t two_inv = bb_inv(2);
    bb4_t b = {{beta[0], beta[1], beta[2], beta[3]}};
#pragma omp parallel for schedule(static) if (half >= 4096)
    for (long j = 0; j < (long)half; j++) {
        bb4_t lo, hi;
        memcpy(lo.c, in + 8 * (size_t)j, 16);
        memcpy(hi.c, in + 8 * (size_t)j + 4, 16);
        uint32_t x = bb_mul(shift, bb_pow(w, bitrev32((uint32_t)j, log_len - 1)));
        uint32_t hx = bb_mul(two_inv, bb_inv(x));
        bb4_t s = bb4_scale(bb4_add(lo, hi), two_inv);
        bb4_t d = bb4_scale(bb4_sub(lo, hi), hx);
        bb4_t r = bb4_add(s, bb4_mul(b, d));
        memcpy(out + 4 * (size_t)j, r.c, 16);
    }
}
