/* The drop-in boundary used from plain C: no Python, no torch, nothing but include/powdr_b200.h and libpowdr_b200.so.
 *
 *   gcc -std=c99 -O2 -I include examples/abi_demo.c -L powdr_b200/_lib -lpowdr_b200 -Wl,-rpath,$PWD/powdr_b200/_lib -o abi_demo
 *   ./abi_demo [log_n]
 *
 * Proves one segment of a 3-column AIR with the constraints  a*b - c = 0  and  a*(a - 1)*(a - 2) = 0  over a satisfying
 * trace held in pinned host memory, then asks for 4 query openings.  tests/test_gpu_parity.py builds and runs it and compares
 * the printed commitments with the Python binding's on the same trace.
 * (Reference call site this stands in for: sdk.app_prover(exe)?.prove(stdin), /root/reference/openvm-riscv/src/lib.rs:327-332.) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "powdr_b200.h"

#define P 2013265921u
enum { OP_PUSH_APC = 0, OP_PUSH_CONST = 1, OP_ADD = 2, OP_SUB = 3, OP_MUL = 4 };   /* expr_eval.cuh:12-20 */

static uint32_t to_monty(uint32_t x) { return (uint32_t)((((uint64_t)x) << 32) % P); }

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d\n", #call, rc_); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    const size_t log_n = argc > 1 ? (size_t)atoi(argv[1]) : 10, n = (size_t)1 << log_n, width = 3;
    const uint32_t bytecode[] = {
        OP_PUSH_APC, 0, OP_PUSH_APC, 1, OP_MUL, OP_PUSH_APC, 2, OP_SUB,                                   /* a*b - c          */
        OP_PUSH_APC, 0, OP_PUSH_APC, 0, OP_PUSH_CONST, 1, OP_SUB, OP_MUL, OP_PUSH_APC, 0, OP_PUSH_CONST, 2, OP_SUB, OP_MUL,   /* a(a-1)(a-2) */
    };
    const pb_expr_span_t constraints[2] = {{0, 8}, {8, 14}};

    pb_ctx_t* ctx = NULL;
    CHECK(pb_ctx_create(&ctx, 0, NULL));            /* PB_ERR_NO_DEVICE without a GPU: there is no CPU path */
    CHECK(pb_ctx_set_fri_params(ctx, 4, 16));       /* 4 queries, 16 proof-of-work bits */
    pb_air_t* air = NULL;
    CHECK(pb_air_compile(ctx, bytecode, sizeof bytecode / sizeof bytecode[0], constraints, 2, (uint32_t)width, &air));

    uint32_t* trace = NULL;                          /* column-major [3][n], Montgomery words, pinned */
    CHECK(pb_host_alloc((void**)&trace, 4 * width * n));
    uint64_t s = 88172645463325252ull;
    for (size_t r = 0; r < n; r++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const uint32_t a = (uint32_t)(s % 3), b = (uint32_t)((s >> 8) % P), c = (uint32_t)(((uint64_t)a * b) % P);
        trace[0 * n + r] = to_monty(a);
        trace[1 * n + r] = to_monty(b);
        trace[2 * n + r] = to_monty(c);
    }

    pb_segment_proof_t proof;
    CHECK(pb_prove_segment(ctx, air, trace, log_n, width, 0u /* host trace */, &proof));
    printf("jit %d\n", pb_air_is_jit(air));
    printf("trace_root");
    for (int i = 0; i < 8; i++) printf(" %u", proof.trace_root[i]);
    printf("\nquotient_root");
    for (int i = 0; i < 8; i++) printf(" %u", proof.quotient_root[i]);
    printf("\nfri_layers %u final_len %u final_poly", proof.n_fri_layers, proof.final_len);
    for (uint32_t i = 0; i < proof.final_len; i++)
        for (int l = 0; l < 4; l++) printf(" %u", proof.final_poly[i][l]);
    printf("\n");

    size_t wpq = 0;
    CHECK(pb_query_words(log_n, width, 0, &wpq));
    uint32_t* q = (uint32_t*)malloc(4 * wpq * 4);
    CHECK(pb_query_segment(ctx, q, 4 * wpq));            /* 4 queries: pb_ctx_set_fri_params above */
    printf("query_rows %u %u %u %u\n", q[0], q[wpq], q[2 * wpq], q[3 * wpq]);
    float ms[PB_N_STAGES];
    CHECK(pb_last_stage_ms(ctx, ms));
    printf("total_ms %.3f launches %llu\n", ms[PB_N_STAGES - 1], (unsigned long long)pb_launch_count(ctx));

    free(q);
    CHECK(pb_host_free(trace));
    CHECK(pb_air_free(air));
    CHECK(pb_ctx_destroy(ctx));
    return 0;
}
