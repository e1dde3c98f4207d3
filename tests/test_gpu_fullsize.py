"""Sampled parity at the HEADLINE size (BASELINE.json configs[1]: 2^20 rows x 2022 columns): the oracle cannot
prove a segment of this size in test time, but it can check what the GPU produced at sampled positions in seconds --
  * the independent verifier over the whole proof (every challenge, proof of work, per query three Merkle paths to the
    committed roots, the reduced opening against FRI layer 0, all folds, constant final polynomial);
  * the opened main-trace LDE rows against the oracle's own evaluation of the trace polynomials at those domain points;
  * the opened quotient rows against the oracle's constraint fold of the opened main rows:  fold(x) = Z_H(x) * Q(x);
  * for an AIR made of bus interactions only, the LogUp identity at zeta (it holds for any trace), which ties the permutation
    trace, its running sum, the quotient and the two-point openings together at full size with 1734 interactions."""
import ctypes as C

import numpy as np
import pytest

from util import P, bitrev

pytestmark = pytest.mark.gpu

LOG_N, WIDTH = 20, 2022


@pytest.fixture(scope="module")
def big_trace():
    rng = np.random.default_rng(0xB2000001)
    return rng.integers(0, P, size=(WIDTH, 1 << LOG_N), dtype=np.uint32)


def _upload_canonical(ctx, arr):
    buf = ctx.alloc(arr.nbytes).upload(arr)
    assert ctx.lib.pb_to_monty(ctx.h, C.c_void_p(buf.ptr), C.c_size_t(arr.size)) == 0
    ctx.synchronize()
    return buf


def e4_mul(a, b):
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] += a[i] * b[j]
    return [(t[i] + 11 * (t[i + 4] if i + 4 < 7 else 0)) % P for i in range(4)]


def test_keccak_shape_sampled_against_oracle(ctx, orc, big_trace):
    from powdr_b200 import machine as M
    mach = M.synthetic_machine(WIDTH, 187, seed=0xB2000001)
    bc, spans = M.compile_constraints(mach)
    air = ctx.air(bc, spans, mach.width)
    d = _upload_canonical(ctx, big_trace)
    proof = ctx.prove_segment(air, d.ptr, LOG_N, WIDTH, on_device=True)
    queries, ys = ctx.query_segment(LOG_N, WIDTH)
    d.free()
    assert proof["n_fri_layers"] == LOG_N and proof["final_poly"][0] == proof["final_poly"][1]
    assert orc.verify_segment(bc, spans, LOG_N, WIDTH, proof, ys, queries) == 0
    log_m, n = LOG_N + 1, 1 << LOG_N
    w_m = pow(31, (P - 1) >> log_m, P)
    g_n = pow(31, n, P)
    alpha = proof["alpha"]
    n_lde_checks = 4 if orc.fast_available() else 1
    for qi in range(queries.shape[0]):
        r = int(queries[qi, 0])
        i_nat = bitrev(r, log_m)
        x = 31 * pow(w_m, i_nat, P) % P
        row = queries[qi, 1:1 + WIDTH]
        qrow = queries[qi, 1 + WIDTH + 8 * log_m:1 + WIDTH + 8 * log_m + 8]
        if qi < n_lde_checks:          # the opened LDE row IS the trace polynomials evaluated at x (oracle: barycentric / interpolation)
            ev = (orc.fast_eval_at_point if orc.fast_available() else orc.eval_at_point)(big_trace, 1, [x, 0, 0, 0])
            assert (ev[:, 1:] == 0).all() and (ev[:, 0] == row).all(), qi
        # quotient row: alpha-fold of all 187 constraints on the opened row == Z_H(x) * Q_parity(x)
        acc = orc.constraint_fold(bc, spans, row.reshape(WIDTH, 1), alpha)[:, 0].tolist()
        par = i_nat & 1
        zh = (g_n * (P - 1 if par else 1) - 1) % P
        assert [int(v) for v in acc] == [zh * int(qrow[4 * par + l]) % P for l in range(4)], qi


def test_logup_1734_interactions_verify_at_full_size(ctx, orc, big_trace):
    from powdr_b200 import machine as M
    mach = M.SymbolicMachine([], M.synthetic_bus(WIDTH, 1734, seed=0xB2000002))
    w = mach.width                                   # the columns the 1734 interactions reference (1963 of the 2022)
    bus = M.compile_bus(mach, 1)
    air = ctx.air([], [], w, bus)
    assert air.perm_width >= 4 * (800 + 1)          # ~1734 / 2 chunks (interactions with constant-only arguments pack three to a chunk)
    d = _upload_canonical(ctx, big_trace[:w])
    proof = ctx.prove_segment(air, d.ptr, LOG_N, w, on_device=True)
    queries, ys = ctx.query_segment(LOG_N, w, air.perm_width)
    d.free()
    assert proof["final_poly"][0] == proof["final_poly"][1]
    assert orc.verify_segment([], [], LOG_N, w, proof, ys, queries, check_constraints=True, bus=bus) == 0
