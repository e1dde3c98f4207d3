"""CPU tests of the LogUp / bus argument restatement (oracle/logup.c) and of transcript v2 (proof of work, observed openings,
constant final polynomial): algebra checked against pure-Python Ext4 arithmetic, whole proofs through the independent verifier."""
import numpy as np
import pytest

from util import P, rand_field

W11 = 11


def e4(a):
    return [int(a) % P, 0, 0, 0]


def e4_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def e4_mul(a, b):
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] += a[i] * b[j]
    return [(t[i] + W11 * (t[i + 4] if i + 4 < 7 else 0)) % P for i in range(4)]


def e4_pow(a, e):
    r = e4(1)
    while e:
        if e & 1:
            r = e4_mul(r, a)
        a = e4_mul(a, a)
        e >>= 1
    return r


def e4_inv(a):
    return e4_pow(a, P**4 - 2)


def _bus_machine(width, n_ints, seed, quadratic_every=0, constraints=()):
    from powdr_b200 import machine as M
    bus = M.synthetic_bus(width, n_ints, seed, quadratic_every)
    mach = M.SymbolicMachine(list(constraints), bus)
    return M, mach


def test_chunking_follows_the_degree_rule(orc):
    M, mach = _bus_machine(24, 11, 1, quadratic_every=4)
    _, _, cs = orc.logup_perm_trace(np.zeros((mach.width, 2), dtype=np.uint32), M.compile_bus(mach, 1), [1, 0, 0, 0], [2, 0, 0, 0])
    # interactions 3 and 7 carry a degree-2 argument -> chunks of one; the others pair up in declaration order
    assert cs == [0, 2, 3, 4, 6, 7, 8, 10, 11]


def test_perm_trace_against_python_ext4(orc):
    M, mach = _bus_machine(16, 7, 2, quadratic_every=3)
    rng = np.random.default_rng(5)
    n = 4
    trace = rand_field(rng, (mach.width, n))
    al, be = rand_field(rng, 4).tolist(), rand_field(rng, 4).tolist()
    perm, cum, cs = orc.logup_perm_trace(trace, M.compile_bus(mach, 1), al, be)

    def ev(e, r):
        if isinstance(e, int):
            return e % P
        if isinstance(e, str):
            return int(trace[mach.col_of(e), r])
        if len(e) == 2:
            return (-ev(e[1], r)) % P
        a, b = ev(e[0], r), ev(e[2], r)
        return (a + b) % P if e[1] == "+" else (a - b) % P if e[1] == "-" else (a * b) % P

    phi = e4(0)
    for r in range(n):
        for c in range(len(cs) - 1):
            v = e4(0)
            for i in range(cs[c], cs[c + 1]):
                b = mach.bus_interactions[i]
                d, bp = list(al), e4(1)
                for a in b["args"]:
                    d = e4_add(d, e4_mul(bp, e4(ev(a, r))))
                    bp = e4_mul(bp, be)
                d = e4_add(d, e4_mul(bp, e4(b["id"] + 1)))
                v = e4_add(v, e4_mul(e4_inv(d), e4(ev(b["mult"], r))))
            assert [int(perm[4 * c + l, r]) for l in range(4)] == v, (r, c)
            phi = e4_add(phi, v)
        assert [int(perm[4 * (len(cs) - 1) + l, r]) for l in range(4)] == phi
    assert cum.tolist() == phi


@pytest.mark.parametrize("log_n,width,n_ints,quad", [(3, 12, 5, 0), (5, 20, 9, 3), (6, 10, 1, 0)])
def test_logup_proof_verifies_with_constraint_identity_on_any_trace(orc, log_n, width, n_ints, quad):
    """An AIR with bus interactions only: the LogUp constraints hold for EVERY trace (the permutation columns are built to satisfy
    them), so the quotient is a polynomial, the final FRI polynomial is constant and the verifier's identity at zeta holds."""
    M, mach = _bus_machine(width, n_ints, 10 + log_n, quad)
    bus = M.compile_bus(mach, 1)
    rng = np.random.default_rng(log_n)
    trace = rand_field(rng, (mach.width, 1 << log_n))
    proof, ys, q, _ = orc.prove(trace, [], [], bus, n_queries=6, pow_bits=5)
    assert proof["perm_width"] > 0 and proof["final_poly"][0] == proof["final_poly"][1]
    assert orc.verify_segment([], [], log_n, mach.width, proof, ys, q, check_constraints=True, bus=bus) == 0
    if orc.fast_available():
        p2, ys2, q2, _ = orc.prove(trace, [], [], bus, n_queries=6, pow_bits=5, fast=True)
        assert p2 == proof and (ys2 == ys).all() and (q2 == q).all()


def test_with_constraints_and_interactions_on_a_satisfying_trace(orc):
    from powdr_b200 import machine as M
    base = M.synthetic_machine(18, 5, seed=3)
    mach = M.SymbolicMachine(base.constraints, M.synthetic_bus(base, 6, 4))
    bc, spans = M.compile_constraints(mach)
    bus = M.compile_bus(mach, 1)
    log_n = 4
    trace = np.zeros((mach.width, 1 << log_n), dtype=np.uint32)          # padding rows satisfy a guarded APC
    proof, ys, q, _ = orc.prove(trace, bc, spans, bus, n_queries=4, pow_bits=3)
    assert orc.verify_segment(bc, spans, log_n, mach.width, proof, ys, q, check_constraints=True, bus=bus) == 0
    rng = np.random.default_rng(1)
    bad = rand_field(rng, trace.shape)                                     # random trace: transcript fine, identity fails
    proof, ys, q, _ = orc.prove(bad, bc, spans, bus, n_queries=4, pow_bits=3)
    assert orc.verify_segment(bc, spans, log_n, mach.width, proof, ys, q, check_constraints=False, bus=bus) == 0
    assert orc.verify_segment(bc, spans, log_n, mach.width, proof, ys, q, check_constraints=True, bus=bus) == 16


def test_verifier_rejects_tampering(orc):
    M, mach = _bus_machine(10, 4, 9)
    bus = M.compile_bus(mach, 1)
    rng = np.random.default_rng(2)
    log_n = 4
    trace = rand_field(rng, (mach.width, 1 << log_n))
    proof, ys, q, _ = orc.prove(trace, [], [], bus, n_queries=5, pow_bits=6)
    v = lambda p=proof, y=ys, qq=q: orc.verify_segment([], [], log_n, mach.width, p, y, qq, True, bus=bus)
    assert v() == 0
    p = dict(proof); p["cumulative_sum"] = [(proof["cumulative_sum"][0] + 1) % P] + proof["cumulative_sum"][1:]
    assert v(p) == 2                                     # alpha no longer follows from the transcript
    p = dict(proof); p["pow_witness"] = (proof["pow_witness"] + 1) % P
    assert v(p) in (6, 7)
    y = ys.copy(); y[3, 0] = (int(y[3, 0]) + 1) % P
    assert v(y=y) == 4                                   # every opened value is observed
    p = dict(proof); p["final_poly"] = [proof["final_poly"][0], [(proof["final_poly"][1][0] + 1) % P] + proof["final_poly"][1][1:]]
    assert v(p) == 15                                    # final polynomial must be constant
    qq = q.copy(); qq[0, 1 + mach.width + 8 * (log_n + 1)] ^= 1          # first permutation-row word
    assert v(qq=qq) == 9


def test_non_low_degree_codeword_is_rejected(orc):
    """a proof whose FRI input is not low degree must be rejected although every fold, path and challenge is consistent.
    An honest prover's reduced opening is a polynomial for ANY trace (an unsatisfied constraint is caught by the identity at zeta,
    not by FRI), so the dishonest prover here lies about one opened value: (f(x) - y')/(x - zeta) has a pole, the last FRI layer is
    not constant, and only the constancy check of the final polynomial sees it."""
    from powdr_b200 import machine as M
    mach = M.synthetic_machine(12, 4, seed=8)
    bc, spans = M.compile_constraints(mach)
    rng = np.random.default_rng(3)
    log_n = 5
    trace = rand_field(rng, (mach.width, 1 << log_n))
    proof, ys, q, _ = orc.prove(trace, bc, spans, None, n_queries=4, pow_bits=2)
    assert proof["final_poly"][0] == proof["final_poly"][1]
    assert orc.verify_segment(bc, spans, log_n, mach.width, proof, ys, q) == 0
    assert orc.verify_segment(bc, spans, log_n, mach.width, proof, ys, q, check_constraints=True) == 16
    proof, ys, q, _ = orc.prove(trace, bc, spans, None, n_queries=4, pow_bits=2, cheat_opening=True)
    assert proof["final_poly"][0] != proof["final_poly"][1]
    assert orc.verify_segment(bc, spans, log_n, mach.width, proof, ys, q) == 15


def test_grind_finds_the_smallest_witness(orc):
    ch = orc.Challenger()
    ch.observe(np.arange(5, dtype=np.uint32))
    w = orc.grind(ch, 9)
    import copy
    def ok(c):
        t = orc.Challenger(); C = __import__("ctypes"); C.memmove(C.byref(t.s), C.byref(ch.s), C.sizeof(t.s))
        t.observe(np.array([c], dtype=np.uint32))
        return t.sample() & 511 == 0
    assert ok(w) and not any(ok(c) for c in range(w))
