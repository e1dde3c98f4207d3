"""CPU tests: the oracle against (a) everything the reference tree pins for this path -- field facts, expression
semantics, the shipped machines (JSON fixtures + 62 optimized snapshots), the sizes the reference's own tests assert --
(b) algorithm-independent self-checks (naive DFT, restriction, fold identity, Merkle paths) and (c) the committed KATs."""
import json
import os

import numpy as np
import pytest

from util import P, rand_field, bitrev, bitrev_perm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


# ---- field facts pinned in-tree: /root/reference/number/src/baby_bear.rs:46-55 ((p-1)/2 = 0x3c000000) ----
def test_modulus_and_generator():
    assert (P - 1) // 2 == 0x3C000000
    assert P == 15 * 2**27 + 1
    for q in (2, 3, 5):
        assert pow(31, (P - 1) // q, P) != 1          # 31 generates F_p^*
    assert pow(pow(31, 15, P), 2**26, P) == P - 1     # two-adic generator has exact order 2^27


def test_poseidon2_constants_match_published_plonky3_tables():
    """The Grain-LFSR generator reproduces the public Plonky3 BabyBear width-16 tables BABYBEAR_RC16_EXTERNAL_INITIAL,
    BABYBEAR_RC16_INTERNAL and BABYBEAR_RC16_EXTERNAL_FINAL of p3-baby-bear's poseidon2.rs.  The values below are RECOLLECTED
    (the reference tree vendors neither Plonky3 nor a Cargo.lock); all 13 internal constants, the first 8 initial and the first
    8 terminal external words are pinned -- the layout bug of round 1 (only the first word of each 16-block kept for the
    internal rounds) fails this test on internal[1] and on external_terminal[0][0]."""
    doc = json.load(open(os.path.join(os.path.dirname(GOLDEN), "..", "constants", "poseidon2_babybear_w16.json")))
    assert doc["external_initial"][0][:8] == [0x69CBB6AF, 0x46AD93F9, 0x60A00F4E, 0x6B1297CD, 0x23189AFE, 0x732E7BEF, 0x72C246DE, 0x2C941900]
    assert doc["internal"] == [0x5A8053C0, 0x693BE639, 0x3858867D, 0x19334F6B, 0x128F0FD8, 0x4E2B1CCB, 0x61210CE0, 0x3C318939,
                               0x0B5B2F22, 0x2EDB11D5, 0x213EFFDF, 0x0CAC4606, 0x241AF16D]
    assert doc["external_terminal"][0][:8] == [0x7290A80D, 0x6F7E5329, 0x598EC8A8, 0x76A859A0, 0x6559E868, 0x657B83AF, 0x13271D3F, 0x1F876063]
    assert len(doc["external_initial"]) == 4 and len(doc["external_terminal"]) == 4 and len(doc["internal"]) == 13
    d = doc["internal_diag_m1"]
    inv = lambda x: pow(x, P - 2, P)
    assert d == [P - 2, 1, 2, inv(2), 3, 4, P - inv(2), P - 3, P - 4, inv(1 << 8), inv(4), inv(8), inv(1 << 27), P - inv(1 << 8),
                 P - inv(16), P - inv(1 << 27)]


def test_poseidon2_plonky3_default_known_answer(orc):
    """Known-answer test of the whole permutation against Plonky3's own unit test of `default_babybear_poseidon2_16()`
    (p3-baby-bear poseidon2.rs, input 0..15).  The first three output words 1906786279, 1737026427, 1959749225 were written
    down from memory BEFORE the permutation was run here and matched on the first run after the constant-layout fix (93 bits:
    not a coincidence); the remaining 13 words are this repo's output and serve as a regression pin.  Both restatements
    (C oracle, pure-Python) must produce it; the GPU test test_gpu_parity.py::test_poseidon2_plonky3_kat checks the CUDA path."""
    import py_poseidon2 as pp
    exp = [1906786279, 1737026427, 1959749225, 700325316, 1638050605, 1021608788, 1726691001, 1761127344, 1552405120, 417318995,
           36799261, 1215172152, 614923223, 1300746575, 957311597, 304856115]
    assert pp.permute(list(range(16))) == exp
    assert orc.poseidon2_permute(np.arange(16, dtype=np.uint32)).tolist() == exp


# ---- self-checks ----
@pytest.mark.parametrize("log_n", [1, 2, 5, 8])
def test_ntt_vs_naive(orc, log_n):
    rng = np.random.default_rng(log_n)
    a = rand_field(rng, 1 << log_n)
    assert (orc.ntt(a) == orc.dft_naive(a)).all()
    assert (orc.intt(orc.ntt(a)) == a).all()


def test_lde_is_bitreversed_coset_evaluation(orc):
    rng = np.random.default_rng(3)
    log_n = 5
    n = 1 << log_n
    co = rand_field(rng, n)
    lde = orc.lde_batch(orc.ntt(co)[None, :], 1, 31)[0]
    pad = np.zeros(2 * n, dtype=np.uint32)
    pad[:n] = co
    nat = orc.dft_naive(pad, 31)
    assert (lde[bitrev_perm(log_n + 1)] == nat).all()
    # restriction: shift = 1 -> first (bit-reversed) half is the trace
    ev = orc.ntt(co)
    l1 = orc.lde_batch(ev[None, :], 1, 1)[0]
    assert (l1[:n][bitrev_perm(log_n)] == ev).all()


def test_fold_identity(orc):
    """fold(f)(x^2) = f_e(x^2) + beta f_o(x^2): folding a low-degree codeword gives the codeword of the folded polynomial"""
    rng = np.random.default_rng(5)
    log_len = 6
    n = 1 << log_len

    def ext_mul(a, b):
        t = [0] * 7
        for i in range(4):
            for j in range(4):
                t[i + j] = (t[i + j] + int(a[i]) * int(b[j])) % P
        return [(t[i] + 11 * t[i + 4]) % P if i < 3 else t[i] for i in range(4)]

    co = rand_field(rng, (n // 2, 4))
    f = np.zeros((n, 4), dtype=np.uint32)
    for l in range(4):
        pad = np.zeros(n, dtype=np.uint32)
        pad[: n // 2] = co[:, l]
        nat = orc.dft_naive(pad, 31)
        for i in range(n):
            f[bitrev(i, log_len), l] = nat[i]
    beta = rand_field(rng, 4)
    g = orc.fri_fold(f, 31, beta)
    cp = np.zeros((n // 4, 4), dtype=np.uint32)
    for k in range(n // 4):
        m = ext_mul(beta, co[2 * k + 1])
        cp[k] = [(int(co[2 * k][i]) + m[i]) % P for i in range(4)]
    for l in range(4):
        pad = np.zeros(n // 2, dtype=np.uint32)
        pad[: n // 4] = cp[:, l]
        nat = orc.dft_naive(pad, 31 * 31 % P)
        for i in range(n // 2):
            assert g[bitrev(i, log_len - 1), l] == nat[i]


def test_merkle_structure(orc):
    rng = np.random.default_rng(7)
    m0, m1 = rand_field(rng, (11, 8)), rand_field(rng, (3, 8))
    layers = orc.merkle_commit([m0, m1])
    assert [l.shape[0] for l in layers] == [8, 4, 2, 1]
    assert (layers[0][5] == orc.hash_row(np.concatenate([m0[:, 5], m1[:, 5]]))).all()
    assert (layers[1][2] == orc.compress(layers[0][4], layers[0][5])).all()
    assert (layers[3][0] == orc.compress(layers[2][0], layers[2][1])).all()


def test_sponge_is_overwrite_mode(orc):
    """PaddingFreeSponge: a 9-element row = permute(permute(row[0:8] || 0^8) with lane 0 overwritten by row[8])"""
    row = np.arange(100, 109, dtype=np.uint32)
    st = np.zeros(16, dtype=np.uint32)
    st[:8] = row[:8]
    st = orc.poseidon2_permute(st)
    st[0] = row[8]
    st = orc.poseidon2_permute(st)
    assert (orc.hash_row(row) == st[:8]).all()


def test_oracle_kats(orc):
    k = load("oracle_kat.json")
    assert orc.poseidon2_permute(np.zeros(16, dtype=np.uint32)).tolist() == k["perm_zero"]
    assert orc.poseidon2_permute(np.arange(16, dtype=np.uint32)).tolist() == k["perm_iota"]
    assert orc.hash_row(np.array(k["row21"], dtype=np.uint32)).tolist() == k["hash_row21"]
    assert orc.compress(np.arange(8, dtype=np.uint32), np.arange(8, 16, dtype=np.uint32)).tolist() == k["compress"]
    ch = orc.Challenger()
    ch.observe(np.arange(1, 12, dtype=np.uint32))
    assert [ch.sample() for _ in range(10)] == k["challenger_after_11"]
    assert orc.lde_batch(np.array(k["lde_in"], dtype=np.uint32), 1, 31).tolist() == k["lde_out"]
    assert orc.fri_fold(np.array(k["fold_in"], dtype=np.uint32), 31, k["fold_beta"]).tolist() == k["fold_out"]


def test_two_adic_generators_are_plonky3s(orc):
    """The subgroup generators the NTT / LDE / FRI domains are built from, against p3-baby-bear's TWO_ADIC_GENERATORS table
    (recollected public constants: 0x1, 0x78000000, 0x67055c21, 0x5ee99486, 0xbb4c4e4 for 2^0..2^4, 0x67456167 for 2^8,
    0x1a427a41 for 2^27): Plonky3 derives them as 31^((p-1)/2^k) like the restatement does, so the evaluation domains -- not only
    the field -- coincide.  Taken from the oracle itself: the evaluations of the polynomial x over the subgroup."""
    def root(bits):
        e = np.zeros(1 << bits, dtype=np.uint32)
        e[1] = 1
        return int(orc.dft_naive(e, 1)[1])
    known = {1: 0x78000000, 2: 0x67055c21, 3: 0x5ee99486, 4: 0x0bb4c4e4, 8: 0x67456167}
    for bits, w in known.items():
        assert root(bits) == w == pow(31, (P - 1) >> bits, P)
        assert pow(w, 1 << bits, P) == 1 and pow(w, 1 << (bits - 1), P) == P - 1
    assert pow(31, (P - 1) >> 27, P) == 0x1a427a41          # generator of the whole 2-adic subgroup
    assert orc.GENERATOR == 31                               # the coset shift of every LDE is the field generator, as in Plonky3's PCS


def test_poseidon2_is_a_permutation_with_full_diffusion(orc):
    a = orc.poseidon2_permute(np.zeros(16, dtype=np.uint32))
    e = np.zeros(16, dtype=np.uint32)
    e[15] = 1
    b = orc.poseidon2_permute(e)
    assert all(int(x) != int(y) for x, y in zip(a, b))


# ---- expression semantics pinned by the reference (expression/src/lib.rs:179-246) ----
def test_expression_json_codec_example(orc):
    """the reference's own serde test shape: [[5,"*","x"],"-",3]"""
    from powdr_b200 import machine as M
    mach = M.SymbolicMachine([[[5, "*", "x@0"], "-", 3]])
    bc, spans = M.compile_constraints(mach)
    assert bc == [1, 5, 0, 0, 4, 1, 3, 3]            # PUSH_CONST 5, PUSH_APC 0, MUL, PUSH_CONST 3, SUB
    mat = np.array([[7, 0, P - 1]], dtype=np.uint32)
    out = orc.constraint_fold(bc, spans, mat, [1, 0, 0, 0])
    assert out[0].tolist() == [32, P - 3, (5 * (P - 1) - 3) % P] and not out[1:].any()


def test_inv_or_zero_semantics(orc):
    bc = [0, 0, 6]                                       # PUSH_APC 0, INV_OR_ZERO
    mat = np.array([0, 1, 2, P - 1], dtype=np.uint32)
    vals = [orc.eval_expr(bc, mat, r) for r in range(4)]
    assert vals == [0, 1, (P + 1) // 2, P - 1]


def test_snapshot_machines_vanish_on_zero_row(orc):
    """All 62 optimized APC snapshots of the reference (778 constraints, degree <= 3) hold on the all-zero padding row
    (guard invariant, /root/reference/autoprecompiles/src/lib.rs:415-453,470-524; padding at cuda/mod.rs:264-269)."""
    from powdr_b200 import machine as M
    snaps = load("apc_snapshots.json")
    assert len(snaps) == 62
    total = 0
    for key, s in snaps.items():
        mach = M.SymbolicMachine(s["constraints"], s["bus_interactions"])
        total += len(mach.constraints)
        assert max([M.degree(c) for c in mach.constraints] + [0]) <= 3
        for b in mach.bus_interactions:
            assert max(M.degree(e) for e in [b["mult"]] + b["args"]) <= 2
        bc, spans = M.compile_constraints(mach)
        zero = np.zeros((mach.width, 2), dtype=np.uint32)
        assert not orc.constraint_fold(bc, spans, zero, [3, 1, 4, 1]).any(), key
    assert total == 778


def test_reference_fixture_sizes():
    """sizes asserted by the reference's own optimizer tests (/root/reference/autoprecompiles/tests/optimizer.rs:72-83,...)"""
    st = load("fixture_stats.json")
    assert (st["keccak_apc_pre_opt"]["columns"], st["keccak_apc_pre_opt"]["bus_interactions"], st["keccak_apc_pre_opt"]["constraints"]) == (27521, 13262, 28627)
    assert st["keccak_apc_pre_opt"]["degree_hist"] == [1551, 4978, 12029, 10069]
    assert st["single_div_nondet"]["constraints"] == 74 and st["single_div_nondet"]["columns"] == 59


def test_single_div_fixture_against_python_evaluator(orc):
    """independent check of the bytecode path: evaluate the fixture's expression trees directly in Python"""
    from powdr_b200 import machine as M
    mach = M.SymbolicMachine.from_json_file(os.path.join(GOLDEN, "single_div_nondet.machine.json"))
    bc, spans = M.compile_constraints(mach)
    rng = np.random.default_rng(41)
    mat = rand_field(rng, (mach.width, 3))

    def ev(e, r):
        if isinstance(e, int):
            return e % P
        if isinstance(e, str):
            return int(mat[mach.col_of(e), r])
        if len(e) == 2:
            return (-ev(e[1], r)) % P
        a, b = ev(e[0], r), ev(e[2], r)
        return (a + b) % P if e[1] == "+" else (a - b) % P if e[1] == "-" else (a * b) % P

    for k, c in enumerate(mach.constraints):
        o, l = spans[k]
        sub = [(0, l)]
        got = orc.constraint_fold(bc[o:o + l], sub, mat, [1, 0, 0, 0])[0]
        assert got.tolist() == [ev(c, r) for r in range(3)]


def test_quotient_times_vanishing_is_fold(orc):
    from powdr_b200 import machine as M
    mach = M.synthetic_machine(9, 4, seed=2)
    bc, spans = M.compile_constraints(mach)
    rng = np.random.default_rng(43)
    log_n = 4
    n = 1 << log_n
    lde = rand_field(rng, (mach.width, 2 * n))
    alpha = rand_field(rng, 4)
    q = orc.quotient(bc, spans, lde, log_n, alpha)
    fold = orc.constraint_fold(bc, spans, lde, alpha)
    gn = pow(31, n, P)
    for r in range(2 * n):
        i = bitrev(r, log_n + 1)
        z = (gn * (1 if i % 2 == 0 else P - 1) - 1) % P
        chunk, j = r >> log_n, r & (n - 1)
        assert chunk == i % 2
        for l in range(4):
            assert int(q[chunk, l, j]) * z % P == int(fold[l, r])


# ---- stage 0 oracle vs a direct numpy statement of the reference kernels ----
def test_tracegen_oracle(orc):
    rng = np.random.default_rng(47)
    a0, a1 = rand_field(rng, (5, 16)), rand_field(rng, (3, 32))
    subs = [(0, 2, 0, 0), (1, 1, 1, 3), (0, 4, 0, 1)]
    out = orc.apc_tracegen(8, 4, [(a0, 1), (a1, 2)], subs, 6)
    assert (out[0, :6] == a0[2, :6]).all() and not out[0, 6:].any()
    assert (out[3, :6] == a1[1, 1:13:2]).all() and not out[3, 6:].any()
    assert (out[1, :6] == a0[4, :6]).all()
    # derived: col2 = QuotientOrZero(col0, col1) = col0 / col1 (0 when col1 == 0), absolute-offset bytecode
    H = 8
    out[1, 2] = 0
    bc = [0, 1 * H, 6, 0, 0 * H, 4]
    orc.apc_apply_derived(out, 6, [(2, 0, len(bc))], bc)
    for r in range(6):
        d = int(out[1, r])
        assert int(out[2, r]) == (int(out[0, r]) * pow(d, P - 2, P)) % P
    assert not out[2, 6:].any()


def test_eval_at_point_is_polynomial_evaluation(orc):
    rng = np.random.default_rng(73)
    log_n, shift = 6, 31 * 5 % P
    co = rand_field(rng, 1 << log_n)
    ev = orc.dft_naive(co, shift)
    zeta = rand_field(rng, 4)

    def emul(a, b):
        t = [0] * 7
        for i in range(4):
            for j in range(4):
                t[i + j] = (t[i + j] + int(a[i]) * int(b[j])) % P
        return [(t[i] + 11 * t[i + 4]) % P if i < 3 else t[i] for i in range(4)]

    acc = [0, 0, 0, 0]
    for c in co[::-1]:
        acc = emul(acc, zeta)
        acc[0] = (acc[0] + int(c)) % P
    assert orc.eval_at_point(ev[None, :], shift, zeta)[0].tolist() == acc


def test_segment_reduced_opening_folds_to_a_constant(orc):
    """whole-segment invariant of the oracle pipeline: random (unsatisfying) trace, the FRI final polynomial is constant"""
    from powdr_b200 import machine as M
    mach = M.synthetic_machine(10, 4, seed=21)
    bc, spans = M.compile_constraints(mach)
    trace = rand_field(np.random.default_rng(79), (mach.width, 1 << 7))
    proof, _ = orc.prove_segment(trace, bc, spans)
    assert proof["final_len"] == 2 and proof["final_poly"][0] == proof["final_poly"][1]
    assert proof["n_fri_layers"] == 7


# ---- the oracle's verifier (validity is how the reference itself pins its prover: openvm-riscv/src/lib.rs:337-341) ----
def _bool_machine():
    from powdr_b200 import machine as M
    return M.SymbolicMachine([["b@0", "*", ["b@0", "-", 1]], [["b@0", "*", "c@1"], "-", "d@2"]])


def test_verifier_accepts_honest_and_rejects_tampered(orc):
    from powdr_b200 import machine as M
    mach = M.synthetic_machine(10, 4, seed=21)
    bc, spans = M.compile_constraints(mach)
    trace = rand_field(np.random.default_rng(79), (mach.width, 1 << 7))
    proof, ys, q = orc.prove_segment_q(trace, bc, spans, 6)
    p0 = orc.prove_segment(trace, bc, spans, n_queries=6, pow_bits=4)[0]
    assert proof == p0
    assert orc.verify_segment(bc, spans, 7, mach.width, proof, ys, q) == 0
    q2 = q.copy(); q2[2, 5] ^= 1
    assert orc.verify_segment(bc, spans, 7, mach.width, proof, ys, q2) == 8          # trace row no longer matches its path
    q3 = q.copy(); q3[0, -3] ^= 1
    assert orc.verify_segment(bc, spans, 7, mach.width, proof, ys, q3) == 12         # FRI layer path
    q4 = q.copy(); q4[1, 1 + mach.width + 8 * 8 + 8 + 8 * 8 + 2] ^= 1               # a layer-0 pair value
    assert orc.verify_segment(bc, spans, 7, mach.width, proof, ys, q4) in (11, 12, 13)
    ys2 = ys.copy(); ys2[3, 1] = (int(ys2[3, 1]) + 1) % P
    assert orc.verify_segment(bc, spans, 7, mach.width, proof, ys2, q) == 4          # every opened value is bound by the transcript
    bad = dict(proof); bad["final_poly"] = [[1, 2, 3, 4], [1, 2, 3, 4]]
    assert orc.verify_segment(bc, spans, 7, mach.width, bad, ys, q) in (6, 7)        # final polynomial is observed before the PoW


def test_verifier_constraint_identity_on_satisfying_and_unsatisfying_traces(orc):
    from powdr_b200 import machine as M
    mach = _bool_machine()
    bc, spans = M.compile_constraints(mach)
    rng = np.random.default_rng(5)
    b = rng.integers(0, 2, 128).astype(np.uint32)
    c = rand_field(rng, 128)
    d = (b.astype(np.uint64) * c % P).astype(np.uint32)
    trace = np.stack([b, c, d])
    proof, ys, q = orc.prove_segment_q(trace, bc, spans, 5)
    assert orc.verify_segment(bc, spans, 7, 3, proof, ys, q, check_constraints=True) == 0
    trace[2, 17] = (int(trace[2, 17]) + 1) % P                                       # one bad cell
    proof, ys, q = orc.prove_segment_q(trace, bc, spans, 5)
    assert orc.verify_segment(bc, spans, 7, 3, proof, ys, q, check_constraints=False) == 0    # the PCS part is still sound
    assert orc.verify_segment(bc, spans, 7, 3, proof, ys, q, check_constraints=True) == 16


def test_large_preopt_fixture_against_python_evaluator(orc):
    """apc_reth_op_bug: a real pre-optimisation APC (5869 columns, 9168 constraints, degree <= 3).  Sizes as recorded from the
    reference fixture; 300 sampled constraints evaluated directly on the expression trees must equal the bytecode path."""
    from powdr_b200 import machine as M
    mach = M.SymbolicMachine.from_json_file(os.path.join(GOLDEN, "apc_reth_op_bug.machine.json.gz"))
    st = load("fixture_stats.json")["apc_reth_op_bug"]
    assert (mach.width, len(mach.constraints), len(mach.bus_interactions)) == (st["columns"], st["constraints"], st["bus_interactions"]) == (5869, 9168, 3117)
    hist = [0, 0, 0, 0]
    for c in mach.constraints:
        hist[M.degree(c)] += 1
    assert hist == st["degree_hist"]
    bc, spans = M.compile_constraints(mach)
    rng = np.random.default_rng(43)
    mat = rand_field(rng, (mach.width, 2))

    def ev(e, r):
        if isinstance(e, int):
            return e % P
        if isinstance(e, str):
            return int(mat[mach.col_of(e), r])
        if len(e) == 2:
            return (-ev(e[1], r)) % P
        a, b = ev(e[0], r), ev(e[2], r)
        return (a + b) % P if e[1] == "+" else (a - b) % P if e[1] == "-" else (a * b) % P

    for k in rng.choice(len(mach.constraints), size=300, replace=False):
        o, l = spans[k]
        got = orc.constraint_fold(bc[o:o + l], [(0, l)], mat, [1, 0, 0, 0])[0]
        assert got.tolist() == [ev(mach.constraints[k], r) for r in range(2)], k


def test_c_oracle_matches_an_independent_python_poseidon2(orc):
    """two restatements that share only the constants file: oracle/poseidon2.c and tests/py_poseidon2.py (pure Python integers)"""
    import py_poseidon2 as pp
    rng = np.random.default_rng(71)
    for _ in range(5):
        st = rand_field(rng, 16)
        assert orc.poseidon2_permute(st).tolist() == pp.permute(st)
    assert orc.poseidon2_permute(np.zeros(16, dtype=np.uint32)).tolist() == pp.permute([0] * 16)
    for width in (1, 7, 8, 9, 16, 23):
        row = rand_field(rng, width)
        assert orc.hash_row(row).tolist() == pp.hash_row(row), width
    l, r = rand_field(rng, 8), rand_field(rng, 8)
    assert orc.compress(l, r).tolist() == pp.compress(l, r)
    mat = rand_field(rng, (5, 8))                       # 8 rows of 5 columns
    assert orc.merkle_commit([mat])[-1][0].tolist() == pp.merkle_root(mat.T.tolist())


def test_c_oracle_challenger_and_dft_match_independent_python(orc):
    import py_poseidon2 as pp
    rng = np.random.default_rng(73)
    a, b = orc.Challenger(), pp.DuplexChallenger()
    for n_obs, n_samp in [(8, 4), (3, 1), (0, 9), (11, 2), (16, 8), (1, 1)]:
        v = rand_field(rng, n_obs)
        a.observe(v)
        b.observe(v.tolist())
        assert [a.sample() for _ in range(n_samp)] == [b.sample() for _ in range(n_samp)]
    # the O(n^2) DFT the NTT tests lean on, against Python integers; the 8th root of unity is 31^((p-1)/8)
    co = rand_field(rng, 8)
    w = int(orc.dft_naive(np.array([0, 1, 0, 0, 0, 0, 0, 0], dtype=np.uint32), 1)[1])
    assert pow(w, 8, P) == 1 and pow(w, 4, P) == P - 1
    for shift in (1, 31):
        exp = [sum(int(c) * pow(shift * pow(w, i, P) % P, k, P) for k, c in enumerate(co)) % P for i in range(8)]
        assert orc.dft_naive(co, shift).tolist() == exp


def test_size_independent_properties_of_the_oracle(orc):
    """linearity of the LDE and of the FRI fold, and sensitivity of the Merkle root -- the properties the GPU tests rely on at sizes
    the oracle cannot reach quickly"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=12, deadline=None)
    @given(st.integers(1, 7), st.integers(0, 2**32 - 1), st.integers(0, P - 1), st.integers(1, P - 1))
    def lde_is_linear(log_n, seed, k, shift):
        rng = np.random.default_rng(seed)
        a, b = rand_field(rng, (1, 1 << log_n)), rand_field(rng, (1, 1 << log_n))
        comb = ((a.astype(np.uint64) * np.uint64(k) + b) % np.uint64(P)).astype(np.uint32)
        la, lb, lc = (orc.lde_batch(x, 1, shift)[0].astype(np.uint64) for x in (a, b, comb))
        assert (lc == (la * np.uint64(k) + lb) % np.uint64(P)).all()

    @settings(max_examples=12, deadline=None)
    @given(st.integers(2, 8), st.integers(0, 2**32 - 1))
    def fold_is_affine_in_beta(log_len, seed):
        rng = np.random.default_rng(seed)
        f = rand_field(rng, (1 << log_len, 4))
        b0 = orc.fri_fold(f, 31, [0, 0, 0, 0]).astype(np.int64)
        b1 = orc.fri_fold(f, 31, [1, 0, 0, 0]).astype(np.int64)
        b5 = orc.fri_fold(f, 31, [5, 0, 0, 0]).astype(np.int64)
        assert ((b0 + 5 * (b1 - b0)) % P == b5).all()

    @settings(max_examples=8, deadline=None)
    @given(st.integers(1, 5), st.integers(1, 20), st.integers(0, 2**32 - 1))
    def merkle_root_binds_every_element(log_h, width, seed):
        rng = np.random.default_rng(seed)
        m = rand_field(rng, (width, 1 << log_h))
        root = orc.merkle_commit([m])[-1][0].copy()
        c, r = int(rng.integers(width)), int(rng.integers(1 << log_h))
        m[c, r] = (int(m[c, r]) + 1) % P
        assert (orc.merkle_commit([m])[-1][0] != root).any()

    lde_is_linear()
    fold_is_affine_in_beta()
    merkle_root_binds_every_element()


# ---- stage 0 with the REAL substitution tables of the reference fixtures (tests/golden/stage0_subs.json.gz) ----
def _stage0_fixture(name):
    import gzip
    with gzip.open(os.path.join(GOLDEN, "stage0_subs.json.gz"), "rt") as f:
        return json.load(f)[name]


def stage0_tables(name):
    """(n_columns, airs [(key, row_block_size, width)], substs) of a fixture, via the host mirror of cuda/mod.rs:268-326"""
    from powdr_b200 import machine as M
    fx = _stage0_fixture(name)
    airs, substs = M.compile_substitutions(fx["opcodes"], fx["subs"], {i: i for i in range(fx["n_columns"])}, M.rv32_air_of_opcode)
    return fx, airs, substs


def test_rv32_opcode_classes_agree_with_the_fixture_widths():
    """the recollected opcode classes against what the tree pins: all instructions that the class table sends to one AIR use the same
    number of original columns (an AIR has ONE width), different classes differ, and the nondeterministic-division fixture is a DivRem
    instruction"""
    from powdr_b200 import machine as M
    fx = _stage0_fixture("keccak_apc_pre_opt")
    widths = {}
    for op, ss in zip(fx["opcodes"], fx["subs"]):
        widths.setdefault(M.rv32_air_of_opcode(op), set()).add((len(ss), max(o for o, _ in ss) + 1))
    assert widths == {"BaseAlu": {(36, 36)}, "Shift": {(53, 53)}, "LoadStore": {(41, 41)}, "BranchEqual": {(26, 26)}, "JalLui": {(18, 18)}}
    assert M.rv32_air_of_opcode(_stage0_fixture("single_div_nondet")["opcodes"][0]) == "DivRem"


@pytest.mark.parametrize("name,n_instr,n_cols,n_airs", [("single_div_nondet", 1, 59, 1), ("wasm_register_reuse", 2, 64, 2),
                                                         ("keccak_apc_pre_opt", 677, 27521, 5)])
def test_substitution_tables_of_the_reference_fixtures(orc, name, n_instr, n_cols, n_airs):
    """sizes pinned by SURVEY App. A (keccak: 677 instructions, 27 521 columns); every APC column is gathered from exactly one original
    cell; an AIR's row block is as long as the number of its instructions in the block; the CPU mirror of `_apc_tracegen` then
    reproduces a direct numpy statement of the gather on synthetic original traces"""
    fx, airs, substs = stage0_tables(name)
    assert len(fx["opcodes"]) == n_instr and fx["n_columns"] == n_cols and len(airs) == n_airs
    assert sorted(s[3] for s in substs) == list(range(n_cols))
    assert len({(a, c, r) for a, c, r, _ in substs}) == len(substs)
    assert sum(rbs for _, rbs, _ in airs) == sum(1 for ss in fx["subs"] if ss)
    for ai, (_, rbs, width) in enumerate(airs):
        mine = [s for s in substs if s[0] == ai]
        assert max(s[2] for s in mine) == rbs - 1 and max(s[1] for s in mine) == width - 1
    num_calls, H = 5, 8
    rng = np.random.default_rng(n_cols)
    traces = [rand_field(rng, (width, rbs * H)) for _, rbs, width in airs]
    out = orc.apc_tracegen(H, n_cols, [(t, rbs) for t, (_, rbs, _) in zip(traces, airs)], substs, num_calls)
    for ai, col, row, apc_col in substs[:: max(1, len(substs) // 500)]:
        rbs = airs[ai][1]
        assert (out[apc_col, :num_calls] == traces[ai][col, row: rbs * num_calls: rbs]).all()
        assert not out[apc_col, num_calls:].any()
