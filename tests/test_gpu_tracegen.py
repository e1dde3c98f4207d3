"""GPU parity for stage 0: the three drop-in replacements (_apc_tracegen, _apc_apply_derived_expr, _apc_apply_bus) against
the CPU mirror of the reference kernels, on the reference's own machine fixtures where they apply."""
import ctypes as C
import os

import numpy as np
import pytest

from util import P, rand_field

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _structs():
    from powdr_b200 import capi
    return capi


def _upload_struct_array(ctx, arr):
    raw = np.frombuffer(bytes(arr), dtype=np.uint8)
    return ctx.alloc(raw.nbytes).upload(raw)


@pytest.mark.parametrize("H,num_calls", [(8, 6), (1024, 1024), (4096, 3000), (2048, 0)])
def test_apc_tracegen_gather(ctx, orc, H, num_calls):
    capi = _structs()
    rng = np.random.default_rng(H + num_calls)
    # two original AIRs: one consumed row-per-call, one with a row block of 3
    a0 = rand_field(rng, (7, max(H, 8)))
    a1 = rand_field(rng, (5, 3 * max(H, 8)))
    width = 9
    subs = [(0, 2, 0, 0), (1, 1, 1, 3), (0, 6, 0, 1), (1, 4, 2, 8), (1, 0, 0, 5), (0, 0, 0, 7)]
    exp = orc.apc_tracegen(H, width, [(a0, 1), (a1, 3)], subs, num_calls)
    d0, d1 = ctx.to_device(a0, monty=False), ctx.to_device(a1, monty=False)
    airs = (capi.OriginalAir * 2)()
    airs[0].width, airs[0].height, airs[0].buffer, airs[0].row_block_size = 7, a0.shape[1], d0.ptr, 1
    airs[1].width, airs[1].height, airs[1].buffer, airs[1].row_block_size = 5, a1.shape[1], d1.ptr, 3
    S = (capi.Subst * len(subs))()
    for i, s in enumerate(subs):
        S[i].air_index, S[i].col, S[i].row, S[i].apc_col = s
    d_airs, d_subs = _upload_struct_array(ctx, airs), _upload_struct_array(ctx, S)
    init = np.full((width, H), 0xDEADBEEF % P, dtype=np.uint32)     # untouched columns must stay as the caller left them
    d_out = ctx.alloc(init.nbytes).upload(init)
    ctx.apc_tracegen(d_out.ptr, H, d_airs.ptr, d_subs.ptr, len(subs), num_calls)
    got = d_out.download((width, H))
    assert (got == exp).all()


def test_tracegen_rejects_non_power_of_two_height(ctx):
    from powdr_b200.capi import PbError
    with pytest.raises(PbError):
        ctx.apc_tracegen(1, 12, 1, 1, 1, 1)


def test_apc_apply_derived_expr(ctx, orc):
    capi = _structs()
    from powdr_b200 import machine as M
    rng = np.random.default_rng(3)
    H, num_calls = 512, 300
    mach = M.SymbolicMachine(
        [["x@0", "*", "y@1"]], [],
        [["z@2", {"QuotientOrZero": [["x@0", "+", 5], ["y@1", "-", "x@0"]]}], ["k@3", {"Constant": 77}],
         ["q@4", {"QuotientOrZero": ["z@2", "k@3"]}]])        # q depends on the derived z and k of the same row
    specs, bc = M.compile_derived(mach, H)
    base = rand_field(rng, (mach.width, H))
    base[1, :7] = base[0, :7]                                  # zero denominators -> 0
    base[2:] = 123
    exp = orc.apc_apply_derived(base.copy(), num_calls, specs, bc)
    d_out = ctx.to_device(base)                                # Montgomery on device
    D = (capi.DerivedExprSpec * len(specs))()
    for i, (c, o, l) in enumerate(specs):
        D[i].col_base, D[i].span.off, D[i].span.len = c * H, o, l
    d_specs = _upload_struct_array(ctx, D)
    d_bc = ctx.to_device(np.array(bc, dtype=np.uint32), monty=False)
    ctx.apc_apply_derived_expr(d_out.ptr, H, num_calls, d_specs.ptr, len(specs), d_bc.ptr)
    got = ctx.to_host(d_out, base.shape)
    assert (got == exp).all()
    assert not got[2:, num_calls:].any()


def _bus_case(ctx, orc, mach, trace, num_calls, var=(3, 1 << 18), tuple2=(7, 256, 2048), bitwise=6):
    capi = _structs()
    from powdr_b200 import machine as M
    H = trace.shape[1]
    ints, spans, bc = M.compile_bus(mach, H)
    exp = orc.apc_apply_bus(trace, num_calls, bc, ints, spans, var, tuple2, bitwise)
    d_tr = ctx.to_device(trace)
    d_bc = ctx.to_device(np.array(bc, dtype=np.uint32), monty=False)
    I = (capi.DevInteraction * len(ints))()
    for i, (b, n, o) in enumerate(ints):
        I[i].bus_id, I[i].num_args, I[i].args_index_off = b, n, o
    SP = (capi.Span * len(spans))()
    for i, (o, l) in enumerate(spans):
        SP[i].off, SP[i].len = o, l
    d_i, d_s = _upload_struct_array(ctx, I), _upload_struct_array(ctx, SP)
    d_var = ctx.alloc(4 * var[1]).zero()
    d_t2 = ctx.alloc(4 * tuple2[1] * tuple2[2]).zero()
    d_bw = ctx.alloc(4 << 17).zero()
    ctx.apc_apply_bus(d_tr.ptr, num_calls, d_bc.ptr, len(bc), d_i.ptr, len(ints), d_s.ptr, len(spans), var[0], d_var.ptr, var[1],
                      tuple2[0], d_t2.ptr, tuple2[1], tuple2[2], bitwise, d_bw.ptr)
    got = (d_var.download(var[1]), d_t2.download(tuple2[1] * tuple2[2]), d_bw.download(1 << 17))
    for g, e in zip(got, exp):
        assert (g == e).all()
    # the per-AIR generated kernel (pb_bus_compile / pb_bus_apply, column-index bytecode) must fill the same histograms
    handle = ctx.bus_compile(M.compile_bus(mach, 1), mach.width, var[0], tuple2[0], bitwise)
    d_var.zero(); d_t2.zero(); d_bw.zero()
    ctx.bus_apply(handle, d_tr.ptr, H, num_calls, d_var.ptr, var[1], d_t2.ptr, tuple2[1], tuple2[2], d_bw.ptr)
    ctx.synchronize()
    got = (d_var.download(var[1]), d_t2.download(tuple2[1] * tuple2[2]), d_bw.download(1 << 17))
    for g, e in zip(got, exp):
        assert (g == e).all()
    ctx.bus_free(handle)
    return exp


def test_apc_apply_bus_synthetic(ctx, orc):
    from powdr_b200 import machine as M
    rng = np.random.default_rng(5)
    H, num_calls = 1024, 900
    # columns: v (17-bit value), b (byte), c (byte), m (multiplicity 0..3), t (0..255), s (0..2047)
    tr = np.zeros((6, H), dtype=np.uint32)
    tr[0] = rng.integers(0, 1 << 17, H)
    tr[1] = rng.integers(0, 256, H)
    tr[2] = rng.integers(0, 256, H)
    tr[3] = rng.integers(0, 4, H)
    tr[4] = rng.integers(0, 256, H)
    tr[5] = rng.integers(0, 2048, H)
    tr[1, :64] = 7          # many lanes hitting the same bins -> exercises the warp aggregation
    tr[2, :64] = 9
    mach = M.SymbolicMachine([], [
        {"id": 3, "mult": "m@3", "args": ["v@0", 17]},
        {"id": 3, "mult": 1, "args": [["b@1", "+", "c@2"], 9]},
        {"id": 6, "mult": "m@3", "args": ["b@1", "c@2", 0, 0]},
        {"id": 6, "mult": 2, "args": ["b@1", "c@2", 0, 1]},
        {"id": 7, "mult": ["m@3", "*", "m@3"], "args": ["t@4", "s@5"]},
        {"id": 1, "mult": 1, "args": ["v@0", "b@1"]},           # memory bus: ignored by the periphery replay
    ])
    var_h, t2_h, bw_h = _bus_case(ctx, orc, mach, tr, num_calls)
    assert var_h.sum() == int(tr[3, :num_calls].sum()) + num_calls
    assert bw_h[:65536].sum() == int(tr[3, :num_calls].sum()) and bw_h[65536:].sum() == 2 * num_calls


def test_apc_apply_bus_reference_snapshot(ctx, orc):
    """bus interactions of a real optimized APC (complex/rotate.txt: 18 interactions on buses 0,1,3,6) on a synthetic
    trace; periphery interactions whose evaluated arguments fall outside the lookup tables are dropped first"""
    import json
    from powdr_b200 import machine as M
    snaps = json.load(open(os.path.join(GOLDEN, "apc_snapshots.json")))
    s = snaps["complex/rotate.txt"]
    mach = M.SymbolicMachine(s["constraints"], s["bus_interactions"])
    assert mach.width == 26 and len(mach.bus_interactions) == 18
    rng = np.random.default_rng(9)
    H, num_calls = 256, 200
    tr = rng.integers(0, 128, size=(mach.width, H)).astype(np.uint32)
    tr[mach.col_of([n for n in mach.column_names if n.startswith("is_valid")][0])] = 1
    ints, spans, bc = M.compile_bus(mach, H)
    flat = tr.reshape(-1)
    keep = []
    for b, (bus, nargs, off) in zip(mach.bus_interactions, ints):
        ok = True
        if bus in (3, 6):
            for r in range(num_calls):
                vals = [orc.eval_expr(bc[o:o + l], flat, r) for (o, l) in spans[off:off + nargs + 1]]
                if bus == 3 and not (vals[2] <= 17 and vals[1] < (1 << vals[2])):
                    ok = False
                if bus == 6 and not (vals[1] < 256 and vals[2] < 256 and vals[4] <= 1):
                    ok = False
                if not ok:
                    break
        if ok:
            keep.append(b)
    mach2 = M.SymbolicMachine(mach.constraints, keep)
    tr2 = np.ascontiguousarray(tr[[mach.id_to_index[pid] for pid in mach2.column_ids]])
    assert len(keep) >= 8
    _bus_case(ctx, orc, mach2, tr2, num_calls)
