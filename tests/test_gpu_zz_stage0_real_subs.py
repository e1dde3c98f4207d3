"""Stage 0 on the GPU with the REAL gather pattern of the reference's keccak APC (tests/golden/stage0_subs.json.gz: 677 instructions on 10
opcodes, 27 521 substituted cells -- the `block` / `subs` sections of /root/reference/autoprecompiles/tests/keccak_apc_pre_opt.json.gz):
`_apc_tracegen` (same symbol and struct layouts as /root/reference/openvm/src/cuda_abi.rs:8-64) driven by tables built the way
`try_generate_witness` builds them (cuda/mod.rs:268-326, host mirror powdr_b200.machine.compile_substitutions).  The original-AIR
traces are synthetic (the record arenas of the OpenVM chips are not in the reference tree); the output must equal the CPU mirror of the
reference kernel.  (File name sorts last on purpose: added after the round's last GPU call, see profiles/README.md.)"""
import numpy as np
import pytest

from util import P, rand_field
from test_oracle import stage0_tables

pytestmark = pytest.mark.gpu


def _upload_struct_array(ctx, arr):
    raw = np.frombuffer(bytes(arr), dtype=np.uint8)
    return ctx.alloc(raw.nbytes).upload(raw)


@pytest.mark.parametrize("name,H,num_calls", [("wasm_register_reuse", 64, 50), ("keccak_apc_pre_opt", 1024, 1000)])
def test_apc_tracegen_with_the_reference_substitution_tables(ctx, orc, name, H, num_calls):
    from powdr_b200 import capi
    fx, airs, substs = stage0_tables(name)
    width = fx["n_columns"]
    rng = np.random.default_rng(H)
    traces = [rand_field(rng, (w, rbs * H)) for _, rbs, w in airs]
    exp = orc.apc_tracegen(H, width, [(t, rbs) for t, (_, rbs, _) in zip(traces, airs)], substs, num_calls)
    bufs = [ctx.to_device(t, monty=False) for t in traces]
    A = (capi.OriginalAir * len(airs))()
    for i, (t, (_, rbs, w)) in enumerate(zip(traces, airs)):
        A[i].width, A[i].height, A[i].buffer, A[i].row_block_size = w, t.shape[1], bufs[i].ptr, rbs
    S = (capi.Subst * len(substs))()
    for i, s in enumerate(substs):
        S[i].air_index, S[i].col, S[i].row, S[i].apc_col = s
    d_airs, d_subs = _upload_struct_array(ctx, A), _upload_struct_array(ctx, S)
    init = np.full((width, H), 0xDEADBEEF % P, dtype=np.uint32)
    d_out = ctx.alloc(init.nbytes).upload(init)
    ctx.apc_tracegen(d_out.ptr, H, d_airs.ptr, d_subs.ptr, len(substs), num_calls)
    got = d_out.download((width, H))
    assert (got == exp).all()
    # every column of the APC trace is covered by exactly one substitution, so nothing of the initial fill survives
    assert not (got == 0xDEADBEEF % P).all(axis=1).any()
