#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/.  Run in the build container (needs /root/reference for the
reference-derived fixtures and the oracle for the generated vectors):  python tests/golden/make_golden.py

  single_div_nondet.machine.json   machine section of /root/reference/autoprecompiles/tests/single_div_nondet.json.gz
  wasm_register_reuse.machine.json machine section of .../wasm_register_reuse.json.gz
  apc_reth_op_bug.machine.json.gz  machine section of .../apc_reth_op_bug.json.gz: a real PRE-optimisation APC (5869 columns,
                                   9168 constraints, 3117 bus interactions) -- the large-AIR case for the compiler, the
                                   interpreter and the chunked JIT
  apc_snapshots.json               the 62 optimized machines of /root/reference/openvm-riscv/tests/apc_snapshots/**
                                   re-serialised in the reference JSON expression schema (constraints + bus interactions)
  fixture_stats.json               sizes/degree histograms of the big reference fixtures (keccak/sha256/ecrecover ...)
  oracle_kat.json                  known-answer vectors produced BY THE ORACLE (Poseidon2, sponge, compress, challenger,
                                   small LDE / fold) -- regression pins for oracle and GPU alike
  segment_2p8_w12.json             a whole-segment proof produced by the oracle
"""
import glob
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import orc  # noqa: E402
from powdr_b200 import machine as M  # noqa: E402
from util import rand_field  # noqa: E402

REF = "/root/reference"


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes")


def main():
    if os.path.isdir(REF):
        for name in ("single_div_nondet", "wasm_register_reuse"):
            doc = json.load(gzip.open(os.path.join(REF, "autoprecompiles/tests/%s.json.gz" % name), "rt"))
            dump(name + ".machine.json", {"machine": doc["machine"], "bus_map": doc.get("bus_map")})
        doc = json.load(gzip.open(os.path.join(REF, "autoprecompiles/tests/apc_reth_op_bug.json.gz"), "rt"))
        with gzip.GzipFile(os.path.join(HERE, "apc_reth_op_bug.machine.json.gz"), "wb", mtime=0) as f:      # 5869 columns, 9168 constraints
            f.write(json.dumps({"machine": doc["machine"], "bus_map": doc.get("bus_map")}, separators=(",", ":")).encode())
        snaps = {}
        for p in sorted(glob.glob(os.path.join(REF, "openvm-riscv/tests/apc_snapshots/*/*.txt"))):
            mach = M.SymbolicMachine.from_snapshot_text(open(p).read())
            key = os.path.relpath(p, os.path.join(REF, "openvm-riscv/tests/apc_snapshots"))
            snaps[key] = {"columns": mach.snapshot_columns, "constraints": mach.constraints, "bus_interactions": mach.bus_interactions}
        dump("apc_snapshots.json", snaps)
        stats = {}
        for name in ("keccak_apc_pre_opt", "ecrecover_apc_pre_opt", "sha256_apc_pre_opt", "single_div_nondet", "apc_reth_op_bug"):
            doc = json.load(gzip.open(os.path.join(REF, "autoprecompiles/tests/%s.json.gz" % name), "rt"))
            mach = M.SymbolicMachine(doc["machine"]["constraints"], doc["machine"]["bus_interactions"], doc["machine"]["derived_columns"])
            hist = [0, 0, 0, 0]
            for c in mach.constraints:
                hist[M.degree(c)] += 1
            stats[name] = {"columns": mach.width, "constraints": len(mach.constraints), "bus_interactions": len(mach.bus_interactions),
                           "degree_hist": hist}
        dump("fixture_stats.json", stats)
    rng = np.random.default_rng(2024)
    kat = {}
    kat["perm_zero"] = orc.poseidon2_permute(np.zeros(16, dtype=np.uint32)).tolist()
    st = np.arange(16, dtype=np.uint32)
    kat["perm_iota"] = orc.poseidon2_permute(st).tolist()
    row = rand_field(rng, 21)
    kat["row21"] = row.tolist()
    kat["hash_row21"] = orc.hash_row(row).tolist()
    kat["compress"] = orc.compress(np.arange(8, dtype=np.uint32), np.arange(8, 16, dtype=np.uint32)).tolist()
    ch = orc.Challenger()
    ch.observe(np.arange(1, 12, dtype=np.uint32))
    kat["challenger_after_11"] = [ch.sample() for _ in range(10)]
    tr = rand_field(rng, (2, 8))
    kat["lde_in"] = tr.tolist()
    kat["lde_out"] = orc.lde_batch(tr, 1, 31).tolist()
    f = rand_field(rng, (8, 4))
    beta = rand_field(rng, 4)
    kat["fold_in"], kat["fold_beta"] = f.tolist(), beta.tolist()
    kat["fold_out"] = orc.fri_fold(f, 31, beta).tolist()
    dump("oracle_kat.json", kat)

    g = {"width": 12, "n_constraints": 5, "seed": 8, "trace_seed": 4242, "log_n": 8}
    mach = M.synthetic_machine(g["width"], g["n_constraints"], seed=g["seed"])
    bc, spans = M.compile_constraints(mach)
    trace = rand_field(np.random.default_rng(g["trace_seed"]), (mach.width, 1 << g["log_n"]))
    g["proof"], _ = orc.prove_segment(trace, bc, spans)
    dump("segment_2p8_w12.json", g)


if __name__ == "__main__":
    main()
