"""world_size-2 gloo test (CPU) of the multi-GPU host logic: deterministic LPT sharding of independent chips and the single
all-gather of Merkle caps.  Commitments are computed with the CPU oracle so the test needs no GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import orc
    from powdr_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 7 "chips" of different shapes; each rank commits the ones the plan gives it
    shapes = [(5, 6), (3, 4), (9, 5), (2, 7), (4, 4), (8, 3), (1, 6)]       # (width, log_height)
    costs = [w * (1 << lh) for w, lh in shapes]
    mine = parallel.my_units(costs, rank, world)
    kmax = max(len(p) for p in parallel.lpt_assign(costs, world))
    caps = np.zeros((kmax, 8), dtype=np.int32)
    for slot, i in enumerate(mine):
        w, lh = shapes[i]
        mat = np.random.default_rng(1000 + i).integers(0, orc.P, size=(w, 1 << lh), dtype=np.uint32)
        caps[slot] = orc.merkle_commit([mat])[-1][0].astype(np.int32)
    allc = parallel.all_gather_caps(torch.from_numpy(caps), dist)
    ret[rank] = (mine, allc.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_plan_is_balanced_and_complete():
    sys.path.insert(0, ROOT)
    from powdr_b200 import parallel
    costs = [7.0, 3.0, 9.0, 1.0, 4.0, 4.0, 8.0, 2.0]
    plan = parallel.lpt_assign(costs, 3)
    assert sorted(i for p in plan for i in p) == list(range(8))
    mx, mean = parallel.plan_summary(costs, 3)
    assert mx <= mean + max(costs)          # LPT guarantee
    assert parallel.lpt_assign(costs, 3) == plan   # deterministic


def test_two_ranks_gather_caps_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle import orc
    from powdr_b200 import parallel
    orc.build()
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (m0, a0), (m1, a1) = ret[0], ret[1]
    assert sorted(m0 + m1) == list(range(7)) and not set(m0) & set(m1)
    assert (a0 == a1).all() and a0.shape[0] == 2
    # rank r's slot s holds the oracle commitment of chip plan[r][s]
    shapes = [(5, 6), (3, 4), (9, 5), (2, 7), (4, 4), (8, 3), (1, 6)]
    for r, mine in enumerate((m0, m1)):
        for slot, i in enumerate(mine):
            w, lh = shapes[i]
            mat = np.random.default_rng(1000 + i).integers(0, orc.P, size=(w, 1 << lh), dtype=np.uint32)
            assert (a0[r, slot].astype(np.uint32) == orc.merkle_commit([mat])[-1][0]).all()
