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


# ---------------------------------------------------------------------------------------------------------------------
# One segment on several ranks (pb_prove_segment_sharded, csrc/shard_api.inl), restated with the oracle's arithmetic and REAL
# torch.distributed collectives (gloo): column-sharded trace -> fold per destination block -> all-to-all -> sub-coset
# evaluation of my row block -> subtree root -> all-gather of roots -> top of the tree.  The GPU library is exercised by
# tests/test_gpu_sharded.py; this is the CPU check of the decomposition and of the data flow between ranks.
def _bitrev(x, bits):
    r = 0
    for i in range(bits):
        r = (r << 1) | ((x >> i) & 1)
    return r


def _sharded_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import orc
    from powdr_b200.sharded import shard_columns
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = orc.P
    log_n, width = 6, 5
    n = 1 << log_n
    g = world.bit_length() - 1
    g1, G1 = g - 1, world // 2
    n_blk = n // G1                       # rows per block = 2n / world
    log_blk = log_n - g1
    trace = np.random.default_rng(77).integers(0, P, size=(width, n), dtype=np.uint32)      # same on every rank
    first, count = shard_columns(width, world, rank)
    per = -(-width // world)

    def root_of_unity(bits):              # taken from the oracle itself: evaluations of the polynomial x over the subgroup
        e = np.zeros(1 << bits, dtype=np.uint32)
        e[1] = 1
        return int(orc.dft_naive(e, 1)[1]) if bits else 1

    w_n, w_2n = root_of_unity(log_n), root_of_unity(log_n + 1)

    def block_shift(blk):
        c, h = blk >> g1, blk & (G1 - 1)
        return orc.GENERATOR * pow(w_2n, c, P) * pow(w_n, _bitrev(h, g1), P) % P

    # 1. inverse transform of MY columns, fold for every destination block
    send = np.zeros((world, per, n_blk), dtype=np.int64)
    for i in range(count):
        a = orc.intt(trace[first + i]).astype(object)          # natural-order coefficients
        for blk in range(world):
            lam = pow(block_shift(blk), n_blk, P)
            b = [sum(int(a[r + q * n_blk]) * pow(lam, q, P) for q in range(G1)) % P for r in range(n_blk)]
            send[blk, i] = b
    # 2. all-to-all (gloo has no all_to_all: gather every rank's send buffer and keep the slices addressed to me)
    gathered = [torch.zeros((world, per, n_blk), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(send))
    recv = np.stack([gathered[src][rank].numpy() for src in range(world)]).reshape(world * per, n_blk)[:width]
    # 3. evaluate every column on my sub-coset, rows in bit-reversed order
    sp = block_shift(rank)
    block = np.zeros((width, n_blk), dtype=np.uint32)
    for c in range(width):
        ev = orc.dft_naive(recv[c].astype(np.uint32), sp)
        for m in range(n_blk):
            block[c, _bitrev(m, log_blk)] = ev[m]
    # 4. subtree root, all-gather, top of the tree
    layers = orc.merkle_commit([block])                 # my subtree: layers[k][j] = node j of level k
    my_root = layers[-1][0]
    roots = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(roots, torch.from_numpy(my_root.astype(np.int64)))
    sub_roots = [r.numpy().astype(np.uint32) for r in roots]
    nodes = list(sub_roots)
    while len(nodes) > 1:
        nodes = [orc.compress(nodes[2 * i], nodes[2 * i + 1]) for i in range(len(nodes) // 2)]
    # 5. query phase (pb_query_segment_sharded): the owner of a row block contributes the row and the bottom of its path into a zeroed
    #    share, one all-gather, the shares add up; the top log2(world) siblings come from the gathered subtree roots
    log_ms = log_blk
    log_m = log_n + 1
    idx = [int(x) for x in np.random.default_rng(5).integers(0, 2 * n, size=6)]      # same indices on every rank
    words = width + 8 * log_m
    share = np.zeros((len(idx), words), dtype=np.int64)
    for qi, r in enumerate(idx):
        if r >> log_ms == rank:
            rl = r & (n_blk - 1)
            share[qi, :width] = block[:, rl]
            for k in range(log_ms):
                share[qi, width + 8 * k: width + 8 * k + 8] = layers[k][(rl >> k) ^ 1]
    shares = [torch.zeros((len(idx), words), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(shares, torch.from_numpy(share))
    q = sum(t.numpy() for t in shares)
    for qi, r in enumerate(idx):
        top, lvl, blk = list(sub_roots), 0, r >> log_ms
        while len(top) > 1:
            q[qi, width + 8 * (log_ms + lvl): width + 8 * (log_ms + lvl) + 8] = top[(blk >> lvl) ^ 1]
            top = [orc.compress(top[2 * i], top[2 * i + 1]) for i in range(len(top) // 2)]
            lvl += 1
    ret[rank] = (block, nodes[0], idx, q.astype(np.uint32))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_commit_data_flow_gloo(world):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle import orc
    orc.build()
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    trace = np.random.default_rng(77).integers(0, orc.P, size=(5, 64), dtype=np.uint32)
    lde = orc.lde_batch(trace, 1, orc.GENERATOR)
    ms = lde.shape[1] // world
    root = orc.merkle_commit([lde])[-1][0]
    full = orc.merkle_commit([lde])
    log_m = len(full) - 1
    for r in range(world):
        block, top, idx, q = ret[r]
        assert (block == lde[:, r * ms:(r + 1) * ms]).all(), "rank %d: row block is not the sub-coset evaluation" % r
        assert (top == root).all(), "rank %d: combined root differs from the single-process commitment" % r
        # the assembled openings are the single-process ones: row, then the sibling at every level of the whole tree
        for qi, i in enumerate(idx):
            assert (q[qi, :5] == lde[:, i]).all()
            for k in range(log_m):
                assert (q[qi, 5 + 8 * k: 5 + 8 * k + 8] == full[k][(i >> k) ^ 1]).all(), (r, i, k)
