"""Multi-GPU host logic: one process per GPU (torch.distributed), independent units sharded across ranks, one all-gather
of Merkle caps.  The reference has no distributed code at all (single process, segments proved one after another,
/root/reference/openvm/src/trace_generation.rs:113-140); the units that are independent there -- continuation segments and,
inside a segment, AIR chips -- are what gets partitioned here.  No data-path collective: the only
exchange is the 8-word commitments."""
from typing import List, Sequence, Tuple


def lpt_assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time bin packing of unit indices onto `world` ranks (cost = height * width of a chip).
    Deterministic (ties by index), so every rank computes the same plan without communicating."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    plan: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += costs[i]
    for p in plan:
        p.sort()
    return plan


def my_units(costs: Sequence[float], rank: int, world: int) -> List[int]:
    return lpt_assign(costs, world)[rank]


def all_gather_caps(local_caps, dist=None):
    """local_caps: int32 tensor [k, 8] (k commitments of this rank, padded to the same k on every rank).
    Returns [world, k, 8] on every rank -- the path's single collective (NCCL over NVLink on GPUs, gloo in CPU tests)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_caps.unsqueeze(0)
    world = dist.get_world_size()
    flat = local_caps.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=local_caps.dtype, device=local_caps.device)
    dist.all_gather_into_tensor(out, flat)
    return out.view((world,) + tuple(local_caps.shape))


def plan_summary(costs: Sequence[float], world: int) -> Tuple[float, float]:
    """(max rank load, mean rank load) of the LPT plan -- the imbalance bound for per-chip sharding."""
    plan = lpt_assign(costs, world)
    loads = [sum(costs[i] for i in p) for p in plan]
    return max(loads), sum(loads) / world
