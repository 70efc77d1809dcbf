"""Source-pose sharding of the hot path across the GPUs of one box.

The correspondence search is independent per SOURCE pose (each source point carries its own
cap state across the ascending target loop), so rank r owns a contiguous source-pose range
balanced by point count; scans and trees are replicated.  Residual blocks live on the rank
that found them (pose_index0 is in its range).  The only exchange is one all-reduce of the
packed normal-equation blocks per Gauss-Newton / LM iteration.
"""
import numpy as np


def shard_ranges(offsets, world_size):
    """Contiguous [lo, hi) source-pose ranges, one per rank, balanced by number of points."""
    offsets = np.asarray(offsets, np.int64)
    n = len(offsets) - 1
    total = int(offsets[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        b = int(np.searchsorted(offsets, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_ranges_by_work(work, world_size):
    """Contiguous [lo, hi) source-pose ranges cut at equal measured work (hitl_get_stf_work summed
    over the ranks of the previous search).  Falls back to equal pose counts when no work is known."""
    work = np.asarray(work, np.float64)
    n = len(work)
    if n == 0 or work.sum() <= 0:
        return [(n * r // world_size, n * (r + 1) // world_size) for r in range(world_size)]
    cum = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for r in range(1, world_size):
        b = int(np.searchsorted(cum, cum[-1] * r / world_size, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def concat_stf(parts):
    """Concatenate per-shard CSR results (in rank order) into the full reference-ordered result."""
    pair_i = np.concatenate([p["pair_i"] for p in parts]) if parts else np.zeros(0, np.uint32)
    pair_j = np.concatenate([p["pair_j"] for p in parts]) if parts else np.zeros(0, np.uint32)
    k = np.concatenate([p["k"] for p in parts]) if parts else np.zeros(0, np.uint32)
    idx = np.concatenate([p["idx"] for p in parts]) if parts else np.zeros(0, np.uint32)
    offs, base = [np.zeros(1, np.uint64)], 0
    for p in parts:
        o = np.asarray(p["pair_off"], np.uint64)
        offs.append(o[1:] + np.uint64(base))
        base += int(o[-1])
    return dict(pair_i=pair_i, pair_j=pair_j, pair_off=np.concatenate(offs), k=k, idx=idx,
                n_queries=sum(int(p.get("n_queries", 0)) for p in parts))


def gather_stf(local, group=None):
    """All ranks: exchange the per-shard CSR results (host objects) and return the full
    reference-ordered result.  Not on the per-iteration path — the search result normally
    stays on the GPU that produced it; this is for callers that want the whole list on the host."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    keep = {k: local[k] for k in ("pair_i", "pair_j", "pair_off", "k", "idx")}
    keep["n_queries"] = int(local.get("n_queries", 0))
    dist.all_gather_object(parts, keep, group=group)
    return concat_stf(parts)


def allreduce_normal_equations(packed, group=None):
    """The one collective of a Gauss-Newton / LM iteration: sum the packed per-pose blocks
    [H_diag (N x 9) | g (N x 3) | cost] over the ranks, in place (NCCL on the device buffer
    returned by hitl_normal_eq_device, gloo on host tensors in the CPU tests)."""
    import torch.distributed as dist
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed
