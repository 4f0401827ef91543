"""Rollout sharding across the GPUs of one box.

The path shards naturally across rollouts: lanes never interact in the forward tape and the only
cross-lane operation is the reverse of `batch` (dexopt/src/ops.h:22-34), i.e. the sum that forms the
gradient of the shared policy parameters.  Each rank evaluates its block of rollouts on its own GPU
and one all-reduce(sum) of {loss, gradient} per evaluation completes the step.

Lane-independent residual rows (weight decay, neural.h:159-171) are scalar-typed outputs that every
rank computes; they must count once, so every rank except rank 0 subtracts their contribution.
Plumbing only: torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_lanes(total_lanes: int, rank: int, world: int, lane_width: int = 8):
    """contiguous block of lane tiles for `rank`: (first_lane, n_lanes)"""
    if total_lanes % lane_width:
        raise ValueError("total_lanes must be a multiple of the lane width")
    tiles = total_lanes // lane_width
    base, extra = divmod(tiles, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first * lane_width, count * lane_width


def shard_wide_parameters(ports, wide: np.ndarray, total_lanes: int, first_lane: int, n_lanes: int) -> np.ndarray:
    """slice a wide packed parameter buffer for `total_lanes` rollouts down to one rank's block"""
    out = []
    off = 0
    for p in ports:
        if p.lanes == 1:
            n = p.size // 8
            out.append(wide[off:off + n])
            off += n
        else:
            comps = p.size // (8 * p.lanes)
            block = wide[off:off + comps * total_lanes].reshape(comps, total_lanes)
            out.append(block[:, first_lane:first_lane + n_lanes].reshape(-1))
            off += comps * total_lanes
    return np.concatenate(out) if out else np.zeros(0)


def scalar_row_correction(loss_scalar_rows: float, g_scalar_rows: np.ndarray, rank: int):
    """what a rank subtracts from its local {loss, gradient} so that lane-independent residual rows are
    counted once over the whole job (rank 0 keeps them)"""
    if rank == 0:
        return 0.0, np.zeros_like(g_scalar_rows)
    return loss_scalar_rows, g_scalar_rows


def allreduce_loss_gradient(buffer, group=None):
    """in-place sum of a {loss, gradient} tensor over all ranks (NCCL over NVLink on GPUs)"""
    import torch.distributed as dist
    dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
    return buffer
