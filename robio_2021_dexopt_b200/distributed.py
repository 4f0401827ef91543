"""Rollout sharding across the GPUs of one box.

The path shards naturally across rollouts: lanes never interact in the forward tape and the only
cross-lane operation is the reverse of `batch` (dexopt/src/ops.h:22-34), i.e. the sum that forms the
gradient of the shared policy parameters.  Each rank evaluates its block of rollouts on its own GPU
and one all-reduce(sum) of {loss, gradient} per evaluation completes the step.

Lane-independent residual rows (weight decay, neural.h:159-171) are scalar-typed outputs that every
rank computes; they must count once, so every rank except rank 0 subtracts their contribution.
The all-reduce itself is issued by the native library on the objective's stream (include/trx_b200.h,
trx_comm_* / trx_objective_set_comm), so the device-resident solver loops run sharded without the host
in between.  torch.distributed is plumbing only: it carries the 128-byte communicator id from rank 0 to
the others (NCCL or gloo), and stands in for the collective in the CPU (gloo) tests of the host logic.
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


def allreduce_loss_gradient(buffer, group=None):
    """in-place sum of a {loss, gradient} tensor over all ranks through torch.distributed: the host-side
    equivalent of what an objective with a Communicator does on the device (CPU tests, gloo)"""
    import torch.distributed as dist
    dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
    return buffer


class Communicator:
    """the native library's NCCL communicator over the ranks that share a rollout batch (trx_comm_*)"""

    def __init__(self, unique_id: bytes, n_ranks: int, rank: int, device: int):
        import ctypes
        from . import _native
        assert len(unique_id) == 128
        self._h = ctypes.c_void_p()
        self.n_ranks, self.rank, self.device = n_ranks, rank, device
        buf = ctypes.create_string_buffer(unique_id, 128)
        _native.check(_native.lib().trx_comm_create(buf, n_ranks, rank, device, ctypes.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        import ctypes
        from . import _native
        buf = ctypes.create_string_buffer(128)
        _native.check(_native.lib().trx_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int, group=None):
        """rank 0 draws the id, torch.distributed broadcasts it, every rank joins"""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(box[0], world, rank, device)

    def allreduce(self, device_ptr: int, count: int, stream: int = 0) -> None:
        from . import _native
        _native.check(_native.lib().trx_comm_allreduce(self._h, device_ptr, count, stream))

    def close(self) -> None:
        from . import _native
        if self._h:
            _native.lib().trx_comm_destroy(self._h)
            self._h = None
