"""Chain sharding over torch.distributed ranks.

Chains are independent units: rank r owns one contiguous block of global chain
ids and no chain state is ever exchanged.  The Philox counter carries the GLOBAL
chain id, so a chain's stream does not depend on the number of ranks.  The only
collectives on the path are tiny FP64 SUM all-reduces of pooled statistics
(adaptation moments, R-hat/ESS sums); they go through ``torch.distributed``
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import os
from dataclasses import dataclass


@dataclass(frozen=True)
class ChainShard:
    start: int
    stop: int
    total: int

    @property
    def count(self) -> int:
        return self.stop - self.start

    @staticmethod
    def for_rank(total: int, rank: int, world: int) -> "ChainShard":
        """Contiguous, balanced partition: the first ``total % world`` ranks get one extra chain."""
        if world < 1 or not (0 <= rank < world):
            raise ValueError(f"bad rank/world {rank}/{world}")
        if total < world:
            raise ValueError(f"{total} chains cannot be split over {world} ranks")
        base, rem = divmod(total, world)
        start = rank * base + min(rank, rem)
        return ChainShard(start, start + base + (1 if rank < rem else 0), total)

    @staticmethod
    def for_current_rank(total: int) -> "ChainShard":
        rank, world = rank_and_world()
        return ChainShard.for_rank(total, rank, world)


def rank_and_world():
    """(rank, world_size) from torch.distributed when initialised, else from the torchrun env."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def pool_groups(start: int, count: int, group: int) -> int:
    """Number of `group`-aligned blocks of global chain ids that [start, start + count) touches
    (= rsfm_pooled_groups of a sampler holding that range)."""
    return (start + count - 1) // group - start // group + 1


def max_pool_groups(total: int, world: int, group: int) -> int:
    """Largest per-rank group count of the balanced partition: the padded row count of the all-gather."""
    return max(pool_groups(s.start, s.count, group) for s in (ChainShard.for_rank(total, r, world) for r in range(world)))


def all_gather_rows(out, rows):
    """Gather every rank's `rows` [R, W] into `out` [world * R, W] in rank order (= global chain order for the
    contiguous partition).  One small collective on the current stream; a plain copy for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_gather_into_tensor(out, rows)
    else:
        out.copy_(rows)
    return out


def all_reduce_sum_(tensor):
    """In-place SUM all-reduce over the default process group; a no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
