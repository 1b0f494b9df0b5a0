"""Multi-GPU layout of the generation path: independent sample streams, one process per GPU, no
data-path collective (the reference has no multi-GPU code at all).

A *sample id* is the unit of work.  Rank r of R generates ids ``r, r+R, r+2R, ...``; every random
draw of a sample is a function of ``(base_seed, sample_id)`` only — host scalars through
``sample_seed`` (numpy reseeded per sample), per-voxel noise through the Philox key
``(base_seed, sample_id, stage)`` — so the generated data do not depend on R.  torch.distributed
is used for the timing barrier and the max-over-ranks reduction of the benchmark, nothing else.
"""
from __future__ import annotations

import os


def rank_info() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_ids(first: int, count: int, rank: int, world: int) -> list[int]:
    """Sample ids ``first .. first+count-1`` owned by ``rank``: the ones congruent to rank mod world."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    return [i for i in range(first, first + count) if i % world == rank]


def step_ids(step: int, batch: int, rank: int, world: int) -> list[int]:
    """Ids of one weak-scaling step: every rank generates ``batch`` samples, the job ``world*batch``."""
    base = step * batch * world
    return [base + k * world + rank for k in range(batch)]


def sample_seed(base_seed: int, sample_id: int) -> int:
    """32-bit numpy seed of a sample: splitmix64 of (base_seed, sample_id)."""
    z = (int(base_seed) * 0x9E3779B97F4A7C15 + int(sample_id) + 0x632BE59BD9B4E019) & (2**64 - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    return int((z ^ (z >> 31)) & 0xFFFFFFFF)


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX all-reduce of a scalar (the timing rule of the benchmark); identity for one process."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
