"""Multi-GPU plumbing of the hot path: utterance sharding and the statistics exchange.

The path shards by utterance with no data-path collective; the only exchange is the packed
(2·D+1)-float [sum of per-utterance means | sum of per-utterance stds | utterance count] buffer of
batch/global normalisation, all-reduced once per batch (the reference issues three all_gathers,
feat_norm.py:146-150,239-266)."""
from typing import List

import torch


def shard_indices(n: int, rank: int, world: int, descending: bool = True) -> List[int]:
    """Utterances of a length-sorted batch that rank `rank` of `world` processes: the reference's DDP
    rule ``batch[start::world]`` with start = world-1-rank for descending order (iterator/abs.py:226-236),
    which also balances the total number of samples per rank."""
    start = world - rank - 1 if descending else rank
    return list(range(start, n, world))


def all_reduce_packed(packed: torch.Tensor) -> torch.Tensor:
    """Sum the packed statistics over all ranks (in place).  No-op without an initialised group."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM)
    return packed


def batch_stats_from_packed(packed: torch.Tensor, D: int):
    """(batch_mean, batch_std) = (sum means, sum stds) / N  (feat_norm.py:250,265)."""
    n = packed[2 * D]
    return packed[:D] / n, packed[D:2 * D] / n
