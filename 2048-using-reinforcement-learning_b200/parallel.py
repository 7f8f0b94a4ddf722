"""Multi-GPU plumbing: games are independent, so they shard by global game id with no
collective on the hot path; only the final statistics vector is all-reduced (NCCL over NVLink
on the GPU box, gloo in CPU tests)."""
from __future__ import annotations

from ._lib import STATS_LEN, STATS_MAXSCORE


def shard_range(total: int, rank: int, world: int):
    """Contiguous game-id range [lo, hi) of `rank` (rank r owns g*R/total == r)."""
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def all_reduce_stats(stats, group=None):
    """In-place all-reduce of a g2048_stats_reduce vector: sums everywhere, max for the best score."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    best = stats[STATS_MAXSCORE:STATS_MAXSCORE + 1].clone()
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(best, op=dist.ReduceOp.MAX, group=group)
    stats[STATS_MAXSCORE] = best[0]
    return stats


def describe_stats(stats) -> dict:
    """Readable view of the statistics vector (evaluate_beam_search.py:127-135,201-213)."""
    s = [int(x) for x in stats.tolist()]
    games = max(s[22], 1)
    return {
        "games": s[22],
        "highest_tile_histogram": {str(1 << e if e else 0): s[e] for e in range(18) if s[e]},
        "average_score": s[18] / games, "max_score": s[STATS_MAXSCORE], "average_moves": s[19] / games,
        "valid_moves": s[20], "invalid_moves": s[21],
        "milestone_games": {str(64 << m): s[24 + m] for m in range(8)},
        "nodes": s[32],
    }
