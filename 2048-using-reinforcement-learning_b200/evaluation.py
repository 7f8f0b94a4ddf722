"""Whole-game evaluation driver: evaluate_beam_search.run_evaluation (evaluate_beam_search.py:100-217)
with the N sequential Python games replaced by one `play_games` launch per GPU.

The returned dict and the `overall_results.json` it writes have the reference's layout
(`scores`, `highest_tiles`, `moves`, `valid_moves`, `invalid_moves`, `milestones`, `best_games`,
`parameters`), so the reference's `create_visualizations(results, save_dir, k, d)` consumes them
unchanged.  With torch.distributed initialised, games are sharded over the ranks by global
game id; the per-game columns are all-gathered as one int32 tensor and the statistics vector is
all-reduced (NCCL over NVLink).
"""
from __future__ import annotations

import datetime
import json
import os
import random as _random

import numpy as np

from .beam import BatchedBeamSearch
from .parallel import all_reduce_stats, describe_stats, shard_range

MILESTONES = (64, 128, 256, 512, 1024, 2048, 4096, 8192)      # evaluate_beam_search.py:42-43


def best_games(scores, k=5):
    """Indices of the top-k games exactly as evaluate_beam_search.py:145-151 maintains them."""
    best = []
    for i, s in enumerate(scores):
        if len(best) < k:
            best.append(i)
            best.sort(key=lambda idx: scores[idx], reverse=True)
        elif s > scores[best[-1]]:
            best[-1] = i
            best.sort(key=lambda idx: scores[idx], reverse=True)
    return best


GAME_KEYS = ("score", "highest_exp", "moves", "valid", "invalid", "milestone")


def gather_games(out, num_games, group=None):
    """Per-rank per-game device tensors (each rank holds the contiguous game-id range of `shard_range`)
    -> host arrays in global game order, on every rank.  The 13 per-game columns travel as ONE int32
    tensor per rank in a tensor all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests): no pickling
    through the host.  Without torch.distributed it is just the device->host copy."""
    import torch
    import torch.distributed as dist
    cols = torch.cat([out["score"].to(torch.int32).unsqueeze(1), out["highest_exp"].to(torch.int32).unsqueeze(1),
                      out["moves"].to(torch.int32).unsqueeze(1), out["valid"].to(torch.int32).unsqueeze(1),
                      out["invalid"].to(torch.int32).unsqueeze(1), out["milestone"].to(torch.int32)], dim=1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        sizes = [shard_range(num_games, r, world) for r in range(world)]
        most = max(hi - lo for lo, hi in sizes)
        padded = torch.zeros(most, cols.shape[1], dtype=torch.int32, device=cols.device)
        padded[:cols.shape[0]] = cols
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        cols = torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    host = cols.cpu().numpy()
    return {"score": host[:, 0], "highest_exp": host[:, 1], "moves": host[:, 2], "valid": host[:, 3],
            "invalid": host[:, 4], "milestone": host[:, 5:13]}


def compile_results(score, highest_exp, moves, valid, invalid, milestone, beam_width, search_depth):
    """Per-game arrays (host) -> the reference's `results` dict (evaluate_beam_search.py:127-135)."""
    score = [int(v) for v in score]
    results = {
        "scores": score,
        "highest_tiles": [int(1 << int(e)) if e else 0 for e in highest_exp],
        "moves": [int(v) for v in moves],
        "valid_moves": [int(v) for v in valid],
        "invalid_moves": [int(v) for v in invalid],
        "milestones": {m: [int(v) for v in milestone[:, j] if v >= 0] for j, m in enumerate(MILESTONES)},
        "best_games": best_games(score),
    }
    results["parameters"] = {"beam_width": beam_width, "search_depth": search_depth, "num_games": len(score)}
    return results


def write_overall_results(results, save_dir):
    """evaluate_beam_search.py:198-214: same keys, same types, indent=4."""
    os.makedirs(save_dir, exist_ok=True)
    doc = {k: results[k] for k in ("scores", "highest_tiles", "moves", "valid_moves", "invalid_moves")}
    doc["milestones"] = {str(k): v for k, v in results["milestones"].items()}
    doc["best_games"] = results["best_games"]
    doc["parameters"] = results["parameters"]
    path = os.path.join(save_dir, "overall_results.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=4)
    return path


def run_evaluation(num_games=1000, beam_width=15, search_depth=20, render_freq=None, save_dir="results",
                   max_moves=10000, seed=None, device=None, timestamp=True):
    """Drop-in for evaluate_beam_search.run_evaluation.  `render_freq` is accepted and ignored
    (no per-move rendering of batched games).  Returns the results dict (the same on every rank)."""
    import torch
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    if device is None:
        # under torch.distributed one process drives one GPU: LOCAL_RANK unless the caller already chose a device
        local = os.environ.get("LOCAL_RANK")
        if distributed and local is not None and torch.cuda.current_device() == 0:
            device = f"cuda:{int(local) % max(torch.cuda.device_count(), 1)}"
        else:
            device = f"cuda:{torch.cuda.current_device()}"
    if seed is None:
        seed = _random.getrandbits(64)
        if distributed:                       # every rank must use rank 0's seed
            box = [seed]
            dist.broadcast_object_list(box, src=0)
            seed = box[0]
    lo, hi = shard_range(num_games, rank, world)
    search = BatchedBeamSearch(beam_width, search_depth, device, seed=seed)
    out = search.play_games(hi - lo, max_moves=max_moves, game0=lo)
    stats = all_reduce_stats(out["stats"])
    host = gather_games(out, num_games)
    results = compile_results(host["score"], host["highest_exp"], host["moves"], host["valid"], host["invalid"],
                              host["milestone"], beam_width, search_depth)
    results["parameters"]["num_games"] = num_games
    results["summary"] = describe_stats(stats)
    results["seed"] = seed
    if rank == 0 and save_dir is not None:
        if timestamp:                         # evaluate_beam_search.py:116-118
            save_dir = f"{save_dir}_{datetime.datetime.now().strftime('%Y%m%d_%H%M%S')}"
        results["save_dir"] = save_dir
        write_overall_results(results, save_dir)
        print(f"\nEvaluation complete. Results saved to {save_dir}")
    return results
