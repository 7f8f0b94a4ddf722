"""World-size-2 `gloo` test of the only collective on the path: the final statistics all-reduce
(sum everywhere, max for the best score), and of the game sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import g2048_b200 as G
from g2048_b200 import _lib


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, total_games, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = G.shard_range(total_games, rank, world)
    # per-rank statistics as g2048_stats_reduce lays them out, from deterministic fake per-game results
    g = np.arange(lo, hi)
    score = (g * 37 % 1000 + 100).astype(np.int64)
    hexp = (g % 5 + 7).astype(np.int64)
    stats = torch.zeros(_lib.STATS_LEN, dtype=torch.int64)
    for e in hexp:
        stats[e] += 1
    stats[18] = int(score.sum()); stats[19] = int((g % 11).sum()); stats[22] = len(g)
    stats[_lib.STATS_MAXSCORE] = int(score.max())
    G.all_reduce_stats(stats)
    if rank == 0:
        out.put(stats.tolist())
    dist.destroy_process_group()


def test_stats_all_reduce_world2():
    total = 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.arange(total)
    score = g * 37 % 1000 + 100
    assert got[22] == total and got[18] == int(score.sum()) and got[_lib.STATS_MAXSCORE] == int(score.max())
    assert got[19] == int((g % 11).sum())
    hist = np.bincount(g % 5 + 7, minlength=18)
    assert got[:18] == hist[:18].tolist()
    d = G.describe_stats(torch.tensor(got))
    assert d["games"] == total and d["max_score"] == int(score.max())


def _gather_main(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from g2048_b200.evaluation import gather_games, compile_results
    lo, hi = G.shard_range(total, rank, world)
    g = np.arange(lo, hi)
    ms = -np.ones((hi - lo, 8), np.int32); ms[:, 0] = g
    t = torch.from_numpy
    host = gather_games({"score": t((g * 3).astype(np.int32)), "highest_exp": t((g % 9 + 3).astype(np.uint8)),
                         "moves": t((g + 100).astype(np.int32)), "valid": t(g.astype(np.int32)),
                         "invalid": t((g % 5).astype(np.int32)), "milestone": t(ms)}, total)
    assert len(host["score"]) == total                       # complete on every rank
    if rank == 0:
        res = compile_results(host["score"], host["highest_exp"], host["moves"], host["valid"], host["invalid"],
                              host["milestone"], 20, 40)
        out.put((res["scores"], res["milestones"][64], res["best_games"]))
    dist.destroy_process_group()


def test_sharded_games_gather_in_global_order_world3():
    total, world = 50, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_main, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    scores, ms64, best = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scores == [3 * g for g in range(total)] and ms64 == list(range(total))
    assert best == [49, 48, 47, 46, 45]


def test_single_process_all_reduce_is_identity():
    s = torch.arange(_lib.STATS_LEN, dtype=torch.int64)
    assert torch.equal(G.all_reduce_stats(s.clone()), s)
