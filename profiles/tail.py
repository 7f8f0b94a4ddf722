"""Latency of the whole-game paths: us per move of a lone game, cfg 4 (100 games, 15/20), the per-GPU
share of cfg 5 at 8 GPUs (1,250 games, 20/40) and cfg 5 itself (10,000 games), for each scheduling path
of g2048_play_games (teams of four warps / one warp per game / one warp then teams for the tail).
usage: python profiles/tail.py [quick]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import g2048_b200 as G
from g2048_b200 import _lib

lib = _lib.use_device(0)
quick = len(sys.argv) > 1
res = {}


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best, out


def knobs(direct, tail):
    _lib.check(lib.g2048_set_tuning(1, direct)); _lib.check(lib.g2048_set_tuning(2, tail))


PATHS = {"team": (1 << 30, -1), "warp": (0, 0), "auto": (-1, -1)}
for (W, D) in ((20, 40), (15, 20)):
    s = G.BatchedBeamSearch(W, D, "cuda:0", seed=1234)
    # one get_action on one root, both forms
    roots = torch.empty(1, dtype=torch.int64, device="cuda:0")
    _lib.check(lib.g2048_synthetic_boards(roots.data_ptr(), 1, 5, 3, torch.cuda.current_stream().cuda_stream))
    out = s.new_outputs(1)
    for mode, name in ((1, "warp"), (2, "team")):
        _lib.check(lib.g2048_set_tuning(0, mode))
        ms, _ = timed(lambda: s.get_actions(roots, call=1, out=out), reps=20)
        res[f"get_action_1root_{W}_{D}_{name}_us"] = ms * 1e3
    _lib.check(lib.g2048_set_tuning(0, 0))
    # a lone game: us per move
    for name in ("team", "warp"):
        knobs(*PATHS[name])
        ms, o = timed(lambda: s.play_games(1, max_moves=10000, game0=7, stats=False), reps=2)
        res[f"lone_game_{W}_{D}_{name}_us_per_move"] = ms * 1e3 / int(o["moves"][0])
        res[f"lone_game_{W}_{D}_moves"] = int(o["moves"][0])
    for n in ((100, 1250) if quick else (100, 1250, 10000)):
        for name in PATHS:
            if name == "warp" and n == 10000 and quick:
                continue
            knobs(*PATHS[name])
            ms, o = timed(lambda: s.play_games(n, max_moves=10000, game0=0))
            st = G.describe_stats(o["stats"])
            res[f"games{n}_{W}_{D}_{name}_s"] = ms / 1e3
            res[f"games{n}_{W}_{D}_nodes"] = st["nodes"]
        if n == 1250:                                    # tail threshold sweep on the many-games path
            for thr in (296, 592, 888):
                knobs(0, thr)
                ms, o = timed(lambda: s.play_games(n, max_moves=10000, game0=0))
                res[f"games{n}_{W}_{D}_warp_tail{thr}_s"] = ms / 1e3
    if (W, D) == (20, 40) and not quick:
        for thr in (296, 592, 888, 1776):
            knobs(0, thr)
            ms, o = timed(lambda: s.play_games(10000, max_moves=10000, game0=0))
            res[f"games10000_{W}_{D}_warp_tail{thr}_s"] = ms / 1e3
knobs(-1, -1)
print(json.dumps(res))
