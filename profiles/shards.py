"""Per-rank share of BASELINE cfg 5 in strong scaling, measured on ONE GPU: the R shards of the 10,000 games
(rank r plays games [10000 r / R, 10000 (r+1) / R)) one after the other.  The R-GPU run takes the slowest shard
plus the statistics all-reduce, so this is what bounds `games.cfg5_strong` in bench.py at R GPUs.
usage: python profiles/shards.py [R ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import g2048_b200 as G

ranks = [int(x) for x in sys.argv[1:]] or [2, 4, 8]
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=1234)
s.play_games(8, max_moves=40)
torch.cuda.synchronize()
res = {}
for R in ranks:
    times, longest = [], []
    for r in range(R):
        lo, hi = G.shard_range(10000, r, R)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = s.play_games(hi - lo, max_moves=10000, game0=lo)
        b.record(); torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e-3)
        longest.append(int(out["valid"].max()))
    res[f"R{R}"] = {"shard_seconds": times, "max": max(times), "min": min(times), "longest_productive_game_valid_moves": longest}
print(json.dumps(res))
