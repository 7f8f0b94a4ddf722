"""us per move of a lone game (20/40 and 15/20) and the 1,250-game shard / 10,000-game times, default policy.
usage: python profiles/lone.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
res = {}
for W, D in ((20, 40), (15, 20)):
    s = G.BatchedBeamSearch(W, D, "cuda:0", seed=1234)
    s.play_games(1, max_moves=50, game0=7, stats=False)
    best = 1e9
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); o = s.play_games(1, max_moves=10000, game0=7, stats=False); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e3 / int(o["moves"][0]))
    res[f"lone_{W}_{D}_us_per_move"] = round(best, 2)
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=1234)
for n, g0 in ((100, 0), (1250, 0), (1250, 5000), (10000, 0)):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); s.play_games(n, max_moves=10000, game0=g0); b.record(); torch.cuda.synchronize()
    res[f"games{n}_from{g0}_s"] = round(a.elapsed_time(b) * 1e-3, 4)
print(json.dumps(res))
