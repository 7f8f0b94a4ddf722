"""Hand-over threshold sweep of g2048_play_games (one warp per game -> teams once `thr` games are alive).
usage: python profiles/thr_sweep.py [thr ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
from g2048_b200 import _lib

lib = _lib.use_device(0)
thrs = [int(x) for x in sys.argv[1:]] or [296, 444, 592, 888]
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=1234)
s.play_games(8, max_moves=40)
res = {}
def timed(n, g0):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); s.play_games(n, max_moves=10000, game0=g0); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3
for thr in thrs:
    _lib.check(lib.g2048_set_tuning(1, 0)); _lib.check(lib.g2048_set_tuning(2, thr))
    res[f"thr{thr}"] = {"shard0": timed(1250, 0), "shard1": timed(1250, 1250), "shard4": timed(1250, 5000), "all": timed(10000, 0)}
_lib.check(lib.g2048_set_tuning(1, -1)); _lib.check(lib.g2048_set_tuning(2, -1))
res["auto"] = {"shard0": timed(1250, 0), "shard1": timed(1250, 1250), "shard4": timed(1250, 5000), "all": timed(10000, 0)}
print(json.dumps(res))
