import json, os, sys
sys.path.insert(0, '/root/repo')
import torch
import g2048_b200 as G
from g2048_b200 import _lib
lib = _lib.use_device(0)
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=1234)
s.play_games(8, max_moves=40)
def timed(n, g0):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); s.play_games(n, max_moves=10000, game0=g0); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3
res = {}
for split in (1, 0):
    _lib.check(lib.g2048_set_tuning(4, split))
    res[f"split{split}"] = {"shard0": timed(1250, 0), "shard4": timed(1250, 5000), "all": timed(10000, 0)}
print(json.dumps(res))
