"""us per step of the per-step kernel (CUDA graph of 64 steps) against the block size, per batch size.
usage: python profiles/step_geometry.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
from g2048_b200 import _lib
lib = _lib.use_device(0)
res = {}
pdl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.check(lib.g2048_set_tuning(6, pdl))
if os.environ.get('STEP_TABLES'): _lib.check(lib.g2048_set_tuning(3, int(os.environ['STEP_TABLES'])))
sizes = [int(x) for x in os.environ.get('STEP_SIZES', '16384,65536,131072').split(',')]
widths = [int(x) for x in os.environ.get('STEP_WIDTHS', '-1,1,2,4,7,8,14,16,28,32').split(',')]
for n in sizes:
    env = G.BatchedGame2048Env(n, "cuda:0", seed=1)
    env.rollout(300)
    acts = torch.randint(0, 4, (64, n), device="cuda:0", dtype=torch.uint8)
    for bw in widths:
        _lib.check(lib.g2048_set_tuning(5, bw))
        for name, obs in (("obs", True), ("noobs", False)):
            def sixty_four():
                for i in range(64): env.step_fused(acts[i], auto_reset=True, want_obs=obs)
            g = env.graph(sixty_four)
            g.replay(); torch.cuda.synchronize()
            best = 1e9
            for rep in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(8): g.replay()
                e.record(); torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e) * 1e3 / (8 * 64))
            res[f"n{n}_w{bw}_{name}"] = round(best, 3)
_lib.check(lib.g2048_set_tuning(5, -1)); _lib.check(lib.g2048_set_tuning(6, -1))
print(json.dumps(res))
