"""Per-step API (one launch per env step): us per step and board-steps/s for the fused step kernel, as
plain launches and as a CUDA graph, at several batch sizes, for both row-move forms.
usage: python profiles/step_sweep.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import g2048_b200 as G
from g2048_b200 import _lib

lib = _lib.use_device(0)
res = {}
for n in (16384, 65536, 1048576):
    env = G.BatchedGame2048Env(n, "cuda:0", seed=1)
    acts = torch.randint(0, 4, (64, n), device="cuda:0", dtype=torch.uint8)
    for tables in (0, 1):
        _lib.check(lib.g2048_set_tuning(3, tables))
        for name, fn in (("step_autoreset", lambda i: env.step(acts[i], auto_reset=True)),
                         ("fused_obs", lambda i: env.step_fused(acts[i], auto_reset=True, want_obs=True)),
                         ("fused_noobs", lambda i: env.step_fused(acts[i], auto_reset=True, want_obs=False))):
            def sixty_four():
                for i in range(64):
                    fn(i)
            g = env.graph(sixty_four)
            g.replay(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(8):
                g.replay()
            e.record(); torch.cuda.synchronize()
            us = s.elapsed_time(e) * 1e3 / (8 * 64)
            key = f"n{n}_{'tables' if tables else 'swar'}_{name}"
            res[key + "_graph_us_per_step"] = us
            res[key + "_graph_steps_per_s"] = n / us * 1e6
            sixty_four(); torch.cuda.synchronize()
            s.record()
            for _ in range(4):
                sixty_four()
            e.record(); torch.cuda.synchronize()
            res[key + "_eager_us_per_step"] = s.elapsed_time(e) * 1e3 / (4 * 64)
_lib.check(lib.g2048_set_tuning(3, -1))
print(json.dumps(res))
