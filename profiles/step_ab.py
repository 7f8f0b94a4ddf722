import json, os, sys
sys.path.insert(0, '/root/repo')
import torch
import g2048_b200 as G
from g2048_b200 import _lib
lib = _lib.use_device(0)
if os.environ.get('STEP_OUTPUTS'): lib.g2048_set_tuning(8, int(os.environ['STEP_OUTPUTS']))
sizes = [int(x) for x in os.environ.get('STEP_SIZES', '16384,65536,98304,131072').split(',')]
res = {}
for n in sizes:
    env = G.BatchedGame2048Env(n, "cuda:0", seed=1)
    env.rollout(300)
    acts = torch.randint(0, 4, (64, n), device="cuda:0", dtype=torch.uint8)
    for name, fn in (("obs", lambda i: env.step_fused(acts[i], auto_reset=True, want_obs=True)),
                     ("noobs", lambda i: env.step_fused(acts[i], auto_reset=True, want_obs=False))):
        def sixty_four():
            for i in range(64): fn(i)
        g = env.graph(sixty_four)
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(8): g.replay()
            e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) * 1e3 / (8 * 64))
        res[f"n{n}_{name}"] = round(best, 3)
print(json.dumps(res))
