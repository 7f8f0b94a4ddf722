"""us per step of the per-step kernel (CUDA graph of 64 steps, default launch policy) per batch size.
usage: python profiles/step_sizes.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
res = {}
for n in (4096, 16384, 32768, 49152, 65536, 131072, 262144, 1048576):
    env = G.BatchedGame2048Env(n, "cuda:0", seed=1)
    env.rollout(300)
    acts = torch.randint(0, 4, (64, n), device="cuda:0", dtype=torch.uint8)
    for name, obs in (("obs", True), ("noobs", False)):
        def sixty_four():
            for i in range(64): env.step_fused(acts[i], auto_reset=True, want_obs=obs)
        g = env.graph(sixty_four)
        g.replay(); torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(8): g.replay()
            e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) * 1e3 / (8 * 64))
        res[f"n{n}_{name}"] = {"us_per_step": round(best, 3), "steps_per_s": n / best * 1e6}
print(json.dumps(res))
