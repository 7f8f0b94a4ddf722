"""Times the per-step API (g2048_env_step, one launch per env step) as a CUDA graph of 64 steps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
n = int(os.environ.get("SWEEP_ENVS", 65536))
env = G.BatchedGame2048Env(n, "cuda:0", seed=1)
env.reset()
acts = torch.randint(0, 4, (64, n), device="cuda", dtype=torch.uint8)
def body():
    for i in range(64):
        if os.environ.get('SWEEP_AUTORESET'): env.step(acts[i], auto_reset=True)
        else: env.step(acts[i]); env.reset_done()
g = env.graph(body)
for _ in range(3): g.replay()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): g.replay()
e.record(); torch.cuda.synchronize()
us = s.elapsed_time(e) * 1e3 / 640
print(json.dumps({"lib": os.path.basename(G.LIB_PATH), "envs": n, "us_per_step_plus_reset": us, "steps_per_s": n / us * 1e6}))
