"""Times the two hot kernels for the library named by G2048_LIB_PATH (kernel-variant experiments).
usage: G2048_LIB_PATH=... python profiles/sweep.py [tag]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
from g2048_b200 import _lib

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(G.LIB_PATH)
dev = "cuda:0"
N = int(os.environ.get("SWEEP_ENVS", 65536))
env = G.BatchedGame2048Env(N, dev, seed=1234)
env.reset()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(reps):
        flush.fill_(i)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / reps
ms = timed(lambda: env.rollout(2000), 10)
res = {"tag": tag, "envs": N, "rollout_ms": ms, "steps_per_s": N * 2000 / ms * 1e3}
if os.environ.get("SWEEP_ROLLOUT_ONLY"):
    print(json.dumps(res)); sys.exit(0)
for (W, D) in ((20, 40), (15, 20)):
    roots = torch.empty(10000, dtype=torch.int64, device=dev)
    _lib.check(_lib.use_device(0).g2048_synthetic_boards(roots.data_ptr(), 10000, 1234, 0, torch.cuda.current_stream().cuda_stream))
    search = G.BatchedBeamSearch(W, D, dev, seed=1234)
    out = search.get_actions(roots, call=7)
    nodes = int(out["nodes"].sum().item())
    ms = timed(lambda: search.get_actions(roots, call=7), 10)
    res[f"beam_{W}_{D}_ms"] = ms
    res[f"beam_{W}_{D}_nodes_per_s"] = nodes / ms * 1e3
print(json.dumps(res))
