"""Latency of one get_action (1 root) and of small batches, device-side (CUDA events)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
from g2048_b200 import _lib
dev = "cuda:0"
res = {}
for (W, D) in ((20, 40), (15, 20)):
    s = G.BatchedBeamSearch(W, D, dev, seed=5)
    for n in (1, 148, 148 * 4, 148 * 24):
        roots = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.check(_lib.use_device(0).g2048_synthetic_boards(roots.data_ptr(), n, 5, 0, torch.cuda.current_stream().cuda_stream))
        out = s.new_outputs(n)
        for _ in range(3): s.get_actions(roots, call=1, out=out)
        torch.cuda.synchronize()
        ts = []
        for i in range(20):
            torch.cuda._sleep(500000)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.get_actions(roots, call=1, out=out); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        res[f"W{W}_D{D}_n{n}_us_median"] = ts[len(ts) // 2]
        res[f"W{W}_D{D}_n{n}_levels"] = None
print(json.dumps({k: v for k, v in res.items() if v is not None}))
