import os, sys
sys.path.insert(0, "/root/repo")
import torch
import g2048_b200 as G
from g2048_b200 import _lib
dev = "cuda:0"
env = G.BatchedGame2048Env(65536, dev, seed=1234)
env.reset(); torch.cuda.synchronize(); print("reset ok")
for i in range(4):
    env.rollout(2000); torch.cuda.synchronize(); print("rollout ok", i, env.t)
roots = torch.empty(10000, dtype=torch.int64, device=dev)
_lib.check(_lib.use_device(0).g2048_synthetic_boards(roots.data_ptr(), 10000, 1234, 0, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize(); print("synthetic ok")
for (W, D) in ((20, 40), (15, 20)):
    search = G.BatchedBeamSearch(W, D, dev, seed=1234)
    for call in (7, 7, 8):
        out = search.get_actions(roots, call=call); torch.cuda.synchronize(); print("beam ok", W, D, call, int(out["nodes"].sum()))
