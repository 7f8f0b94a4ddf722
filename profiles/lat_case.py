import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G
from g2048_b200 import _lib
n = 148
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=5)
roots = torch.empty(n, dtype=torch.int64, device="cuda")
_lib.check(_lib.use_device(0).g2048_synthetic_boards(roots.data_ptr(), n, 5, 0, torch.cuda.current_stream().cuda_stream))
out = s.new_outputs(n)
for _ in range(4): s.get_actions(roots, call=1, out=out)
torch.cuda.synchronize(); print("ok")
