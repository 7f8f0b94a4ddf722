"""One short launch of a hot kernel, for `ncu --set full -k regex:<kernel>` (see profiles/README.md).
usage: python profiles/ncu_cases.py rollout | beam | step | step_plain | lone | games | games10k"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import g2048_b200 as G
from g2048_b200 import _lib

case = sys.argv[1]
dev = "cuda:0"
lib = _lib.use_device(0)
if case == "rollout":                       # env_rollout_kernel: 65,536 envs x 2,000 steps
    env = G.BatchedGame2048Env(65536, dev, seed=1234)
    for _ in range(3):
        env.rollout(2000)
elif case == "beam":                        # beam_search_kernel: 10,000 roots, width 20 depth 40
    n = 10000
    roots = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(lib.g2048_synthetic_boards(roots.data_ptr(), n, 1234, 0, torch.cuda.current_stream().cuda_stream))
    s = G.BatchedBeamSearch(20, 40, dev, seed=1234)
    for c in range(3):
        s.get_actions(roots, call=c)
elif case == "step_plain":                  # the same kernel without programmatic dependent launch (one block of 14 warps per SM)
    _lib.check(lib.g2048_set_tuning(6, 0))
    env = G.BatchedGame2048Env(65536, dev, seed=1234)
    env.rollout(300)
    acts = torch.randint(0, 4, (8, 65536), device=dev, dtype=torch.uint8)
    for i in range(8):
        env.step_fused(acts[i], auto_reset=True, want_obs=True)
elif case == "step":                        # env_step_fused_kernel: 65,536 envs, all outputs incl. observation
    env = G.BatchedGame2048Env(65536, dev, seed=1234)
    env.rollout(300)                        # mid-game boards
    acts = torch.randint(0, 4, (8, 65536), device=dev, dtype=torch.uint8)
    for i in range(8):
        env.step_fused(acts[i], auto_reset=True, want_obs=True)
elif case == "lone":                        # team_games_kernel: one game, one team, 400 moves
    s = G.BatchedBeamSearch(20, 40, dev, seed=1234)
    s.play_games(1, max_moves=400, game0=7, stats=False)
elif case == "games10k":                    # cfg 5 on one GPU: play_games_kernel (one warp per game) then team_games_kernel
    s = G.BatchedBeamSearch(20, 40, dev, seed=1234)
    s.play_games(10000, max_moves=10000, game0=0, stats=False)
elif case == "games":                       # team_games_kernel: the per-GPU share of cfg 5 at 8 GPUs
    s = G.BatchedBeamSearch(20, 40, dev, seed=1234)
    s.play_games(1250, max_moves=10000, game0=0, stats=False)
torch.cuda.synchronize()
print("ok", case)
