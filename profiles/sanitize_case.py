"""Smallest case that touches every kernel once (for compute-sanitizer; one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import g2048_b200 as G
from g2048_b200 import _lib

env = G.BatchedGame2048Env(600, "cuda:0", seed=3)
env.reset()
a = torch.randint(0, 4, (600,), device="cuda", dtype=torch.uint8)
env.step(a); env.reset_done(); env.rollout(40); env.legal_masks(); env.legal_masks(agent=True); env.observe(); env.values()
env.ppo_features()
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=3)
out = s.get_actions(env.boards)
g = s.play_games(30, max_moves=60)
e = G.Game2048Env(seed=1); e.step(0); e.get_valid_moves(); e.simulate_move(e.get_state(), 1)
G.BeamSearchAgent(5, 6, seed=1).get_action(e.get_state())
torch.cuda.synchronize()
print("sanitize case ok", int(out["nodes"].sum()), G.describe_stats(g["stats"])["games"])
