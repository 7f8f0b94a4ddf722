"""Where a team search spends its cycles: phases A (expand + ballots), B (spawn + evaluate), C (counting
top-k) of beam_search_team, per level, for a lone game.  Needs the profiling build:
    make -C 2048-using-reinforcement-learning_b200/csrc ../libg2048_prof.so
    G2048_LIB_PATH=2048-using-reinforcement-learning_b200/libg2048_prof.so python profiles/team_phase.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import g2048_b200 as G
from g2048_b200 import _lib

lib = _lib.use_device(0)
prof = lib.g2048_debug_team_profile
prof.argtypes = [C.POINTER(C.c_ulonglong)]
out = (C.c_ulonglong * 8)()
res = {}
_lib.check(lib.g2048_set_tuning(1, 1 << 30))
for (W, D) in ((20, 40), (15, 20)):
    s = G.BatchedBeamSearch(W, D, "cuda:0", seed=1234)
    s.play_games(1, max_moves=200, game0=7, stats=False)
    prof(out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); o = s.play_games(1, max_moves=10000, game0=7, stats=False); b.record(); torch.cuda.synchronize()
    prof(out)
    v = list(out)
    levels, searches = max(v[3], 1), max(v[4], 1)
    res[f"{W}_{D}"] = {"moves": int(o["moves"][0]), "us_per_move": a.elapsed_time(b) * 1e3 / int(o["moves"][0]),
                       "searches_with_levels": v[4], "levels_per_search": levels / searches,
                       "cycles_per_level": {"A_expand": v[0] / levels, "B_spawn_eval": v[1] / levels, "C_topk": v[2] / levels},
                       "cycles_per_search": v[5] / searches}
print(json.dumps(res))
