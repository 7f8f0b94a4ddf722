"""Which games end a g2048_play_games call, and by which path (profiling build): per-game event times.
    G2048_LIB_PATH=.../libg2048_prof.so python profiles/game_events.py [games] [game0] [W] [D]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import g2048_b200 as G
from g2048_b200 import _lib

lib = _lib.use_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
game0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W = int(sys.argv[3]) if len(sys.argv) > 3 else 20
D = int(sys.argv[4]) if len(sys.argv) > 4 else 40
buf = np.zeros((7, 16384), np.uint64)
times = lib.g2048_debug_game_times
times.argtypes = [C.c_void_p]
s = G.BatchedBeamSearch(W, D, "cuda:0", seed=1234)
s.play_games(8, max_moves=50, game0=0, stats=False)
times(buf.ctypes.data)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); o = s.play_games(n, max_moves=10000, game0=game0, stats=False); b.record(); torch.cuda.synchronize()
times(buf.ctypes.data)
t0 = int(buf[6, 0])
if t0 == 2**64 - 1:                      # no one-warp kernel in this run: first event
    t0 = int(buf[:6][buf[:6] > 0].min())
ms = lambda x: round((int(x) - t0) / 1e6, 2) if x else None      # noqa: E731
moves = o["moves"].cpu().numpy(); invalid = o["invalid"].cpu().numpy(); nodes = o["nodes"].cpu().numpy()
order = np.argsort(buf[0, :n].astype(np.int64))[::-1]
last = []
for g in order[:24]:
    last.append({"game": int(g), "written_ms": ms(buf[0, g]), "parked_ms": ms(buf[1, g]), "to_team_ms": ms(buf[2, g]),
                 "migrated_ms": ms(buf[3, g]), "claimed_ms": ms(buf[4, g]), "split_ms": ms(buf[5, g]),
                 "moves": int(moves[g]), "invalid": int(invalid[g]), "nodes": int(nodes[g])})
w = np.sort((buf[0, :n].astype(np.int64) - t0) / 1e6)
to_team = buf[2, :n]; to_team = to_team[to_team > 0]
print(json.dumps({"games": n, "game0": game0, "event_ms": a.elapsed_time(b),
                  "written_ms_percentiles": {str(p): float(np.percentile(w, p)) for p in (10, 50, 75, 90, 95, 99, 100)},
                  "handed_to_team_kernel": int(len(to_team)),
                  "team_kernel_first_hand_over_ms": ms(to_team.min()) if len(to_team) else None,
                  "team_kernel_last_hand_over_ms": ms(to_team.max()) if len(to_team) else None,
                  "parked_games": int((buf[1, :n] > 0).sum()), "migrated_games": int((buf[3, :n] > 0).sum()),
                  "valid_moves_percentiles": {str(p): float(np.percentile(moves - invalid, p)) for p in (50, 75, 90, 95, 99, 100)},
                  "last_games": last}))
