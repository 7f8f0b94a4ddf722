"""Timeline of one g2048_play_games call on the team path (profiling build): when the queue ran dry,
when blocks turned into stall breakers, when the last normal game ended, when the kernel ended.
    G2048_LIB_PATH=.../libg2048_prof.so python profiles/games_timeline.py [games] [W] [D]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import g2048_b200 as G
from g2048_b200 import _lib

lib = _lib.use_device(0)
prof = lib.g2048_debug_games_profile
prof.argtypes = [C.POINTER(C.c_ulonglong)]
out = (C.c_ulonglong * 8)()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
W = int(sys.argv[2]) if len(sys.argv) > 2 else 20
D = int(sys.argv[3]) if len(sys.argv) > 3 else 40
s = G.BatchedBeamSearch(W, D, "cuda:0", seed=1234)
s.play_games(8, max_moves=50, game0=0, stats=False)
prof(out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); o = s.play_games(n, max_moves=10000, game0=0, stats=False); b.record(); torch.cuda.synchronize()
prof(out)
v = list(out)
t0 = v[0]
ms = lambda x: (x - t0) / 1e6          # noqa: E731
moves = o["moves"].cpu(); invalid = o["invalid"].cpu()
print(json.dumps({"games": n, "W": W, "D": D, "event_ms": a.elapsed_time(b),
                  "queue_empty_ms": ms(v[1]), "last_team_game_retired_ms": ms(v[2]),
                  "first_block_idle_ms": ms(v[3]), "last_block_idle_ms": ms(v[4]), "kernel_end_ms": ms(v[5]),
                  "stalled_games": v[6], "spec_rounds": v[7],
                  "longest_valid_chain": int((moves - invalid).max()), "games_at_cap": int((moves == 10000).sum()),
                  "mean_moves": float(moves.float().mean())}))
