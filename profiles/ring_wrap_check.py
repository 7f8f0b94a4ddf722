import sys, json
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import gpu_common as X
from oracle import pyoracle as orc
SEED = 0x0B200B200B200
res = {}
for (n, W, D, cap, g0) in ((300, 5, 7, 4000, 9000), (160, 6, 8, 2500, 4000), (64, 20, 40, 10000, 0)):
    ref = orc.play_games(SEED, g0, n, W, D, max_moves=cap)
    for ring in (2, 8, 32, -1):
        for path in ("team", "warp+tail"):
            knobs = dict(X.PLAY_PATHS[path]); knobs[X.TUNE_PENDING_CAP] = ring
            with X.tuning(knobs):
                out = X.host_play(n, W, D, SEED, game0=g0, max_moves=cap)
            ok = all((out["score"][i], out["moves"][i], out["valid"][i], out["invalid"][i], out["nodes"][i]) ==
                     (ref[i].score, ref[i].moves, ref[i].valid_moves, ref[i].invalid_moves, ref[i].nodes) for i in range(n))
            res[f"n{n}_W{W}_ring{ring}_{path}"] = ok
            print(n, W, D, ring, path, ok, flush=True)
print(json.dumps(res))
