import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import g2048_b200 as G
s = G.BatchedBeamSearch(20, 40, "cuda:0", seed=1234)
out = s.play_games(10000, max_moves=10000)
mv = out["moves"].cpu().numpy(); inv = out["invalid"].cpu().numpy(); val = out["valid"].cpu().numpy()
capped = mv >= 10000
print(json.dumps({"capped_games": int(capped.sum()), "capped_invalid_mean": float(inv[capped].mean()) if capped.any() else 0,
  "capped_valid_mean": float(val[capped].mean()) if capped.any() else 0,
  "moves_hist": np.histogram(mv, bins=[0,500,1000,1500,2000,3000,5000,9999,10001])[0].tolist(),
  "invalid_in_uncapped": int(inv[~capped].sum()), "valid_in_uncapped": int(val[~capped].sum()),
  "max_valid": int(val.max()), "p99_valid": float(np.percentile(val, 99))}))
