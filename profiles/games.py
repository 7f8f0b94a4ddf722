"""Whole-game runs (BASELINE configs 4 and 5): play_games timing, nodes/s and per-game statistics.
usage: python profiles/games.py <games> <beam_width> <search_depth> [max_moves]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import g2048_b200 as G

games, W, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
search = G.BatchedBeamSearch(W, D, "cuda:0", seed=1234)
search.play_games(min(games, 64), max_moves=50)          # warm-up
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
out = search.play_games(games, max_moves=cap)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e)
st = G.describe_stats(out["stats"])
moves = out["moves"].cpu()
print(json.dumps({"games": games, "W": W, "D": D, "seconds": ms / 1e3, "nodes_per_s": st["nodes"] / ms * 1e3,
                  "moves_per_s": int(moves.sum()) / ms * 1e3, "max_moves_in_a_game": int(moves.max()),
                  "mean_moves": float(moves.float().mean()), "stats": st}))
