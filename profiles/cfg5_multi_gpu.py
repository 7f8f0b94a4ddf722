"""BASELINE config 5: BeamSearchAgent(20, 40) over 10,000 whole games sharded across the ranks of one
box (one process per GPU, NCCL), final statistics all-reduced.  Launch with torchrun.
usage: torchrun --nproc-per-node N profiles/cfg5_multi_gpu.py [games] [W] [D]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import g2048_b200 as G

games = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 20
D = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
G.BatchedBeamSearch(W, D, f"cuda:{local}", seed=1).play_games(8, max_moves=20)      # warm-up (tables, attributes)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res = G.run_evaluation(games, W, D, save_dir=None, seed=1234, device=f"cuda:{local}")
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local}", dtype=torch.float64)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    s = res["summary"]
    print(json.dumps({"games": games, "W": W, "D": D, "n_gpus": world, "seconds": float(dt.item()),
                      "nodes_per_s": s["nodes"] / float(dt.item()), "summary": s,
                      "scores_checksum": int(sum(res["scores"])), "len_scores": len(res["scores"])}))
if world > 1:
    dist.destroy_process_group()
