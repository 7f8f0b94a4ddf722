"""Randomised differential soak: CUDA engine (through the C ABI) vs the CPU oracle over many random
configurations.  Prints one JSON summary line; any mismatch raises.
usage: python profiles/soak.py [rounds] [seed]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import g2048_b200 as G
from oracle import pyoracle as O
from tests import gpu_common as X

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
t0 = time.time()
tot = {"env_board_steps": 0, "beam_roots": 0, "beam_nodes": 0, "games": 0, "game_moves": 0, "configs": 0}
for it in range(rounds):
    seed = int(rng.integers(0, 2**63))
    # ---- env: random n / steps / start step / game offset; boards continue into the beam test
    n = int(rng.choice([1, 7, 64, 333, 1024, 3000])); steps = int(rng.integers(1, 400)); t_start = int(rng.integers(0, 5000))
    game0 = int(rng.integers(0, 2**31))
    ob = np.zeros((n, 16), np.int32); osc = np.zeros(n, np.int64); ohi = np.zeros(n, np.int32); octr = np.zeros(n, np.uint32)
    for i in range(n):
        e = O.Env(seed, game0 + i, ctor_reset=False); e.reset()
        ob[i] = e.board; ohi[i] = e.s.highest_tile; octr[i] = e.s.spawn_ctr
    ors = np.zeros(n); oep = np.zeros(n, np.int32)
    b, s, h, c = X.host_reset(n, seed, game0)
    rs = np.zeros(n); ep = np.zeros(n, np.int32)
    X.host_rollout(b, s, h, c, rs, ep, steps, t_start, seed, game0)
    O.rollout(ob, osc, ohi, octr, ors, oep, steps, t_start, seed, game0)
    assert (b == G.pack_boards(ob)).all() and (s == osc).all() and (c == octr).all() and (rs == ors).all() and (ep == oep).all()
    tot["env_board_steps"] += n * steps
    # ---- the per-step kernel (one launch per step, auto-reset) over the same envs, actions and step range, in a random
    #      launch form (block size, programmatic dependent launch, row move, optional-array variant): it has to arrive
    #      where the fused rollout and the oracle did, float64 reward sum in step order included
    if it % 2 == 0:
        import torch
        ns = min(n, 1024); st = min(steps, 120)
        knobs = {X.TUNE_STEP_TABLES: int(rng.choice([-1, 0, 1])), X.TUNE_STEP_OUTPUTS: int(rng.choice([-1, 0, 1])),
                 X.TUNE_PDL: int(rng.choice([-1, 0, 1])), X.TUNE_STEP_BLOCK_WARPS: int(rng.choice([-1, 1, 3, 7, 14, 28, 32]))}
        acts = np.array([[O.lib().orc_random_action(seed, game0 + i, t_start + t) for i in range(ns)] for t in range(st)], np.uint8)
        env = G.BatchedGame2048Env(ns, "cuda:0", seed=seed, game0=game0)
        env.reset()
        dev_acts = torch.from_numpy(acts).cuda()
        total = torch.zeros(ns, dtype=torch.float64, device="cuda:0")
        fused = rng.random() < 0.5
        with X.tuning(knobs):
            for t in range(st):
                if fused: env.step_fused(dev_acts[t], auto_reset=True, want_obs=bool(t & 1))
                else: env.step(dev_acts[t], auto_reset=True)
                total += env.reward
        sb = np.zeros((ns, 16), np.int32); ssc = np.zeros(ns, np.int64); shi = np.zeros(ns, np.int32); sctr = np.zeros(ns, np.uint32)
        for i in range(ns):
            e = O.Env(seed, game0 + i, ctor_reset=False); e.reset()
            sb[i] = e.board; shi[i] = e.s.highest_tile; sctr[i] = e.s.spawn_ctr
        srs = np.zeros(ns); sep = np.zeros(ns, np.int32)
        O.rollout(sb, ssc, shi, sctr, srs, sep, st, t_start, seed, game0)
        assert (env.boards_u64() == G.pack_boards(sb)).all() and (env.score.cpu().numpy() == ssc).all(), (it, knobs)
        assert (env.spawn_ctr.cpu().numpy().astype(np.uint32) == sctr).all() and (env.episodes.cpu().numpy() == sep).all(), (it, knobs)
        assert (total.cpu().numpy() == srs).all(), (it, knobs)
        tot["per_step_board_steps"] = tot.get("per_step_board_steps", 0) + ns * st
    # ---- beam: random width / depth / thresholds / caller-supplied legality on the boards just reached
    W = int(rng.integers(1, 33)) if rng.random() < 0.85 else int(rng.integers(33, 129))      # 15 %: wide-beam path
    D = int(rng.integers(1, 46))
    early = int(2 ** rng.integers(3, 12)); mid = early * int(2 ** rng.integers(0, 3))
    m = min(n, 400 if W <= 32 else 48)
    vals = ob[:m]; packed = b[:m].copy()
    legal = None
    if rng.random() < 0.5:
        legal = np.array([O.env_legal_mask(v) for v in vals], np.uint8)
    call = rng.integers(0, 10000, m).astype(np.uint32)
    a, p, best, k = X.host_beam(packed, W, D, seed, game0=game0, legal=legal, call=call, early=early, mid=mid)
    for i in range(m):
        o = O.beam_get_action(vals[i], None if legal is None else int(legal[i]), W, D, seed, game0 + i, int(call[i]), early, mid)
        assert (a[i], p[i], k[i], best[i]) == (o.action, o.prob, o.nodes, o.best_score), (it, i, W, D)
    tot["beam_roots"] += m; tot["beam_nodes"] += int(k.sum())
    # ---- a few whole games with a small search (includes the stall-breaker path when games stall)
    if it % 4 == 0:
        g = 24; Wg = int(rng.integers(2, 9)); Dg = int(rng.integers(2, 10)); cap = int(rng.choice([150, 600, 2000]))
        out = X.host_play(g, Wg, Dg, seed, game0=game0, max_moves=cap, early=early, mid=mid)
        for i in range(g):
            r = O.play_game(seed, game0 + i, Wg, Dg, cap, early, mid)
            assert (out["score"][i], out["moves"][i], out["valid"][i], out["invalid"][i], out["nodes"][i]) == \
                (r.score, r.moves, r.valid_moves, r.invalid_moves, r.nodes), (it, i, Wg, Dg)
            assert list(out["milestone"][i]) == list(r.milestone_move)
        tot["games"] += g; tot["game_moves"] += int(out["moves"].sum())
    tot["configs"] += 1
tot["seconds"] = time.time() - t0
tot["overflow_count"] = G.overflow_count()
tot["result"] = "all identical"
print(json.dumps(tot))
