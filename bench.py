#!/usr/bin/env python
"""bench.py -- env board-steps/s (headline), beam-search nodes/s and whole-game runs on B200, beside the CPU oracle.

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine, one process per GPU
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Headline workload (BASELINE.json configs[1]): 65,536 boards per GPU x 2,000 random-policy env steps.
One bench "step" = one pass of that workload = one fused rollout launch (131,072,000 board-steps
per GPU).  Boards are independent, so N GPUs each run their own 65,536 boards (weak scaling, no
collective on the data path); `value` = all ranks' board-steps / max-over-ranks device time.

The JSON line also carries
  e2e           same metric through the host-buffer C ABI, H2D + D2H inside the timed region
  roofline      algorithmic HBM bytes of SURVEY 8d (22 B per board-step) over the measured HBM peak
  issue         ALU-pipe / issue-slot view of the same launch; the per-instruction constants come from the
                committed ncu capture (profiles/ncu_constants.json) and are flagged `stale` when csrc/ changed since
  sustained     the same rollout back to back for >= 2 s
  rollout_1m    1,048,576 envs per GPU (is the headline's ALU fraction the kernel or cfg 2's 443 threads per SM?)
  per_step_api  one launch per env step (g2048_env_step_fused as a CUDA graph) at 65,536 and 16,384 envs (cfg 3 env side)
  beam          beam-search nodes/s, width 20 depth 40, batched get_action
  games         BASELINE configs 4 and 5: whole games to game over -- cfg 4 (100 games, 15/20), cfg 5 strong
                (10,000 games over all ranks) and weak (10,000 per rank); the NCCL statistics all-reduce is
                inside the timed region; `checksum` is independent of the GPU count
  cpu_baseline  oracle port on the host cores (bounded sample) and the Python reference (measured where the
                reference tree exists, else the BASELINE.md probe, labelled)
  clocks, gpu_launches
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "env board-steps/sec (random-policy rollout, float64 shaped reward, auto-reset)"
UNIT = "board-steps/s"
ENVS = 65536
ENV_STEPS = 2000
BYTES_PER_STEP = 22          # 8 rd board + 1 rd action + 8 wr board + 4 wr reward + 1 wr done
BYTES_PER_NODE = 21          # 8 rd parent + 8 wr child + 4 wr score + 1 wr first action
SEED = 1234
BEAM_W, BEAM_D, BEAM_ROOTS = 20, 40, 10000
SLEEP_CYCLES = 1_000_000     # torch.cuda._sleep before each timed launch (see run_ours)
REF_SAMPLE = (16384, 200)    # bounded sample of the cfg-2 workload for the CPU arm's later steps


def workload_config(n, env_steps):
    """`config` of both arms: the ours-arm and the reference-arm line describe the same workload."""
    return {"workload": f"cfg2: {n} boards/GPU x {env_steps} random-policy env steps per bench step "
                        f"(on-device / host Philox actions and spawns, float64 shaped reward, auto-reset)",
            "envs_per_gpu": n, "env_steps": env_steps, "seed": SEED}


def csrc_hash():
    """sha256 over the kernel sources, so that ncu-derived constants can be tied to the code they describe."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "2048-using-reinforcement-learning_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(csrc, name), "rb") as f:
                h.update(name.encode()); h.update(f.read())
    return h.hexdigest()[:16]


def ncu_constants():
    """Per-instruction figures of the hot kernels from the committed ncu captures (profiles/summarize.py
    writes the file together with the hash of csrc/ it was captured from)."""
    with open(os.path.join(ROOT, "profiles", "ncu_constants.json")) as f:
        c = json.load(f)
    c["stale"] = c.get("csrc_sha16") != csrc_hash()
    return c


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val == "Active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU side: the oracle port on the host cores (cpu_baseline leg and --impl reference)
# ----------------------------------------------------------------------------------------------
def oracle_env_state(O, n, seed, game0):
    boards = np.zeros((n, 16), np.int32); score = np.zeros(n, np.int64); hi = np.zeros(n, np.int32)
    ctr = np.zeros(n, np.uint32)
    for i in range(n):
        e = O.Env(seed, game0 + i, ctor_reset=False)
        e.reset()
        boards[i] = e.board; hi[i] = e.s.highest_tile; ctr[i] = e.s.spawn_ctr
    return boards, score, hi, ctr


def cpu_env_seconds(O, n, steps, threads):
    """Seconds the oracle takes for n envs x steps on `threads` host threads (building the states is not timed)."""
    st = oracle_env_state(O, n, SEED, 0)
    rs = np.zeros(n, np.float64); ep = np.zeros(n, np.int32)
    t = time.perf_counter()
    O.rollout(st[0], st[1], st[2], st[3], rs, ep, steps, 0, SEED, 0, threads)
    return time.perf_counter() - t


def cpu_beam_sample(O, roots, threads):
    vals = np.stack([O.synthetic_board(SEED, g) for g in range(roots)])
    t = time.perf_counter()
    _, _, nodes, _ = O.beam_batch(vals, BEAM_W, BEAM_D, SEED, 0, 0, threads)
    return int(nodes.sum()) / (time.perf_counter() - t)


def _python_reference_worker(args):
    """cfg 1 loop of the UNMODIFIED reference env on one core: get_valid_moves, a uniform
    legal action, step; reset at done or 2,000 steps; for `seconds`."""
    root, seed, seconds = args
    import random
    sys.dont_write_bytecode = True
    sys.path.insert(0, root)
    from environment.game_2048 import Game2048Env
    random.seed(seed)
    env = Game2048Env()
    env.reset()
    steps = in_game = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        legal = [a for a, ok in enumerate(env.get_valid_moves()) if ok]
        _, _, done, _ = env.step(random.choice(legal) if legal else 0)
        steps += 1; in_game += 1
        if done or in_game >= 2000:
            env.reset(); in_game = 0
    return steps / (time.perf_counter() - t0)


def python_reference_leg(seconds=6.0):
    """The reference's own Python env on every host core (multiprocessing; Python is GIL-bound so cores =
    processes).  Only where the reference tree exists; elsewhere the BASELINE.md probe is quoted, labelled."""
    root = os.environ.get("G2048_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(root, "environment", "game_2048.py")):
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context("spawn").Pool(cores) as pool:
            rates = pool.map(_python_reference_worker, [(root, 100 + i, seconds) for i in range(cores)])
        return {"kind": "reference-python", "value": float(sum(rates)), "unit": "board-steps/s", "cores": cores,
                "per_core": float(np.mean(rates)), "measured": "here",
                "sample": f"cfg 1 loop (get_valid_moves + random legal step), {seconds:.0f} s on each of {cores} processes"}
    return {"kind": "reference-python", "value": 976.0, "unit": "board-steps/s", "cores": 1, "per_core": 976.0,
            "measured": "NOT on this box: BASELINE.md section 2 probe (survey container, one core of an 8-vCPU Xeon); the "
                        "reference is pure Python and its tree does not travel to the GPU box",
            "sample": "cfg 1 loop, one game of 182 steps; 1,564 steps/s with uniform random actions incl. invalid ones"}


def run_reference(args):
    """The reference's algorithm for this path on the host cores.  The reference itself is pure Python and
    does not exist on the GPU box, so this is the C oracle port (kind "port") on every host thread.  The first
    timed step is one FULL pass of the cfg-2 workload (same shape as our arm); further steps are bounded
    samples of it, so that --steps K ends within minutes.  value = board-steps done / seconds taken."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import pyoracle as O
    threads = O.max_threads()
    n_full, steps_full = args.envs, args.env_steps
    n_s, steps_s = REF_SAMPLE
    for _ in range(args.warmup):
        cpu_env_seconds(O, n_s, 10, threads)
    work = secs = 0.0
    for k in range(args.steps):
        n, steps = (n_full, steps_full) if k == 0 else (n_s, steps_s)
        secs += cpu_env_seconds(O, n, steps, threads)
        work += n * steps
    value = work / secs
    sample = (f"step 0: the full {n_full} x {steps_full} workload; steps 1..{args.steps - 1}: {n_s} envs x {steps_s} steps each "
              f"(oracle/orc2048.c, {threads} pthreads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (n_full * steps_full / value), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 packed boards (int32 ALU) + f64 reward", "data": "synthetic",
        "config": workload_config(n_full, steps_full),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "seconds": secs,
        "note": "ms_per_step is the time of one full cfg-2 pass at the measured rate",
    }))


# ----------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import g2048_b200 as G
    from g2048_b200 import _lib

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_over_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_over_ranks(x, dist.ReduceOp.MAX if world > 1 else None)

    def min_over_ranks(x):
        return reduce_over_ranks(x, dist.ReduceOp.MIN if world > 1 else None)

    def event_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    n, env_steps = args.envs, args.env_steps
    env = G.BatchedGame2048Env(n, dev, seed=SEED, game0=rank * n)      # rank r owns games [r*n, (r+1)*n)
    env.reset()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- headline: fused rollout, state resident in HBM ------------------------------------------
    for _ in range(args.warmup):
        env.rollout(env_steps)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = G.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                          # evict L2 between timed iterations
        torch.cuda._sleep(SLEEP_CYCLES)                                # GPU idles ~0.5 ms so the host is always ahead:
        starts[i].record()                                             # no enqueue latency between the two events
        env.rollout(env_steps)
        ends[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = G.launch_count() - launches0
    step_times = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    dev_ms = max_over_ranks(sum(step_times))
    value = world * n * env_steps * args.steps / (dev_ms * 1e-3)
    kernel_ms = dev_ms / args.steps

    # ---- sustained: the same launch back to back for >= 2 s (clocks, power) -----------------------------
    reps = max(8, int(2.5 / (kernel_ms * 1e-3)))
    barrier()
    s0, s1 = event_pair()
    s0.record()
    for _ in range(reps):
        env.rollout(env_steps)
    s1.record()
    barrier()
    sustained_s = max_over_ranks(s0.elapsed_time(s1)) * 1e-3
    sustained_value = world * n * env_steps * reps / sustained_s

    # ---- 1,048,576 envs per GPU: enough threads to fill every scheduler ---------------------------------
    big_n = 1 << 20
    big = G.BatchedGame2048Env(big_n, dev, seed=SEED, game0=(world + rank) * big_n)
    big_steps = 500
    for _ in range(2):
        big.rollout(big_steps)
    barrier()
    s0, s1 = event_pair()
    s0.record()
    for _ in range(4):
        big.rollout(big_steps)
    s1.record()
    barrier()
    big_value = world * big_n * big_steps * 4 / (max_over_ranks(s0.elapsed_time(s1)) * 1e-3)
    del big

    # ---- per-step API: the fused step (step + reset + legal mask + observation) as a CUDA graph -----------
    def graph_steps_per_s(e, want_obs):
        acts = torch.randint(0, 4, (64, e.n), device=dev, dtype=torch.uint8)

        def sixty_four_steps():
            for i in range(64):
                e.step_fused(acts[i], auto_reset=True, want_obs=want_obs)
        g = e.graph(sixty_four_steps)
        g.replay()
        barrier()
        a, b = event_pair()
        a.record()
        for _ in range(8):
            g.replay()
        b.record()
        barrier()
        return world * e.n * 64 * 8 / (max_over_ranks(a.elapsed_time(b)) * 1e-3)

    per_step_65k = graph_steps_per_s(env, False)
    per_step_65k_obs = graph_steps_per_s(env, True)
    cfg3 = G.BatchedGame2048Env(16384, dev, seed=SEED, game0=(3 * world + rank) * 16384)
    per_step_cfg3_obs = graph_steps_per_s(cfg3, True)
    acts = torch.randint(0, 4, (64, n), device=dev, dtype=torch.uint8)
    for i in range(8):
        env.step(acts[i])
    barrier()
    e0, e1 = event_pair()
    e0.record()
    for i in range(64):
        env.step(acts[i])
    e1.record()
    barrier()
    per_step_eager = world * n * 64 / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H per call) ---------------------------
    P = _lib.np_ptr
    lib = _lib.use_device(local_rank)
    pin = lambda dt: torch.empty(n, dtype=dt).pin_memory().numpy()     # noqa: E731
    hb = pin(torch.int64).view(np.uint64); hs = pin(torch.int32); hh = pin(torch.uint8)
    hc = pin(torch.int32).view(np.uint32); hr = pin(torch.float64); he = pin(torch.int32)
    hb[:] = 0; hs[:] = 0; hh[:] = 0; hc[:] = 0; hr[:] = 0; he[:] = 0
    _lib.check(lib.g2048_host_env_reset(P(hb), P(hs), P(hh), P(hc), n, SEED, rank * n))
    h2d = n * (8 + 4 + 1 + 4 + 8 + 4)
    d2h = h2d
    for w in range(max(1, args.warmup)):
        _lib.check(lib.g2048_host_env_rollout(P(hb), P(hs), P(hh), P(hc), P(hr), P(he), n, env_steps, w * env_steps, SEED, rank * n))
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        _lib.check(lib.g2048_host_env_rollout(P(hb), P(hs), P(hh), P(hc), P(hr), P(he), n, env_steps,
                                              (args.warmup + i) * env_steps, SEED, rank * n))
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * env_steps * args.steps / e2e_s
    # one env.step per call with host buffers (the reference's call granularity, batched)
    ha = pin(torch.uint8); ha[:] = np.random.default_rng(0).integers(0, 4, n).astype(np.uint8)
    hrw = pin(torch.float64); hv = pin(torch.uint8); hl = pin(torch.uint8); hd = pin(torch.uint8)
    for _ in range(3):
        _lib.check(lib.g2048_host_env_step(P(hb), P(ha), None, P(hs), P(hh), P(hc), P(hrw), None, P(hv), P(hl), P(hd), n, SEED, rank * n))
    t0 = time.perf_counter()
    for _ in range(20):
        _lib.check(lib.g2048_host_env_step(P(hb), P(ha), None, P(hs), P(hh), P(hc), P(hrw), None, P(hv), P(hl), P(hd), n, SEED, rank * n))
    e2e_step_value = world * n * 20 / max_over_ranks(time.perf_counter() - t0)

    # ---- secondary metric: beam-search nodes/s (width 20, depth 40) on synthetic boards --------------
    roots = torch.empty(args.beam_roots, dtype=torch.int64, device=dev)
    _lib.check(lib.g2048_synthetic_boards(roots.data_ptr(), args.beam_roots, SEED, rank * args.beam_roots,
                                          torch.cuda.current_stream().cuda_stream))
    search = G.BatchedBeamSearch(BEAM_W, BEAM_D, dev, seed=SEED)
    for w in range(max(1, args.warmup)):
        search.get_actions(roots, call=w, game0=rank * args.beam_roots)
    barrier()
    bs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    be = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    outs = [search.new_outputs(args.beam_roots) for _ in range(args.steps)]     # no allocation inside the timed region
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        torch.cuda._sleep(SLEEP_CYCLES)
        bs[i].record()
        search.get_actions(roots, call=100 + i, game0=rank * args.beam_roots, out=outs[i])
        be[i].record()
    barrier()
    beam_times = [s.elapsed_time(e) for s, e in zip(bs, be)]
    beam_ms = max_over_ranks(sum(beam_times))
    nodes_total = sum(int(o["nodes"].sum().item()) for o in outs)
    if world > 1:
        t = torch.tensor([nodes_total], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        nodes_total = int(t.item())
    beam_value = nodes_total / (beam_ms * 1e-3)
    # beam e2e: pinned host roots in, host actions / probabilities / node counts out, every call
    hroots_t = roots.cpu().pin_memory()
    ha8_t = torch.zeros(args.beam_roots, dtype=torch.uint8).pin_memory()
    hp_t = torch.zeros(args.beam_roots, dtype=torch.float32).pin_memory()
    hk_t = torch.zeros(args.beam_roots, dtype=torch.int32).pin_memory()
    hroots, ha8, hp, hk = hroots_t.numpy().view(np.uint64), ha8_t.numpy(), hp_t.numpy(), hk_t.numpy()

    def beam_host_call(call0):
        _lib.check(lib.g2048_host_beam_search(P(hroots), None, None, call0, P(ha8), P(hp), None, P(hk), args.beam_roots,
                                              BEAM_W, BEAM_D, 512, 1024, SEED, rank * args.beam_roots))
        return int(hk.sum())

    beam_host_call(99)                                   # warm-up: staging arena
    barrier()
    t0 = time.perf_counter()
    e2e_nodes = sum(beam_host_call(100 + i) for i in range(args.steps))
    beam_e2e = world * e2e_nodes / max_over_ranks(time.perf_counter() - t0)

    # ---- whole games: BASELINE configs 4 and 5 --------------------------------------------------------
    def play(total_games, W, D, lo, hi):
        """Games [lo, hi) of a run of `total_games`, to game over (cap 10,000 moves, evaluate_beam_search.py:16),
        timed on the device from the launch to the end of the NCCL statistics all-reduce."""
        s = G.BatchedBeamSearch(W, D, dev, seed=SEED)
        s.play_games(min(hi - lo, 8), max_moves=40, game0=lo)           # warm-up: attributes, allocator pools
        barrier()
        a, b = event_pair()
        a.record()
        out = s.play_games(hi - lo, max_moves=10000, game0=lo)
        stats = G.all_reduce_stats(out["stats"])
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        # order-independent checksum of the per-game results: identical for every GPU count
        ck = (out["score"].to(torch.int64) * 1000003 + out["moves"].to(torch.int64) * 10007 +
              out["highest_exp"].to(torch.int64) * 101 + out["invalid"].to(torch.int64) * 7 +
              out["nodes"] * 3).sum().reshape(1)
        if world > 1:
            dist.all_reduce(ck)
        st = G.describe_stats(stats)
        seconds = max_over_ranks(ms) * 1e-3
        return {"games": st["games"], "beam_width": W, "search_depth": D, "seconds": seconds,
                "rank_seconds_min_max": [min_over_ranks(ms) * 1e-3, seconds],
                "nodes_per_s": st["nodes"] / seconds, "moves_per_s": st["average_moves"] * st["games"] / seconds,
                "games_per_s": st["games"] / seconds, "nodes": st["nodes"], "average_score": st["average_score"],
                "max_score": st["max_score"], "highest_tile_histogram": st["highest_tile_histogram"],
                "invalid_moves": st["invalid_moves"], "checksum": int(ck.item()) & (2**63 - 1)}

    games = {}
    lo, hi = G.shard_range(100, rank, world)
    games["cfg4_100_games_15_20"] = dict(play(100, 15, 20, lo, hi), scaling="strong")
    lo, hi = G.shard_range(args.games, rank, world)
    games["cfg5_strong"] = dict(play(args.games, 20, 40, lo, hi), scaling="strong",
                                note=f"{args.games} games over all ranks, rank r plays games [g*r/R, g*(r+1)/R)")
    if world > 1:
        games["cfg5_weak"] = dict(play(args.games * world, 20, 40, rank * args.games, (rank + 1) * args.games),
                                  scaling="weak", note=f"{args.games} games per rank")
    else:
        games["cfg5_weak"] = dict(games["cfg5_strong"], scaling="weak", note="N=1: the strong-scaling run")
    # a lone game: the sequential chain that bounds strong scaling (us per move)
    lone = G.BatchedBeamSearch(20, 40, dev, seed=SEED)
    lone.play_games(1, max_moves=50, game0=7, stats=False)
    barrier()
    a, b = event_pair()
    a.record()
    o = lone.play_games(1, max_moves=10000, game0=7, stats=False)
    b.record()
    barrier()
    games["lone_game_20_40_us_per_move"] = a.elapsed_time(b) * 1e3 / max(1, int(o["moves"][0]))

    clocks = sampler.stop() if rank == 0 else None      # sampled across all GPU timed regions above

    # final histogram all-reduce of the rollout envs (not on the hot path)
    hist = torch.bincount(env.highest_exp.to(torch.int64), minlength=18)
    if world > 1:
        dist.all_reduce(hist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: oracle port on the host cores, bounded sample ---------------------------------------
    from oracle import pyoracle as O
    threads = O.max_threads()
    cpu_n, cpu_steps, cpu_roots = 16384, 1500, 16384         # ~15-25 core-seconds in total
    if world == 1:                                           # the CPU arm is timed at N=1 only
        cpu_value = cpu_n * cpu_steps / cpu_env_seconds(O, cpu_n, cpu_steps, threads)
        cpu_beam = cpu_beam_sample(O, cpu_roots, threads)
        pyref = python_reference_leg()
    else:
        cpu_value = cpu_beam = pyref = None

    C = ncu_constants()
    R, B = C["rollout"], C["beam"]
    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_STEP * (n * env_steps) / (kernel_ms * 1e-3) / 1e9       # per launch, one rank
    beam_achieved = BYTES_PER_NODE * (nodes_total / world / args.steps) / (beam_ms / args.steps * 1e-3) / 1e9
    # ALU/issue view of the same launch: warp instructions issued per second over the SM issue peak
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6 if clocks else 1965.0e6
    issue_peak = 148 * 4 * sm_hz                                   # 1 warp-instruction / clk / scheduler
    issue_rate = (value / world) / 32.0 * R["warp_inst_per_warp_step"]
    alu_rate = (value / world) / 32.0 * R["alu_warp_inst_per_warp_step"]
    alu_peak = C["int32_peaks"]["alu_pipe_warp_inst_per_s"]
    not_n1 = "not timed at N > 1 (see the N=1 line)"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 packed boards (int32 ALU) + f64 reward", "data": "synthetic",
        "config": dict(workload_config(n, env_steps),
                       l2="flushed between timed iterations (256 MiB write); the 512 KiB working set is re-read from HBM",
                       timing="CUDA events around each rollout launch on the launching stream, max over ranks",
                       wall_s_bracket=wall, ms_min_max=[min(step_times), max(step_times)]),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "g2048_host_env_rollout (pinned host state in/out per call)",
                "per_env_step_call": {"value": e2e_step_value, "unit": UNIT,
                                      "api": "g2048_host_env_step, 65,536 boards per call, all outputs"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": R["dram_bytes_per_launch"], "kernel": "env_rollout_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_STEP * n * env_steps,
                     "note": "algorithmic 22 B per board-step (SURVEY 8d) as if every step round-tripped HBM; the fused "
                             "rollout keeps boards in registers (measured DRAM traffic = `traffic`), so the binding "
                             "roofline is ALU-pipe issue, reported in `issue`"},
        "issue": {"bound": "alu", "achieved": issue_rate, "peak": issue_peak, "unit": "warp-inst/s",
                  "frac": issue_rate / issue_peak, "ncu": R, "stale": C["stale"], "constants": "profiles/ncu_constants.json",
                  "csrc_sha16": {"captured": C.get("csrc_sha16"), "built": csrc_hash()},
                  "alu_pipe": {"achieved": alu_rate, "peak": alu_peak, "frac": alu_rate / alu_peak, "unit": "warp-inst/s",
                               "peak_source": C["int32_peaks"]["source"]},
                  "measured_issue_peak": C["int32_peaks"]["alu_plus_fma_warp_inst_per_s"]},
        "sustained": {"value": sustained_value, "unit": UNIT, "seconds": sustained_s, "launches": reps,
                      "note": "the headline launch back to back, no L2 flush, no idle gaps"},
        "rollout_1m": {"value": big_value, "unit": UNIT, "envs_per_gpu": big_n, "env_steps": big_steps,
                       "alu_pipe_frac": (big_value / world) / 32.0 * R["alu_warp_inst_per_warp_step"] / alu_peak,
                       "note": "16x the envs of cfg 2: every scheduler has all the warps it can hold"},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{cpu_n} envs x {cpu_steps} steps, oracle/orc2048.c on {threads} threads"
                                   if world == 1 else not_n1,
                         "python_reference": pyref if world == 1 else not_n1},
        "per_step_api": {"metric": "one launch per env step: step + reset of finished games + legal mask (+ float32[N,16] "
                                   "observation) in g2048_env_step_fused, 64 steps per CUDA graph",
                         "cuda_graph_65536": {"value": per_step_65k, "unit": UNIT, "us_per_step": world * n / per_step_65k * 1e6},
                         "cuda_graph_65536_with_obs": {"value": per_step_65k_obs, "unit": UNIT},
                         "cfg3_env_side_16384_with_obs": {"value": per_step_cfg3_obs, "unit": UNIT,
                                                         "us_per_step": world * 16384 / per_step_cfg3_obs * 1e6},
                         "eager_65536": {"value": per_step_eager, "unit": UNIT, "note": "g2048_env_step from Python, no graph"}},
        "beam": {"metric": "beam-search nodes/sec (BeamSearchAgent.get_action, width 20 depth 40)", "value": beam_value,
                 "unit": "nodes/s", "roots_per_gpu": args.beam_roots, "nodes_per_step": nodes_total // max(1, args.steps),
                 "ms_per_step": beam_ms / args.steps, "ms_min_max": [min(beam_times), max(beam_times)],
                 "e2e": {"value": beam_e2e, "unit": "nodes/s", "api": "g2048_host_beam_search (pinned host roots in, host results out per call)",
                         "h2d_bytes_per_step": 8 * args.beam_roots, "d2h_bytes_per_step": 9 * args.beam_roots},
                 "roofline": {"bound": "hbm", "achieved": beam_achieved, "peak": peak, "unit": "GB/s",
                              "frac": beam_achieved / peak, "traffic": B["dram_bytes_per_launch"],
                              "kernel": "beam_search_kernel", "ncu": B, "stale": C["stale"]},
                 "cpu_baseline": {"value": cpu_beam, "unit": "nodes/s", "cores": threads, "kind": "port",
                                  "sample": f"{cpu_roots} synthetic roots, oracle/orc2048.c on {threads} threads"
                                            if world == 1 else not_n1}},
        "games": games,
        "highest_tile_histogram": {str(1 << e): int(c) for e, c in enumerate(hist.tolist()) if c},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=ENVS)
    ap.add_argument("--env-steps", type=int, default=ENV_STEPS)
    ap.add_argument("--beam-roots", type=int, default=BEAM_ROOTS)
    ap.add_argument("--games", type=int, default=10000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
