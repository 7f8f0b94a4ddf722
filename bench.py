#!/usr/bin/env python
"""bench.py -- env board-steps/s (headline) and beam-search nodes/s on B200, beside the CPU oracle.

    python bench.py --gpus N --steps K --warmup W            # our CUDA engine, one process per GPU
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

Workload (BASELINE.json configs[1]): 65,536 boards per GPU x 2,000 random-policy env steps.
One bench "step" = one pass of that workload = one fused rollout launch (131,072,000 board-steps
per GPU).  Boards are independent, so N GPUs each run their own 65,536 boards (weak scaling, no
collective on the data path); `value` = all ranks' board-steps / max-over-ranks device time.

The JSON line also carries: `e2e` (same metric through the host-buffer C ABI, H2D + D2H inside
the timed region), `roofline` (algorithmic HBM bytes of SURVEY 8d: 22 B per board-step),
`cpu_baseline` (oracle port on the host cores, bounded sample), `beam` (secondary metric:
beam-search nodes/s, width 20 depth 40), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
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
# From the committed ncu capture of the same command (profiles/ncu_summary_r01.md): executed warp
# instructions per warp-step of env_rollout_kernel, its DRAM traffic per launch, and pipe utilisation.
NCU_ROLLOUT = {"warp_inst_per_warp_step": 336.5, "dram_bytes_per_launch": 2157824, "alu_pipe_pct_of_peak": 67.4,
               "issue_active_pct": 69.6, "fma_pipe_pct_of_peak": 15.4, "source": "profiles/ncu_summary_r01.md",
               # ALU-pipe instructions per warp-step = pct_of_peak x 0.5 inst/clk/SMSP x SMSP cycles per warp-step
               "alu_warp_inst_per_warp_step": 186.3}
# Measured pipe peaks of this pool's B200 (profiles/int32_peak.cu -> profiles/int32_peak_r01.json): a pure
# LOP3/SHF/PRMT stream sustains 5.79e11 warp-inst/s (0.5 per clock per scheduler), an ALU+IMAD mix 1.13e12.
INT32_PEAKS = {"alu_pipe_warp_inst_per_s": 5.79e11, "alu_plus_fma_warp_inst_per_s": 1.13e12,
               "source": "profiles/int32_peak_r01.json"}
NCU_BEAM = {"alu_pipe_pct_of_peak": 73.0, "issue_active_pct": 69.8, "dram_bytes_per_launch": 272896,
            "source": "profiles/ncu_summary_r01.md"}


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


def cpu_env_sample(O, n, steps, threads, return_seconds=False):
    """Board-steps/s of the oracle on `threads` host threads over n envs x steps (same workload, smaller)."""
    st = oracle_env_state(O, n, SEED, 0)
    rs = np.zeros(n, np.float64); ep = np.zeros(n, np.int32)
    t = time.perf_counter()
    O.rollout(st[0], st[1], st[2], st[3], rs, ep, steps, 0, SEED, 0, threads)
    dt = time.perf_counter() - t
    return (n * steps / dt) if not return_seconds else dt


def cpu_beam_sample(O, roots, threads):
    vals = np.stack([O.synthetic_board(SEED, g) for g in range(roots)])
    t = time.perf_counter()
    _, _, nodes, _ = O.beam_batch(vals, BEAM_W, BEAM_D, SEED, 0, 0, threads)
    return int(nodes.sum()) / (time.perf_counter() - t)


def run_reference(args):
    """The reference's algorithm for this path on the host cores.  The reference itself is pure
    Python and does not exist on the GPU box, so this is the C oracle port (kind "port")."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import pyoracle as O
    threads = O.max_threads()
    n, steps = 16384, 200                     # bounded sample of the 65,536 x 2,000 workload per step
    for _ in range(args.warmup):
        cpu_env_sample(O, n, 10, threads)
    dt = 0.0                                  # the rollout itself; building the initial states is not timed
    for _ in range(args.steps):
        dt += cpu_env_sample(O, n, steps, threads, return_seconds=True)
    value = n * steps * args.steps / dt
    sample = f"{n} envs x {steps} steps per bench step (oracle/orc2048.c, pthreads)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 boards / f64 reward", "data": "synthetic",
        "config": {"workload": f"cfg2 sample: {sample}", "seed": SEED},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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

    # ---- per-step API (one launch per env step), device-resident, for reference --------------------
    acts = torch.randint(0, 4, (64, n), device=dev, dtype=torch.uint8)
    for i in range(8):
        env.step(acts[i])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(64):
        env.step(acts[i])
    e1.record()
    barrier()
    per_step_api = max_over_ranks(e0.elapsed_time(e1)) * 1e-3
    per_step_api_value = world * n * 64 / per_step_api
    # the same 64 per-step launches replayed as one CUDA graph (what a PPO loop would do)
    def sixty_four_steps():
        for i in range(64):
            env.step(acts[i], auto_reset=True)
    graph = env.graph(sixty_four_steps)
    graph.replay()
    barrier()
    e0.record()
    for _ in range(4):
        graph.replay()
    e1.record()
    barrier()
    per_step_graph_value = world * n * 64 * 4 / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3)

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
        out = search.get_actions(roots, call=w, game0=rank * args.beam_roots)
    barrier()
    bs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    be = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    nodes_total = 0
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

    clocks = sampler.stop() if rank == 0 else None      # sampled across all GPU timed regions above

    # final histogram all-reduce (the only collective of the workload; not on the hot path)
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
        cpu_value = cpu_env_sample(O, cpu_n, cpu_steps, threads)
        cpu_beam = cpu_beam_sample(O, cpu_roots, threads)
    else:
        cpu_value = cpu_beam = None

    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_STEP * (n * env_steps) / (kernel_ms * 1e-3) / 1e9       # per launch, one rank
    beam_achieved = BYTES_PER_NODE * (nodes_total / world / args.steps) / (beam_ms / args.steps * 1e-3) / 1e9
    # ALU/issue view of the same launch: warp instructions issued per second over the SM issue peak
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6 if clocks else 1965.0e6
    issue_peak = 148 * 4 * sm_hz                                   # 1 warp-instruction / clk / scheduler
    issue_rate = (value / world) / 32.0 * NCU_ROLLOUT["warp_inst_per_warp_step"]
    alu_rate = (value / world) / 32.0 * NCU_ROLLOUT["alu_warp_inst_per_warp_step"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 packed boards (int32 ALU) + f64 reward", "data": "synthetic",
        "config": {"workload": f"cfg2: {n} boards/GPU x {env_steps} random-policy env steps per bench step "
                               f"(fused rollout launch, on-device Philox actions and spawns, auto-reset)",
                   "envs_per_gpu": n, "env_steps": env_steps, "seed": SEED,
                   "l2": "flushed between timed iterations (256 MiB write); the 512 KiB working set is re-read from HBM",
                   "timing": "CUDA events around each rollout launch on the launching stream, max over ranks",
                   "wall_s_bracket": wall, "ms_min_max": [min(step_times), max(step_times)]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "g2048_host_env_rollout (pinned host state in/out per call)",
                "per_env_step_call": {"value": e2e_step_value, "unit": UNIT,
                                      "api": "g2048_host_env_step, 65,536 boards per call, all outputs"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_ROLLOUT["dram_bytes_per_launch"], "kernel": "env_rollout_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_STEP * n * env_steps,
                     "note": "algorithmic 22 B per board-step (SURVEY 8d) as if every step round-tripped HBM; the fused "
                             "rollout keeps boards in registers (measured DRAM traffic = `traffic`), so the binding "
                             "roofline is ALU-pipe issue, reported in `issue`"},
        "issue": {"bound": "alu", "achieved": issue_rate, "peak": issue_peak, "unit": "warp-inst/s",
                  "frac": issue_rate / issue_peak, "ncu": NCU_ROLLOUT,
                  "alu_pipe": {"achieved": alu_rate, "peak": INT32_PEAKS["alu_pipe_warp_inst_per_s"],
                               "frac": alu_rate / INT32_PEAKS["alu_pipe_warp_inst_per_s"], "unit": "warp-inst/s",
                               "peak_source": INT32_PEAKS["source"]},
                  "measured_issue_peak": INT32_PEAKS["alu_plus_fma_warp_inst_per_s"]},
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{cpu_n} envs x {cpu_steps} steps, oracle/orc2048.c on {threads} threads"
                                   if world == 1 else "not timed at N > 1 (see the N=1 line)"},
        "per_step_api": {"value": per_step_api_value, "unit": UNIT,
                         "note": "g2048_env_step, one launch per env step, device-resident tensors",
                         "cuda_graph": {"value": per_step_graph_value, "unit": UNIT,
                                        "note": "64 x g2048_env_step_autoreset (step + reset of finished envs) captured in one CUDA graph"}},
        "beam": {"metric": "beam-search nodes/sec (BeamSearchAgent.get_action, width 20 depth 40)", "value": beam_value,
                 "unit": "nodes/s", "roots_per_gpu": args.beam_roots, "nodes_per_step": nodes_total // max(1, args.steps),
                 "ms_per_step": beam_ms / args.steps, "ms_min_max": [min(beam_times), max(beam_times)],
                 "e2e": {"value": beam_e2e, "unit": "nodes/s", "api": "g2048_host_beam_search (pinned host roots in, host results out per call)",
                         "h2d_bytes_per_step": 8 * args.beam_roots, "d2h_bytes_per_step": 9 * args.beam_roots},
                 "roofline": {"bound": "hbm", "achieved": beam_achieved, "peak": peak, "unit": "GB/s",
                              "frac": beam_achieved / peak, "traffic": NCU_BEAM["dram_bytes_per_launch"],
                              "kernel": "beam_search_kernel", "ncu": NCU_BEAM},
                 "cpu_baseline": {"value": cpu_beam, "unit": "nodes/s", "cores": threads, "kind": "port",
                                  "sample": f"{cpu_roots} synthetic roots, oracle/orc2048.c on {threads} threads"
                                            if world == 1 else "not timed at N > 1 (see the N=1 line)"}},
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
