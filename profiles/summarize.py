"""Turns ncu reports / launch lists brought back in gpurun_out/ into the small text summaries
committed under profiles/ (the .ncu-rep files themselves are scratch).

usage: python profiles/summarize.py <round-tag>     e.g. r01
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_summary(rep, label, units=None):
    rows = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, unit, val = rows[0], rows[1], rows[2]
    lines = [f"## {label}", f"report: {os.path.basename(rep)} (ncu --set full --clock-control none, one launch)", ""]
    d = dict(zip(hdr, zip(unit, val)))
    lines.append(f"kernel: {d['Kernel Name'][1]}")
    for k in RAW_KEYS:
        if k in d:
            lines.append(f"{k:75s} {d[k][1]:>16s} {d[k][0]}")
    stalls = sorted(((float(v[1] or 0), k) for k, v in d.items()
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k), reverse=True)
    lines.append("")
    lines.append("warp stall reasons (warps per issue-active cycle), top 8:")
    for v, k in stalls[:8]:
        lines.append(f"  {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
    if units:
        inst = float(d["smsp__inst_executed.sum"][1])
        lines.append("")
        lines.append(f"warp instructions per {units[0]}: {inst / units[1]:.1f}   ({units[2]})")
    return lines, d


def opcode_hist(rep, per, per_name):
    rows = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv"]).splitlines()))
    hdr = rows[1]
    i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
    hist = collections.Counter()
    for r in rows[2:]:
        if len(r) <= i_ex:
            continue
        ins = r[i_src].strip()
        if ins.startswith("@"):
            ins = ins.split(None, 1)[1]
        hist[ins.split()[0].rstrip(";").split(".")[0]] += int(r[i_ex])
    tot = sum(hist.values())
    lines = ["", f"executed warp instructions by opcode, per {per_name} (top 16):"]
    for op, n in hist.most_common(16):
        lines.append(f"  {op:10s} {n / per:8.2f}  {100 * n / tot:5.1f}%")
    return lines


def launch_list(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) > iv:
            try:
                v = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            k = r[ik].split("(")[0]
            agg[k][0] += 1
            agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"## launch list ({os.path.basename(path)}: ncu --metrics gpu__time_duration.sum --clock-control none, `python bench.py --steps 5 --warmup 3`)",
             "per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
             f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append(f"{k[:72]:72s} {n:8d} {t / 1e6:10.3f} {t / n / 1e3:9.2f} {100 * t / tot:6.1f}%")
    return lines


def stall_by_hot_loop(rep, lo, hi):
    """Stall-reason shares over the instructions whose execution count lies in [lo, hi]: picks the level loop of the
    ONE active team out of a team_games_kernel capture whose other 147 blocks only wait (profiles/ncu_cases.py lone)."""
    rows = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv"]).splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    n = inst = samples = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        ex = int(r[ix["Instructions Executed"]])
        if lo <= ex <= hi:
            n += 1; inst += ex; samples += int(r[ix["# Samples"]])
            for k in stalls:
                tot[k] += int(r[ix[k]])
    lines = [f"instructions of the level loop (executed {lo}..{hi} times): {n}; warp-instructions executed {inst}; samples {samples}",
             "warp state of the team's warps over those samples:"]
    for k, v in tot.most_common(8):
        lines.append(f"  {k:28s} {100 * v / max(1, samples):5.1f} %")
    return lines, inst


def csrc_hash():
    sys.path.insert(0, ROOT)
    import bench
    return bench.csrc_hash()


def main():
    text = [f"# ncu summaries, round {tag} (B200, sm_100a)", ""]
    consts = {"csrc_sha16": csrc_hash(), "round": tag,
              "int32_peaks": {"alu_pipe_warp_inst_per_s": 5.79e11, "alu_plus_fma_warp_inst_per_s": 1.13e12,
                              "source": "profiles/int32_peak_r01.json"}}
    ll = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(ll):
        text += launch_list(ll) + [""]
    n_steps = 65536 * 2000 / 32
    rep = os.path.join(OUT, f"prof_rollout_{tag}.ncu-rep")
    if os.path.exists(rep):
        lines, d = raw_summary(rep, "env_rollout_kernel (headline kernel)", ("board-step (warp-step = 32 board-steps)", n_steps, "65,536 envs x 2,000 steps per launch"))
        text += lines + opcode_hist(rep, n_steps, "warp-step") + [""]
        f = lambda k: float(d[k][1].replace(",", ""))          # noqa: E731
        unit = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
        alu = f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
        smsp_cycles = f("sm__cycles_elapsed.max") * 148 * 4 / n_steps
        consts["rollout"] = {
            "warp_inst_per_warp_step": f("smsp__inst_executed.sum") / n_steps,
            "alu_warp_inst_per_warp_step": alu / 100 * 0.5 * smsp_cycles,
            "dram_bytes_per_launch": int((f("dram__bytes_read.sum") * unit[d["dram__bytes_read.sum"][0]]) +
                                         (f("dram__bytes_write.sum") * unit[d["dram__bytes_write.sum"][0]])),
            "alu_pipe_pct_of_peak": alu, "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct_of_peak": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "source": f"profiles/ncu_summary_{tag}.md"}
    rep = os.path.join(OUT, f"prof_beam_{tag}.ncu-rep")
    if os.path.exists(rep):
        lines, d = raw_summary(rep, "beam_search_kernel (width 20, depth 40, 10,000 roots, one warp per root)")
        text += lines + [""]
        f = lambda k: float(d[k][1].replace(",", ""))          # noqa: E731
        unit = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
        consts["beam"] = {"alu_pipe_pct_of_peak": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                          "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "dram_bytes_per_launch": int((f("dram__bytes_read.sum") * unit[d["dram__bytes_read.sum"][0]]) +
                                                       (f("dram__bytes_write.sum") * unit[d["dram__bytes_write.sum"][0]])),
                          "source": f"profiles/ncu_summary_{tag}.md"}
    rep = os.path.join(OUT, f"prof_step_{tag}.ncu-rep")
    if os.path.exists(rep):
        text += ["## note on the env_step_fused_kernel captures",
                 "Up to 65,536 envs the product launches this kernel as ONE-WARP blocks with programmatic dependent launch: the",
                 "blocks of step t+1 become resident and copy their 4 KiB of tables (seven dependent L2 round trips per warp)",
                 "while step t still computes.  ncu serialises launches, so in the two default-policy captures below that",
                 "prologue stands alone in front of every warp (long_scoreboard 3-9, 9-11 us per launch) although the stream",
                 "never sees it: the same launch costs 4.0 us per step in a CUDA graph (bench.py `per_step_api`,",
                 "profiles/step_sizes_r02.json).  The third capture runs the kernel in its plain form (G2048_TUNE_PDL = 0: one",
                 "block of 14 warps per SM, tables published inside the step), which is what ncu can measure faithfully.", ""]
        lines, d = raw_summary(rep, "env_step_fused_kernel<table-free move> (per-step API, 65,536 envs per launch, all outputs incl. observation)")
        text += lines + [""]
    rep = os.path.join(OUT, f"prof_step_warm_{tag}.ncu-rep")
    if os.path.exists(rep):
        lines, d = raw_summary(rep, "env_step_fused_kernel, the same launch with --cache-control none (state and tables L2-resident, as inside "
                                    "a training loop / CUDA graph; the capture above starts from flushed caches)")
        text += lines + [""]
    rep = os.path.join(OUT, f"prof_step_plain_{tag}.ncu-rep")
    if os.path.exists(rep):
        lines, d = raw_summary(rep, "env_step_fused_kernel in its plain launch form (G2048_TUNE_PDL = 0: 147 blocks of 14 warps, no programmatic "
                                    "dependent launch), --cache-control none", ("warp-step (32 board-steps)", 65536 / 32, "65,536 envs per launch"))
        text += lines + [""]
    for kern, what in (("play", "play_games_kernel, cfg 5 on one GPU (10,000 games at 20/40, one warp per game, 24 warps per SM, until 888 games are "
                                "left): the throughput phase of whole-game runs, 51 % of the bench's GPU time"),
                       ("team", "team_games_kernel, the same run's tail (888 games and fewer, one team of four warps per game, stall ranges "
                                "on the free SMs): chain-bound, so low average utilisation is the point (DESIGN.md 4.5)")):
        rep = os.path.join(OUT, f"prof_games_{kern}_{tag}.ncu-rep")
        if os.path.exists(rep):
            lines, d = raw_summary(rep, what)
            text += lines + [""]
    rep = os.path.join(OUT, f"prof_lone_{tag}.ncu-rep")
    if os.path.exists(rep):
        text += ["## team_games_kernel, ONE game (400 moves at 20/40): the sequential chain that bounds whole-game runs",
                 f"report: {os.path.basename(rep)}; 147 of the 148 blocks only wait (they would break stalls), so the kernel-level",
                 "metrics say nothing about the team; the per-instruction samples of its level loop do:", ""]
        lines, inst = stall_by_hot_loop(rep, 20000, 60000)
        text += lines + [""]
    path = os.path.join(ROOT, "profiles", f"ncu_summary_{tag}.md")
    with open(path, "w") as f:
        f.write("\n".join(text) + "\n")
    print("wrote", path)
    if "rollout" in consts and "beam" in consts:
        cpath = os.path.join(ROOT, "profiles", "ncu_constants.json")
        with open(cpath, "w") as f:
            json.dump(consts, f, indent=1)
        print("wrote", cpath, "for csrc", consts["csrc_sha16"])


if __name__ == "__main__":
    main()
