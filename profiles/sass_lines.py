"""Static SASS instruction counts per source line of one kernel (no GPU needed).
usage: python profiles/sass_lines.py <lib.so> <kernel-name-substring> [min_count]
Extracts the cubins with cuobjdump, disassembles with `nvdisasm --print-line-info` and prints, per source line, how
many SASS instructions carry it (the per-step kernel has no hot loops, so static counts track the executed ones)."""
import collections, glob, os, re, subprocess, sys, tempfile
lib, key = os.path.abspath(sys.argv[1]), sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    text = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    cur, line, take = None, None, False
    per_line, per_op, total = collections.Counter(), collections.Counter(), 0
    for l in text.splitlines():
        if l.startswith(".text."):
            take = key in l
            if take: print("kernel", l.strip())
            continue
        if not take: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
        if m and not m.group(1).startswith("NOP"):
            per_line[line] += 1; per_op[m.group(1).split(".")[0]] += 1; total += 1
    if total:
        print("total", total)
        for (f, n), c in sorted(per_line.items()):
            if c >= thr: print(f"{c:5d}  {f}:{n}")
        print("opcodes:", ", ".join(f"{o} {c}" for o, c in per_op.most_common(14)))
