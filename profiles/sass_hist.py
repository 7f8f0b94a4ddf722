"""Opcode histogram of an `ncu --page source --csv` dump, weighted by executed warp instructions.
usage: python profiles/sass_hist.py <source.csv> <warp_units>   (warp_units = e.g. warp-steps in the launch)"""
import csv, sys, collections
path, units = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[1]
i_src, i_exec = hdr.index("Source"), hdr.index("Instructions Executed")
i_samp = hdr.index("# Samples")
hist, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= i_exec: continue
    ins = r[i_src].strip()
    if ins.startswith("@"): ins = ins.split(None, 1)[1]
    op = ins.split()[0].rstrip(";")
    base = op.split(".")[0]
    n = int(r[i_exec]); hist[base] += n; tot += n; samp[base] += int(r[i_samp])
print(f"total warp-instructions {tot}  per unit {tot/units:.1f}")
for op, n in hist.most_common(40):
    print(f"{op:12s} {n/units:8.2f} per unit  {100*n/tot:5.1f}%   samples {samp[op]}")
