"""Per-source-line executed warp instructions from `ncu --page source --print-source cuda,sass --csv`.
usage: python profiles/line_hist.py <cuda_sass.csv> <warp_units> [min_per_unit]"""
import csv, sys
path, units = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
fname = None; out = []
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0].strip().isdigit() and len(r) > 8:
        try: n = int(r[7])
        except ValueError: continue
        out.append((n / units, fname, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out)
print(f"total per unit {tot:.1f}")
for n, f, ln, src in sorted(out, key=lambda o: (o[1], o[2])):
    if n >= thr: print(f"{n:7.1f}  {f}:{ln:<4d} {src}")
