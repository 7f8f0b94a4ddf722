"""Executed warp instructions per source line split by pipe class, from two ncu source-page dumps:
  ncu -i rep --page source --print-source cuda,sass --csv > lines.csv   (line -> SASS addresses)
  ncu -i rep --page source --csv > sass.csv                             (address -> executed count)
usage: python profiles/alu_by_line.py lines.csv sass.csv <units> [min]"""
import collections, csv, sys
lines_csv, sass_csv, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
ALU = ("LOP3", "SHF", "SEL", "PRMT", "ISETP", "IADD3", "VIADD", "VIMNMX", "LEA", "FSEL", "FSETP", "MOV", "IABS", "FLO", "BREV", "PLOP3", "FMNMX", "SGXT", "BMSK", "P2R", "R2P")
def klass(op):
    base = op.split(".")[0]
    if base in ALU: return "alu"
    if base in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2"): return "fma"
    if base in ("POPC", "I2F", "F2I", "MUFU", "F2F", "I2FP"): return "xu"
    if base in ("DADD", "DMUL", "DFMA", "DSETP"): return "fp64"
    if base in ("LDS", "STS", "LDG", "STG", "LDC", "LDCU", "ATOMG", "RED"): return "mem"
    return "other"
count = {}
for r in csv.reader(open(sass_csv)):
    if len(r) > 5 and r[0].startswith("0x"):
        ins = r[1].strip()
        if ins.startswith("@"): ins = ins.split(None, 1)[1]
        count[r[0]] = (ins.split()[0].rstrip(";"), int(r[5]))
agg = collections.defaultdict(lambda: collections.Counter())
fname = line = src = None
for r in csv.reader(open(lines_csv)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0].strip().isdigit(): line = int(r[0]); src = r[1].strip()[:90]; continue
    if len(r) > 2 and r[2].startswith("0x") and r[2] in count:
        op, n = count[r[2]]
        agg[(fname, line, src)][klass(op)] += n
tot = collections.Counter()
for k, c in agg.items(): tot.update(c)
print("per unit:", {k: round(v / units, 1) for k, v in tot.items()})
for (f, ln, src), c in sorted(agg.items()):
    if c["alu"] / units >= thr:
        print(f"alu {c['alu']/units:6.1f} fma {c['fma']/units:5.1f} other {(sum(c.values())-c['alu']-c['fma'])/units:5.1f}  {f}:{ln:<4d} {src}")
