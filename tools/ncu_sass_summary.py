"""Summarise the SASS source page of an ncu report: executed warp-instructions and stall samples per opcode, and the
hottest instructions.   usage: python tools/ncu_sass_summary.py <report.ncu-rep | source.csv> [units] [top]
`units` divides the executed counts (e.g. problem-iterations in the launch) to print per-unit figures."""
import collections
import csv
import subprocess
import sys

src = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
text = open(src).read() if src.endswith(".csv") else subprocess.run(
    ["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.defaultdict(lambda: [0, 0, collections.Counter()])
insts = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.rstrip(";")
    ex, smp = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
    st = {s: int(r[ix[s]] or 0) for s in stalls}
    b = by_op[op.split(".")[0]]
    b[0] += ex
    b[1] += smp
    b[2].update(st)
    insts.append((smp, ex, sass, st))
tot_ex = sum(b[0] for b in by_op.values())
tot_s = sum(b[1] for b in by_op.values())
print(f"total executed {tot_ex} warp-instr = {tot_ex/units:.1f} per unit; samples {tot_s}")
print(f"{'opcode':12s} {'exec/unit':>10s} {'%exec':>6s} {'%samples':>8s}  top stalls")
for op, (ex, smp, st) in sorted(by_op.items(), key=lambda kv: -kv[1][1])[:22]:
    tops = ", ".join(f"{k[6:]} {100*v/max(1,smp):.0f}%" for k, v in st.most_common(3))
    print(f"{op:12s} {ex/units:10.1f} {100*ex/tot_ex:6.1f} {100*smp/tot_s:8.1f}  {tops}")
allst = collections.Counter()
for b in by_op.values():
    allst.update(b[2])
print("stall totals:", ", ".join(f"{k[6:]} {100*v/tot_s:.1f}%" for k, v in allst.most_common(10)))
print(f"\nhottest {top} instructions")
for smp, ex, sass, st in sorted(insts, key=lambda t: -t[0])[:top]:
    tops = ", ".join(f"{k[6:]} {v}" for k, v in collections.Counter(st).most_common(2))
    print(f"{smp:7d} smp {ex/units:8.2f}/unit  {sass[:70]:70s} {tops}")
