"""Join an ncu SASS-level source page with nvdisasm line info: executed instructions / stall samples per source line.

usage: python tools/ncu_by_line.py <report.ncu-rep> <object.o> <mangled kernel name> <problem-steps>
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kern, steps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
sass = [(r[isrc], int(r[ia] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > ia]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sec = dis.split(".text." + kern + ":")[1].split("\n.section")[0] if (".text." + kern + ":") in dis else dis.split(kern)[2]
lines, cur = [], None
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        outer = inl[-1] if inl else (m.group(1), m.group(2))
        cur = (os.path.basename(outer[0]), int(outer[1]), os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
by_outer = collections.Counter()
by_outer_s = collections.Counter()
for (text, n, s), loc in zip(sass, lines):
    key = (loc[0], loc[1]) if loc else ("?", 0)
    by_outer[key] += n
    by_outer_s[key] += s
tot = sum(by_outer.values())
print(f"total {tot/steps:.1f} instr per problem-step")
srcs = {}
for (f, l), n in sorted(by_outer.items(), key=lambda kv: -kv[1])[:int(__import__("os").environ.get("TOPN","45"))]:
    if f not in srcs:
        for root in ("optimistix_b200/csrc", "."):
            pth = os.path.join(root, f)
            if os.path.exists(pth):
                srcs[f] = open(pth).read().splitlines()
                break
        else:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{n/steps:8.1f} /step {100*n/tot:5.1f}%  samples {by_outer_s[(f,l)]:6d}  {f}:{l}  {text}")
