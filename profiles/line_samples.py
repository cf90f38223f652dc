#!/usr/bin/env python
"""Stall samples per SOURCE LINE of one kernel: python profiles/line_samples.py REPORT.ncu-rep OBJECT.o MANGLED_SUBSTRING [TOPN]
(the object must be the build the report was taken from: SASS instructions are matched by position)."""
import collections, csv, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import parse_disasm  # noqa: E402

rep, obj, sym = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = os.path.join(tmp, "dis.txt")
open(dis, "w").write(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout)
insts = parse_disasm(dis, sym)
rows, hdr = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        rows.append(dict(zip(hdr, r)))
assert len(rows) == len(insts), (len(rows), len(insts))
agg = collections.Counter(); ninst = collections.Counter()
tot = 0
for (off, line, ins), d in zip(insts, rows):
    n = int(d["# Samples"] or 0); tot += n
    agg[line] += n; ninst[line] += int(d["Instructions Executed"] or 0)
print("total samples", tot)
for line, n in agg.most_common(topn):
    print("%6.2f%%  inst %9d  %s" % (100.0 * n / max(tot, 1), ninst[line], line))
