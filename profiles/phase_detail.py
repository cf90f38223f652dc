#!/usr/bin/env python
"""Per-phase stall reasons and the hottest SASS instructions of one phase of the fused KCF kernel.

usage: python profiles/phase_detail.py REPORT.ncu-rep MODE(0=predict,1=update) [PHASE [TOPN]]   (same build as the report)
"""
import collections, csv, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import parse_disasm  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multiple-object-tracking_b200")
SRC = os.path.join(PKG, "csrc", "kcf_fused.cuh")
STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_not_selected", "stall_selected",
          "stall_lg", "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_membar", "stall_tex"]


def main():
    rep, mode = sys.argv[1], int(sys.argv[2])
    phase = sys.argv[3] if len(sys.argv) > 3 else None
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    tmp = tempfile.mkdtemp()
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(PKG, "build", "kcf_inst_32_32.o")], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = os.path.join(tmp, "dis.txt")
    open(dis, "w").write(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout)
    src = open(SRC).read().split("\n")
    marks = [(i + 1, re.search(r"-- (P\d\w?)[: ]", l).group(1)) for i, l in enumerate(src) if re.search(r"// -{20,} P\d", l)]
    sects, cur, shdr = [], None, None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = []; sects.append((r[1], cur)); continue
        if r and r[0] == "Address":
            shdr = r; continue
        if cur is not None and shdr and len(r) == len(shdr):
            cur.append(dict(zip(shdr, r)))
    insts = parse_disasm(dis, "kcf_fused_kernelILi32ELi32ELi%dELb0E" % mode)
    sect = next(s for k, s in sects if len(s) == len(insts) and ("(int)%d, (bool)0" % mode) in k)

    def phase_of(line):
        key = "pre/other"
        if line and line[0] == "fft_reg.cuh":
            return "fft_reg"
        if line and line[0] == "fhog_common.cuh":
            return "P0" if line[1] < 22 else "P1"
        if line and line[0] == "kcf_fused.cuh":
            for ln, nm in marks:
                if line[1] >= ln:
                    key = nm
            if line[1] < marks[0][0]:
                key = "pre/other" if line[1] > 75 else "P0"
        return key

    agg = collections.defaultdict(lambda: collections.Counter())
    rows = []
    for (off, line, ins), d in zip(insts, sect):
        ph = phase_of(line)
        for s in STALLS:
            agg[ph][s] += int(d.get(s) or 0)
        agg[ph]["n"] += int(d["# Samples"] or 0)
        agg[ph]["inst"] += int(d["Instructions Executed"] or 0)
        agg[ph]["wf"] += int(d["L1 Wavefronts Shared"] or 0)
        agg[ph]["wfx"] += int(d["L1 Wavefronts Shared Excessive"] or 0)
        rows.append((ph, off, line, ins, d))
    tot = sum(a["n"] for a in agg.values()) or 1
    for ph, a in agg.items():
        top = ", ".join("%s %.0f%%" % (s.replace("stall_", ""), 100.0 * a[s] / max(a["n"], 1)) for s in sorted(STALLS, key=lambda s: -a[s])[:5])
        print("%-9s %5.1f%% of samples, %9d inst, smem wavefronts %9d (excess %8d) | %s" % (ph, 100.0 * a["n"] / tot, a["inst"], a["wf"], a["wfx"], top))
    if phase:
        print("-" * 60, "hottest instructions of", phase)
        sel = sorted((r for r in rows if r[0] == phase), key=lambda r: -int(r[4]["# Samples"] or 0))[:topn]
        for ph, off, line, ins, d in sel:
            st = sorted(((int(d.get(s) or 0), s.replace("stall_", "")) for s in STALLS), reverse=True)[:2]
            print("%6s  %5d  L%-4s %-60s %s" % (off, int(d["# Samples"] or 0), line[1] if line else "?", ins[:60], " ".join("%s:%d" % (n, v) for v, n in st)))


if __name__ == "__main__":
    main()
