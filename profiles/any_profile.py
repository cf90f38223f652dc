#!/usr/bin/env python
"""Per-phase and per-line breakdown of an `ncu --set full --import-source on` capture of the any-size kernel (csrc/kcf_any_kernel.cuh).

usage: python profiles/any_profile.py REPORT.ncu-rep MODE(0=predict,1=update) [TOP_LINES [NTMAX(512|1024) [STRIPS(0|1)]]]      (same build as the report)
"""
import collections, csv, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import parse_disasm  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multiple-object-tracking_b200")
SRC = os.path.join(PKG, "csrc", "kcf_any_kernel.cuh")
STALLS = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_not_selected", "stall_selected",
          "stall_lg", "stall_dispatch", "stall_no_inst", "stall_branch_resolving", "stall_membar", "stall_tex", "stall_sleeping"]


def main():
    rep, mode = sys.argv[1], int(sys.argv[2])
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(PKG, "build", "kcf_any_inst_%d_0.o" % mode)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = os.path.join(tmp, "dis.txt")
    open(dis, "w").write(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout)
    src = open(SRC).read().split("\n")
    marks = [(i + 1, re.search(r"-{20,} (P[\w./ ]*\d)", l).group(1).strip()) for i, l in enumerate(src) if re.search(r"// -{20,} P\d", l)]
    sub = [(i + 1, l.strip()[8:60]) for i, l in enumerate(src) if re.match(r"\s*// ---- ", l)]
    sects, cur, shdr = [], None, None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = []; sects.append((r[1], cur)); continue
        if r and r[0] == "Address":
            shdr = r; continue
        if cur is not None and shdr and len(r) == len(shdr):
            cur.append(dict(zip(shdr, r)))
    ntmax = int(sys.argv[4]) if len(sys.argv) > 4 else 512
    strips = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    insts = parse_disasm(dis, "kcf_any_kernelILi%dELb0ELi%dELb%dELb0E" % (mode, ntmax, strips))
    sect = next(s for k, s in sects if len(s) == len(insts))

    def phase_of(line):
        if not line:
            return "?"
        if line[0] != "kcf_any_kernel.cuh":
            return line[0]
        key = "pre"
        for ln, nm in marks:
            if line[1] >= ln:
                key = nm
        if line[1] < marks[0][0]:
            # helper functions above the kernel
            for name, (a, b) in HELPERS.items():
                if a <= line[1] <= b:
                    return name
            return "pre"
        for ln, nm in sub:
            if line[1] >= ln and ln > [m for m in marks if m[1] == key][0][0]:
                key2 = nm
        return key

    agg = collections.defaultdict(collections.Counter)
    lines = collections.defaultdict(collections.Counter)
    for (off, line, ins), d in zip(insts, sect):
        ph = phase_of(line)
        for s in STALLS:
            v = int(d.get(s) or 0)
            agg[ph][s] += v; lines[line][s] += v
        for k, col in (("n", "# Samples"), ("inst", "Instructions Executed"), ("wf", "L1 Wavefronts Shared"), ("wfx", "L1 Wavefronts Shared Excessive")):
            v = int(d.get(col) or 0)
            agg[ph][k] += v; lines[line][k] += v
    tot = sum(a["n"] for a in agg.values()) or 1
    toti = sum(a["inst"] for a in agg.values()) or 1
    print("SASS instructions %d; samples %d; warp-instructions executed %d" % (len(insts), tot, toti))
    for ph, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"]):
        top = ", ".join("%s %.0f%%" % (s.replace("stall_", ""), 100.0 * a[s] / max(a["n"], 1)) for s in sorted(STALLS, key=lambda s: -a[s])[:5])
        print("%-22s %5.1f%% smp %5.1f%% inst %9d  smem wf %9d (excess %8d) | %s" % (ph[:22], 100.0 * a["n"] / tot, 100.0 * a["inst"] / toti, a["inst"], a["wf"], a["wfx"], top))
    print("-" * 40, "hottest source lines")
    for line, a in sorted(lines.items(), key=lambda kv: -kv[1]["n"])[:topn]:
        text = src[line[1] - 1].strip()[:90] if line and line[0] == "kcf_any_kernel.cuh" else ""
        st = ", ".join("%s %.0f%%" % (s.replace("stall_", ""), 100.0 * a[s] / max(a["n"], 1)) for s in sorted(STALLS, key=lambda s: -a[s])[:2])
        print("  %-22s %4.1f%% smp %4.1f%% inst wf %8d/%7d  %-30s| %s" % ("%s:%d" % line if line else "?", 100.0 * a["n"] / tot, 100.0 * a["inst"] / toti, a["wf"], a["wfx"], st, text))


def helper_ranges():
    src = open(SRC).read().split("\n")
    out = {}
    names = [("fft_pass_prime", "fft_pass_prime"), ("fft_pass", "fft_pass("), ("fft_batch", "fft_batch("), ("gather_cells", "gather_cells("), ("Bfly", "struct Bfly")]
    for i, l in enumerate(src):
        for key, pat in names:
            if pat in l and ("__device__" in l or "struct" in l) and key not in out:
                out[key] = i + 1
    starts = sorted(out.items(), key=lambda kv: kv[1])
    res = {}
    for q, (k, a) in enumerate(starts):
        b = starts[q + 1][1] - 1 if q + 1 < len(starts) else a + 200
        res[k] = (a, b)
    return res


HELPERS = helper_ranges()

if __name__ == "__main__":
    main()
