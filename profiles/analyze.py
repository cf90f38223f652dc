#!/usr/bin/env python
"""Summarise an ncu report of the fused KCF kernels: headline metrics, stall reasons, and stall samples / executed
instructions aggregated per kernel phase (P0..P7 markers in csrc/kcf_fused.cuh) via nvdisasm line info.

usage: python profiles/analyze.py gpurun_out/prof.ncu-rep [HR WC]   (run from the repo root, same build as the report)
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_lines import parse_disasm  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "multiple-object-tracking_b200")
SRC = os.path.join(PKG, "csrc", "kcf_fused.cuh")

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    hr, wc = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32, 32)
    tmp = tempfile.mkdtemp()
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")])
        for m in METRICS:
            if m in hdr:
                print("  %-62s %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), int(float(r[i] or 0))) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
        tot = sum(v for _, v in st) or 1
        print("  stall samples: " + ", ".join("%s %.0f%%" % (k, 100.0 * v / tot) for k, v in sorted(st, key=lambda kv: -kv[1])[:9]))
    # ---- per-phase aggregation ---------------------------------------------------------------------------------
    src_csv = os.path.join(tmp, "src.csv")
    open(src_csv, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
    obj = os.path.join(PKG, "build", "kcf_inst_%d_%d.o" % (hr, wc))
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = os.path.join(tmp, "dis.txt")
    open(dis, "w").write(subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout)
    src = open(SRC).read().split("\n")
    marks = [(i + 1, re.search(r"-- (P\d\w?)[: ]", l).group(1)) for i, l in enumerate(src) if re.search(r"// -{20,} P\d", l)]
    rows = list(csv.reader(open(src_csv)))
    sects, cur, shdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []; sects.append((r[1], cur)); continue
        if r and r[0] == "Address":
            shdr = r; continue
        if cur is not None and shdr and len(r) == len(shdr):
            cur.append(dict(zip(shdr, r)))
    for mode, name in ((0, "predict"), (1, "update")):
        insts = parse_disasm(dis, "kcf_fused_kernelILi%dELi%dELi%dELb0E" % (hr, wc, mode))
        sect = next((s for k, s in sects if len(s) == len(insts) and ("(int)%d, (bool)0" % mode) in k), None)
        if not sect:
            continue
        agg = collections.OrderedDict((m[1], [0, 0]) for m in marks)
        agg["fft_reg"] = [0, 0]; agg["pre/other"] = [0, 0]
        for (off, line, ins), d in zip(insts, sect):
            key = "pre/other"
            if line and line[0] == "fft_reg.cuh":
                key = "fft_reg"
            elif line and line[0] == "fhog_common.cuh":
                key = "P0" if line[1] < 22 else "P1"                     # bgr_gray / grad_pixel_k
            elif line and line[0] == "kcf_fused.cuh":
                for ln, nm in marks:
                    if line[1] >= ln:
                        key = nm
                if line[1] < marks[0][0]:
                    key = "pre/other" if line[1] > 75 else "P0"      # helper functions (gray, clamp) belong to P0
            agg[key][0] += int(d["# Samples"] or 0); agg[key][1] += int(d["Instructions Executed"] or 0)
        ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
        grid = 1
        print("-" * 100)
        print("%s kernel: %d SASS instructions (%d KB), %d samples, %d warp-instructions" % (name, len(insts), len(insts) * 16 // 1024, ts, ti))
        for k, (s, i) in agg.items():
            print("  %-10s samples %5.1f%%   instructions %5.1f%%" % (k, 100.0 * s / ts, 100.0 * i / ti))


if __name__ == "__main__":
    main()
