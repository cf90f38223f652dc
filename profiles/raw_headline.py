#!/usr/bin/env python
"""stdin: `ncu --page raw --csv`; prints the headline metrics and the stall-reason split of every kernel in the report."""
import csv, sys
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(sys.stdin))
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
