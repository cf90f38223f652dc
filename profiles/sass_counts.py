#!/usr/bin/env python
"""stdin: `cuobjdump -sass libmot_b200.so`; per kernel family, how many of the instructions that prove the copy engines are used:
UTMALDG (2-D TMA tensor load), UBLKCP (1-D bulk copy, the same engine), UBLKPF (bulk L2 prefetch), LDGSTS (cp.async), SYNCS (mbarrier)."""
import collections, re, sys
fam = None
cnt = collections.defaultdict(collections.Counter)
for l in sys.stdin:
    m = re.search(r"Function : (\S+)", l)
    if m:
        n = m.group(1)
        fam = "kcf_fused_kernel" if "kcf_fused" in n else "kcf_any_kernel" if "kcf_any" in n else "munkres_kernel" if "munkres" in n else "other"
        cnt[fam]["kernels"] += 1
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m and fam:
        op = m.group(1).split(".")[0]
        if op in ("UTMALDG", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "UTMAPF", "BAR", "HMMA", "UTCHMMA"):
            cnt[fam][op] += 1
        cnt[fam]["instructions"] += 1
for f, c in sorted(cnt.items()):
    print("%-18s %s" % (f, "  ".join("%s %d" % kv for kv in sorted(c.items()))))
print("(no tcgen05 / HMMA instructions anywhere: the path has no dense contraction)")
