#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with nvdisasm line info and aggregate samples per CUDA source line.

usage: sass_lines.py <ncu_source.csv> <kernel.cubin.txt from `nvdisasm -g -c`> <mangled-name-substring> <source.cu> [top_n]
"""
import collections
import csv
import re
import sys


def parse_disasm(path, fn_pat):
    insts, infn, cur = [], False, None
    for l in open(path):
        m = re.match(r'\s*\.text\.(\S+):', l)
        if m:
            infn = fn_pat in m.group(1); continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            insts.append((int(m.group(1), 16), cur, m.group(2)))
    return insts


def main():
    ncu_csv, disasm, pat, src_path = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    insts = parse_disasm(disasm, pat)
    rows = list(csv.reader(open(ncu_csv)))
    sects, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = []; sects.append(cur); continue
        if r and r[0] == 'Address':
            hdr = r; continue
        if cur is not None and hdr and len(r) == len(hdr):
            cur.append(dict(zip(hdr, r)))
    sect = next(s for s in sects if len(s) == len(insts))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for (off, line, ins), d in zip(insts, sect):
        a = agg[line]
        a[0] += int(d['# Samples'] or 0); a[1] += int(d['Instructions Executed'] or 0); a[2] += 1
    tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
    print('SASS instructions %d (%d KB); samples %d; warp-instructions executed %d' % (len(insts), len(insts) * 16 // 1024, tot, toti))
    src = open(src_path).read().split('\n')
    for line, (s, i, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[line[1] - 1].strip()[:100] if line and line[0] == src_path.split('/')[-1] else ''
        print('  %-20s smp %5d (%4.1f%%) inst %9d (%4.1f%%) sass %5d | %s' % ('%s:%d' % line if line else '?', s, 100 * s / tot, i, 100 * i / toti, n, text))


if __name__ == '__main__':
    main()
