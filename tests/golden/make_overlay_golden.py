#!/usr/bin/env python
"""Generate tests/golden/overlay_v1.npz from the reference (run in the authoring container, where /root/reference exists).

  colormap[256]  the table of top/td.cpp:652-697, parsed from the source text
  tids[N], hashes[N]  hashcolor (top/td.cpp:295-305) evaluated by compiling the function's own text, extracted from the
                 source where it lies into a scratch file under /tmp (nothing of the reference is copied into the repo)
"""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np

REF = "/root/reference/top/td.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(REF).read()
    blk = src[src.index("static uint32_t colormap[]"):]
    blk = blk[:blk.index("};")]
    cmap = np.array([int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{6})", blk)], np.uint32)
    assert len(cmap) == 256
    m = re.search(r"static uint32_t hashcolor\s*\(\s*uint32_t a\)\s*\{.*?\n\}", src, re.S)
    tmp = tempfile.mkdtemp()
    cfile = os.path.join(tmp, "h.c")
    open(cfile, "w").write("#include <stdint.h>\n" + m.group(0).replace("static ", "") + "\n")
    so = os.path.join(tmp, "h.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", cfile, "-o", so])
    lib = ctypes.CDLL(so)
    lib.hashcolor.restype = ctypes.c_uint32
    lib.hashcolor.argtypes = [ctypes.c_uint32]
    rng = np.random.default_rng(7)
    tids = np.concatenate([np.arange(0, 2048, dtype=np.uint32), rng.integers(0, 2 ** 32, 2048, dtype=np.uint64).astype(np.uint32)])
    hashes = np.array([lib.hashcolor(int(t)) for t in tids], np.uint32)
    # The spawn statement itself (top/td.cpp:619-620: `tid = tracker_id++; color = hashcolor(tracker_id) & 255`), compiled from
    # its own two lines around a stand-in tracker_info array: pins WHICH counter value the colour hashes (tid + 1).
    lines = src.split("\n")[618:620]
    assert "tracker_id++" in lines[0] and "hashcolor(tracker_id)" in lines[1], lines
    cfile2 = os.path.join(tmp, "s.c")
    open(cfile2, "w").write("#include <stdint.h>\n" + m.group(0) + "\ntypedef struct { uint32_t tid, color; } ti_t;\n"
                            "void spawn_n(uint32_t first, int n, uint32_t *tid, uint32_t *color)\n{\n    uint32_t tracker_id = first; ti_t tracker_info[1];\n"
                            "    for (int q = 0; q < n; ++q) { const int i = 0;\n" + "\n".join(lines) +
                            "\n        tid[q] = tracker_info[i].tid; color[q] = tracker_info[i].color; }\n}\n")
    so2 = os.path.join(tmp, "s.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", cfile2, "-o", so2])
    lib2 = ctypes.CDLL(so2)
    n_sp = 4096
    sp_tid = np.zeros(2 * n_sp, np.uint32); sp_col = np.zeros(2 * n_sp, np.uint32)
    for half, first in enumerate((0, 2 ** 32 - 100)):              # from the initial counter, and across the 32-bit wrap
        lib2.spawn_n(ctypes.c_uint32(first), n_sp, sp_tid[half * n_sp:].ctypes.data_as(ctypes.c_void_p), sp_col[half * n_sp:].ctypes.data_as(ctypes.c_void_p))
    np.savez_compressed(os.path.join(HERE, "overlay_v1.npz"), colormap=cmap, tids=tids, hashes=hashes, spawn_tid=sp_tid, spawn_color=sp_col)
    print("wrote overlay_v1.npz:", len(cmap), "colours,", len(tids), "hashes,", len(sp_tid), "spawn (tid, colour index) pairs")


if __name__ == "__main__":
    main()
