#!/usr/bin/env python
"""Minimal workload for ncu captures of the any-size kernel: N tracks of one window size, a few predict + update rounds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot_b200 as M
from synth import boxes_array
rows, cols = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "120x160").split("x"))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
W, H = 1920, 1080
rng = np.random.default_rng(3)
frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
ctx = M.Context(W, H, max_tracks=N, n_frame_slots=1, kind=M.TRACKER_KCF)
ctx.upload(0, frame)
b = boxes_array(N)
b["l"] = rng.integers(0, W - cols - 8, N); b["t"] = rng.integers(0, H - rows - 8, N)
b["r"] = b["l"] + cols - 1; b["b"] = b["t"] + rows - 1
h = ctx.new(b); fs = np.zeros(N, np.int32)
ctx.update(h, fs, b)
for _ in range(3):
    ctx.predict(h, fs, b); ctx.update(h, fs, b)
ctx.close()
