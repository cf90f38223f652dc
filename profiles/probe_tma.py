#!/usr/bin/env python
"""Counts the TMA tensor loads the fused predict kernel executes (ncu metric) for a small 128x128-px workload."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot_b200 as M
from synth import boxes_array
N, W, H = 296, 1920, 1080
rng = np.random.default_rng(3)
frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
ctx = M.Context(W, H, max_tracks=N, n_frame_slots=1, kind=M.TRACKER_KCF)
ctx.upload(0, frame)
b = boxes_array(N)
b["l"] = rng.integers(0, W - 136, N); b["t"] = rng.integers(0, H - 136, N); b["r"] = b["l"] + 127; b["b"] = b["t"] + 127
h = ctx.new(b); fs = np.zeros(N, np.int32)
ctx.update(h, fs, b); ctx.predict(h, fs, b); ctx.update(h, fs, b)
ctx.close()
