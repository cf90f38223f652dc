#!/usr/bin/env python
"""Any-size KCF path (csrc/kcf_generic.cu) vs the fused kernels: predict + update time per track for a few window sizes."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot_b200 as M
from synth import boxes_array

W, H, N = 1920, 1080, 256
rng = np.random.default_rng(3)
frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
out = []
for rows, cols in [(128, 128), (64, 64), (100, 60), (120, 160), (200, 90), (300, 150)]:
    ctx = M.Context(W, H, max_tracks=N, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    b = boxes_array(N)
    b["l"] = rng.integers(0, W - cols - 8, N); b["t"] = rng.integers(0, H - rows - 8, N)
    b["r"] = b["l"] + cols - 1; b["b"] = b["t"] + rows - 1
    h = ctx.new(b)
    fs = np.zeros(N, np.int32)
    ctx.update(h, fs, b)
    for _ in range(2):
        ctx.predict(h, fs, b); ctx.update(h, fs, b)
    t0 = time.perf_counter()
    K = 5
    for _ in range(K):
        ctx.predict(h, fs, b); ctx.update(h, fs, b)
    dt = (time.perf_counter() - t0) / K
    out.append({"window_px": [rows, cols], "cells": [rows // 4, cols // 4], "tracks": N, "ms_predict_plus_update": dt * 1e3,
                "track_updates_per_s": N / dt})
    ctx.close()
print(json.dumps(out))
