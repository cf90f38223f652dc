#!/usr/bin/env python
"""KCF track-updates/s per window size: host-array calls with 256 tracks (as in round 1) and device-array calls with 4096 tracks
timed with CUDA events (the steady-state form).  `launches` per predict+update pair shows which path served the size:
2 = a fused kernel (fixed-size or any-size), ~20 = the unfused pipeline (csrc/kcf_generic.cu)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mot_b200 as M
from synth import boxes_array

W, H = 1920, 1080
rng = np.random.default_rng(3)
frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
sizes = [(128, 128), (64, 64), (100, 60), (120, 160), (52, 36), (148, 88), (200, 88), (160, 120), (300, 148)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
out = []
for rows, cols in sizes:
    rec = {"window_px": [rows, cols], "cells": [rows // 4, cols // 4]}
    for N, mode in ((256, "host"), (4096, "dev")):
        ctx = M.Context(W, H, max_tracks=N, n_frame_slots=1, kind=M.TRACKER_KCF)
        ctx.upload(0, frame)
        b = boxes_array(N)
        b["l"] = rng.integers(0, W - cols - 8, N); b["t"] = rng.integers(0, H - rows - 8, N)
        b["r"] = b["l"] + cols - 1; b["b"] = b["t"] + rows - 1
        h = ctx.new(b)
        fs = np.zeros(N, np.int32)
        ctx.update(h, fs, b)
        if mode == "host":
            for _ in range(2):
                ctx.predict(h, fs, b); ctx.update(h, fs, b)
            l0 = ctx.launches()
            t0 = time.perf_counter()
            K = 5
            for _ in range(K):
                ctx.predict(h, fs, b); ctx.update(h, fs, b)
            dt = (time.perf_counter() - t0) / K
            rec["host256_track_updates_per_s"] = N / dt
            rec["launches_per_pair"] = (ctx.launches() - l0) / K - 2      # minus the two staging kernels
        else:
            st = torch.cuda.Stream()
            ctx.set_stream(st.cuda_stream)
            d_h = torch.from_numpy(h.astype(np.int32)).cuda(); d_f = torch.zeros(N, dtype=torch.int32, device="cuda")
            d_b = torch.from_numpy(b.view(np.uint8).reshape(N, 24).copy()).cuda()
            with torch.cuda.stream(st):
                for _ in range(2):
                    ctx.predict_dev(N, d_h.data_ptr(), d_f.data_ptr(), d_b.data_ptr(), 1); ctx.update_dev(N, d_h.data_ptr(), d_f.data_ptr(), d_b.data_ptr())
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                K = 5
                tp = tu = 0.0
                for _ in range(K):
                    e[0].record(st); ctx.predict_dev(N, d_h.data_ptr(), d_f.data_ptr(), d_b.data_ptr(), 1)
                    e[1].record(st); ctx.update_dev(N, d_h.data_ptr(), d_f.data_ptr(), d_b.data_ptr())
                    e[2].record(st); st.synchronize()
                    tp += e[0].elapsed_time(e[1]); tu += e[1].elapsed_time(e[2])
            rec["dev4096_ms_predict"] = tp / K; rec["dev4096_ms_update"] = tu / K
            rec["dev4096_track_updates_per_s"] = N / ((tp + tu) / K * 1e-3)
            ctx.set_stream(None)
        ctx.close()
    out.append(rec)
    print(json.dumps(rec), flush=True)
