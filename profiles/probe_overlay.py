#!/usr/bin/env python
"""Overlay step (top/td.cpp:647-733): device kernel vs the CPU restatement, 64 1080p frames x 128 tracks (run on the GPU box)."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot_b200 as M
import oraclelib
from synth import boxes_array

W, H, NS, NT = 1920, 1080, 64, 128
rng = np.random.default_rng(1)
ctx = M.Context(W, H, max_tracks=8, n_frame_slots=NS, kind=M.TRACKER_KALMAN)
frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
for s in range(NS):
    ctx.upload(s, frame)
n = NS * NT
b = boxes_array(n)
b["l"] = rng.integers(0, W - 140, n); b["t"] = rng.integers(0, H - 140, n); b["r"] = b["l"] + 127; b["b"] = b["t"] + 127
slots = np.repeat(np.arange(NS, dtype=np.int32), NT)
rgb = rng.integers(0, 1 << 24, n, dtype=np.uint32)
for _ in range(3):
    ctx.overlay(slots, b, rgb, 3)
t0 = time.perf_counter()
K = 20
for _ in range(K):
    ctx.overlay(slots, b, rgb, 3)
gpu_ms = (time.perf_counter() - t0) / K * 1e3
L = oraclelib.Oracle("port").kcf
f2 = frame.copy()
t0 = time.perf_counter()
for s in range(NS):
    L.port_overlay(f2.ctypes.data_as(C.c_void_p), C.c_int(f2.strides[0]), C.c_long(f2.nbytes), C.c_int(NT),
                   b[s * NT:(s + 1) * NT].ctypes.data_as(C.c_void_p), rgb[s * NT:(s + 1) * NT].ctypes.data_as(C.c_void_p), C.c_int(3))
cpu_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"what": "overlay, 64 x 1080p frames x 128 tracks x 3 rectangles, host-array call incl. upload of the entry arrays and sync",
                  "gpu_ms_per_call": gpu_ms, "cpu_port_ms_one_core": cpu_ms, "entries": n}))
