#!/usr/bin/env python
"""Launch list of the device-resident KCF frame loop at config 3 (256 tracks, one 1080p stream): run under
ncu --metrics gpu__time_duration.sum to see which kernel a frame's latency goes to."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot_b200 as M
from synth import Scene
W, H = 1920, 1080
sc = Scene(0x5EED0300, W, H, 256, tsize=56, win=128)
ctx = M.Context(W, H, max_tracks=256, n_frame_slots=2, kind=M.TRACKER_KCF)
loop = M.DeviceLoop(ctx, 1, cap=256, max_det=256, cost_mode=0)
loop.kcf_windows([(128, 128)])
for f in range(4):
    sc.step(); ctx.upload(f & 1, sc.render()); loop.frame_base(f & 1); loop.step([sc.windows(jitter=2)])
ctx.sync()
print("tracks:", len(loop.tracks(0)["tid"]))
