"""Config 2 frame with HOST detections every frame (mot_tdd_step): launch-level timing under ncu, or wall-clock per step without."""
import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'multiple-object-tracking_b200')
import mot_b200 as M
from synth import Scene
W, H = 1920, 1080
sc = Scene(0x5EED0200, W, H, 64, tsize=48, win=96)
ctx = M.Context(W, H, max_tracks=128, kind=M.TRACKER_KALMAN)
loop = M.DeviceLoop(ctx, 1, cap=128, max_det=64, cost_mode=0)
dets = []
for f in range(200):
    sc.step(); dets.append(sc.windows(jitter=2))
for f in range(20): loop.step([dets[f]])
ctx.sync(); t0 = time.perf_counter()
for f in range(20, 200): loop.step([dets[f]])
ctx.sync(); print("us per step:", (time.perf_counter() - t0) / 180 * 1e6)
