"""Launch-level timing of one device-resident frame (config 2: 64 tracks x 64 detections, Kalman + Hungarian)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'multiple-object-tracking_b200')
import mot_b200 as M
from synth import Scene
W, H = 1920, 1080
sc = Scene(0x5EED0200, W, H, 64, tsize=48, win=96)
ctx = M.Context(W, H, max_tracks=128, kind=M.TRACKER_KALMAN)
loop = M.DeviceLoop(ctx, 1, cap=128, max_det=64, cost_mode=0)
dev = torch.device('cuda', 0)
nd = torch.full((1,), 64, dtype=torch.int32, device=dev)
for f in range(30):
    sc.step(); d = sc.windows(jitter=2)
    buf = np.zeros((1, 64), M.BBOX_DTYPE); buf[0, :len(d)] = d
    t = torch.from_numpy(buf.view(np.uint8).reshape(1, 64 * 24)).to(dev)
    loop.step_dev(t.data_ptr(), nd.data_ptr())
ctx.sync()
