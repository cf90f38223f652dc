#!/usr/bin/env python
"""Timing of the detector post-processing (mot_yolo_post, host tensors in, bbox array out) next to the reference's own functions
(oracle/_ref/libref_yolo.so, else the C restatement) on one host core: images/s for a 416x416 network, 80 classes, 1280x720 frames.
Results are compared byte for byte on every image."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mot_b200 as M
import oraclelib
from test_oracle_yolo import synth_outputs, run, ANCHORS

th, tw, ih, iw, nc, nobj, obj, nms = 416, 416, 720, 1280, 80, 60, 0.5, 0.45
ref_so = os.path.join(oraclelib.ODIR, "_ref", "libref_yolo.so")
lib, fn = (C.CDLL(ref_so), "ref_yolo_post") if os.path.exists(ref_so) else (oraclelib.Oracle("port").kcf, "port_yolo_post")
rng = np.random.default_rng(1)
imgs = [synth_outputs(rng, th, tw, nc, nobj) for _ in range(32)]
ctx = M.Context(iw, ih, max_tracks=4, kind=M.TRACKER_KALMAN)
ctx.yolo_post(imgs[0], ANCHORS, obj, nms, th, tw, ih, iw, nc)
t0 = time.perf_counter(); got = [ctx.yolo_post(o, ANCHORS, obj, nms, th, tw, ih, iw, nc) for o in imgs]; tg = time.perf_counter() - t0
t0 = time.perf_counter(); want = [run(lib, fn, o, obj, nms, th, tw, ih, iw, nc) for o in imgs]; tc = time.perf_counter() - t0
same = all(a.tobytes() == b.tobytes() for a, b in zip(got, want))
bytes_in = sum(o.nbytes for o in imgs[0])
print(json.dumps({"what": "YOLO3 post-processing, 416x416 net, 80 classes, one image per call, host tensors", "gpu_images_per_s": len(imgs) / tg,
                  "cpu_images_per_s_one_core": len(imgs) / tc, "cpu_kind": fn, "identical": bool(same), "boxes_per_image": float(np.mean([len(g) for g in got])),
                  "input_bytes_per_image": bytes_in}))
ctx.close()
