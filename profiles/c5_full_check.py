#!/usr/bin/env python
"""BASELINE config 5 at full size, EVERY matrix checked: 1024 association problems of 512 x 512 (cost matrix + Munkres on the GPU), each
assignment vector compared bit for bit with the reference's assignmentoptimal run on all host cores (compiled reference when
oracle/_ref is present).  One-off evidence run (a few minutes of CPU); the CI tests check samples and properties."""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multiple-object-tracking_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import mot_b200 as M
import oraclelib
from synth import random_boxes

n, dim = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 512
out = {"config": "C5: %d x (%dx%d), every matrix compared with assignmentoptimal" % (n, dim, dim)}
for mode, name in ((M.COST_REF_CENTROID, "ref_centroid"), (M.COST_IOU_CLAMPED, "iou_clamped")):
    rng = np.random.default_rng(0x5EED0500 + mode)
    trk = np.zeros((n, dim), M.BBOX_DTYPE); det = np.zeros((n, dim), M.BBOX_DTYPE)
    for m in range(n):
        trk[m] = random_boxes(rng, dim, 1920, 1080); det[m] = random_boxes(rng, dim, 1920, 1080)
    ctx = M.Context(1920, 1080, max_tracks=4, kind=M.TRACKER_KALMAN)
    dev = torch.device("cuda", 0)
    d_T = torch.full((n,), dim, dtype=torch.int32, device=dev)
    d_trk = torch.from_numpy(trk.view(np.uint8).reshape(n, dim * 24)).to(dev); d_det = torch.from_numpy(det.view(np.uint8).reshape(n, dim * 24)).to(dev)
    d_dist = torch.zeros((n, dim * dim), dtype=torch.float64, device=dev)
    d_assign = torch.zeros((n, dim), dtype=torch.int32, device=dev); d_cost = torch.zeros(n, dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    rc = M.lib().mot_associate_batch_dev(ctx.h, n, d_T.data_ptr(), d_T.data_ptr(), d_trk.data_ptr(), dim, d_det.data_ptr(), dim, mode,
                                         d_dist.data_ptr(), dim * dim, d_assign.data_ptr(), dim, d_cost.data_ptr(), dim)
    assert rc == 0, M.lib().mot_last_error()
    ctx.sync()
    gpu_s = time.perf_counter() - t0
    assign = d_assign.cpu().numpy(); dist = d_dist.cpu().numpy(); cost = d_cost.cpu().numpy()
    orc = oraclelib.Oracle(oraclelib.best())

    def check(m):
        a, c = orc.assign(dist[m].reshape(dim, dim).T)
        return bool(np.array_equal(a, assign[m])) and c == cost[m]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        ok = list(ex.map(check, range(n)))
    cpu_s = time.perf_counter() - t0
    out[name] = {"matrices": n, "bit_exact": int(sum(ok)), "mismatches": int(n - sum(ok)), "gpu_s_incl_first_launch": gpu_s, "cpu_s_all_cores": cpu_s,
                 "host_cores": os.cpu_count(), "oracle": orc.kind}
    ctx.close()
print(json.dumps(out))
