"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference compiled into oracle/_ref (needs /root/reference).

The reference holds no golden vectors of its own, so these are outputs of the reference itself,
run in the authoring container on seeded inputs, committed as small fixtures together with this script.  FFT inside
the KCF cases is the double-precision DFT shim (FFTW is an absent third-party dependency: parity there is unpinned).
Inputs are regenerated from the seeds by golden_inputs(); only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from synth import BBox, Scene, random_boxes, jittered_detections, boxes_array  # noqa: E402


def textured(rng, h, w):
    base = rng.random((h // 8 + 2, w // 8 + 2)).repeat(8, 0).repeat(8, 1)[:h, :w] * 180
    return (base + rng.random((h, w)) * 20).astype(np.float32)


def golden_inputs():
    """Everything the golden cases are computed from (pure function of fixed seeds)."""
    g = {}
    g["fhog_patches"] = {(h, w): textured(np.random.default_rng(1000 * h + w), h, w) for (h, w) in [(64, 64), (37, 53), (128, 128)]}
    rng = np.random.default_rng(424242)
    g["frame"] = rng.integers(0, 256, size=(240, 320, 3), dtype=np.uint8)
    g["crops"] = [((30, 40, 129, 109), (64, 64)), ((10, 10, 73, 73), (64, 64)), ((100, 50, 260, 199), (128, 64))]   # (l,t,r,b) -> (rows_d, cols_d)
    g["kcf_img"] = textured(np.random.default_rng(7), 240, 320)
    g["kcf_box"] = (80, 60, 64, 64)              # l, t, rows, cols
    g["kcf_shifts"] = [(4, -8), (-4, 4), (8, 8), (0, -4), (-8, 0)]
    mats = []
    rng = np.random.default_rng(99)
    for (T, D, mode) in [(16, 16, 0), (16, 16, 1), (12, 20, 1), (20, 12, 0), (40, 40, 1), (33, 48, 0)]:
        trk = random_boxes(rng, T, 1920, 1080)
        det = jittered_detections(rng, trk, 1920, 1080)[:D] if D <= T else np.concatenate([jittered_detections(rng, trk, 1920, 1080), random_boxes(rng, D - T, 1920, 1080)])
        mats.append((trk, np.ascontiguousarray(det), mode))
    g["assoc"] = mats
    g["ties"] = [np.zeros((6, 6)), np.ones((5, 9)), np.ones((9, 5)), np.random.default_rng(5).integers(0, 3, size=(24, 24)).astype(float),
                 np.random.default_rng(6).integers(0, 4, size=(18, 27)).astype(float), np.random.default_rng(8).integers(0, 4, size=(27, 18)).astype(float)]
    return g


def kalman_trace(orc, n_frames=60):
    rng = np.random.default_rng(3)
    h = orc.kal_new(BBox(100, 50, 120, 180, 0, 1.0))
    pos = np.array([100.0, 50.0]); boxes = []; xs = []
    for i in range(n_frames):
        b = BBox(); orc.kal_predict(h, b); boxes.append(b.tup())
        pos += np.array([1.7, -0.6]) + rng.normal(0, 0.5, 2)
        l, t = int(pos[0]), int(pos[1])
        orc.kal_update(h, BBox(l, t, t + 70, l + 80, 0, 1.0))
        x, P, _ = orc.kal_state(h); xs.append(np.concatenate([x, P.ravel()]))
    orc.kal_delete(h)
    return np.array(boxes, np.int32), np.array(xs)


def kcf_trace(orc, g):
    l, t, rows, cols = g["kcf_box"]
    img = g["kcf_img"].copy()
    b = BBox(l, t, t + rows - 1, l + cols - 1, 1, 1.0)
    h = orc.kcf_new(b)
    patch = lambda im, bb: im[bb.t:bb.b + 1, bb.l:bb.r + 1]
    orc.kcf_update(h, patch(img, b), b)
    out = dict(feat0=orc.kcf_get(h, "xf_tm").copy(), alpha0=orc.kcf_get(h, "alpha").copy())
    boxes, peaks = [], []
    for (dy, dx) in g["kcf_shifts"]:
        img = np.roll(img, (dy, dx), (0, 1))
        orc.kcf_predict(h, patch(img, b), b)
        boxes.append(b.tup()); peaks.append(int(orc.kcf_get(h, "response").argmax()))
        orc.kcf_update(h, patch(img, b), b)
    out["boxes"] = np.array(boxes, np.int32); out["peaks"] = np.array(peaks, np.int32)
    out["alpha_end"] = orc.kcf_get(h, "alpha").copy()
    out["resp_end"] = orc.kcf_get(h, "response").copy()
    orc.kcf_delete(h)
    return out


def td_trace(orc, tracker, n_frames):
    W, H = 640, 480
    sc = Scene(101, W, H, 6 if tracker == "kcf" else 24, tsize=32, win=64)
    td = orc.td_new(tracker, W, H, 64, 0)
    drng = np.random.default_rng(9)
    rows = []
    for f in range(n_frames):
        sc.step(); frame = sc.render(); dets = sc.windows(jitter=2)
        dets = np.ascontiguousarray(dets[drng.random(len(dets)) > 0.1])
        td.step(frame, dets)
        t = td.tracks()
        rows.append(np.concatenate([[len(t["tid"])], t["tid"].astype(np.int64), t["boxes"]["l"], t["boxes"]["t"], t["boxes"]["b"], t["boxes"]["r"], t["age"], t["inv"]]).astype(np.int64))
    td.close()
    width = max(len(r) for r in rows)
    return np.array([np.pad(r, (0, width - len(r)), constant_values=-7) for r in rows])


def compute(orc, g):
    """All golden outputs from backend `orc` (the reference when generating; the port / CUDA path when checking)."""
    import ctypes as C
    out = {}
    for (h, w), I in g["fhog_patches"].items():
        out["fhog_%dx%d" % (h, w)] = orc.fhog(I)
    for i, ((l, t, r, b), (rd, cd)) in enumerate(g["crops"]):
        rs, cs = b - t + 1, r - l + 1
        crop = np.zeros(rs * cs, np.float32); res = np.zeros(rd * cd, np.float32)
        orc.kcf.port_rgb2gray(crop.ctypes.data_as(C.c_void_p), g["frame"].ctypes.data_as(C.c_void_p), g["frame"].strides[0], l, t, r, b)
        orc.kcf.port_resize_gray(res.ctypes.data_as(C.c_void_p), crop.ctypes.data_as(C.c_void_p), rs, cs, rd, cd)
        out["gray_%d" % i] = crop; out["resize_%d" % i] = res
    for k, v in kcf_trace(orc, g).items():
        out["kcf_" + k] = v
    kb, kx = kalman_trace(orc)
    out["kal_boxes"] = kb; out["kal_state"] = kx
    for i, (trk, det, mode) in enumerate(g["assoc"]):
        T, D = len(trk), len(det); nr, nc = (T, D) if T < D else (D, T)
        d = np.zeros(nr * nc, np.float64)
        orc.kcf.port_cost_matrix(d.ctypes.data_as(C.c_void_p), trk.ctypes.data_as(C.c_void_p), T, det.ctypes.data_as(C.c_void_p), D, mode, C.c_double(1.0 / 1920))
        dm = d.reshape(nc, nr).T
        a, c = orc.assign(dm)
        out["assoc_dist_%d" % i] = dm.copy(); out["assoc_assign_%d" % i] = a; out["assoc_cost_%d" % i] = np.array([c])
    for i, d in enumerate(g["ties"]):
        a, c = orc.assign(d)
        out["ties_assign_%d" % i] = a; out["ties_cost_%d" % i] = np.array([c])
    out["td_kal"] = td_trace(orc, "kal", 40)
    out["td_kcf"] = td_trace(orc, "kcf", 10)
    return out


if __name__ == "__main__":
    import oraclelib
    assert oraclelib.have_ref() or os.path.isdir("/root/reference"), "needs the compiled reference"
    if not oraclelib.have_ref():
        oraclelib.build_ref()
    orc = oraclelib.Oracle("ref")
    out = compute(orc, golden_inputs())
    # the cost matrices are produced by the restated cost loop (the original is not compilable stand-alone): keep only
    # what came out of reference code proper + the matrices they were computed on
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays; FFT provider:", orc.kcf.ref_fft_provider().decode())
