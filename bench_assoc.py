#!/usr/bin/env python
"""Secondary benchmarks of the association / Kalman side of the path (BASELINE configs 2 and 5).

  config 2: Kalman tracker + Hungarian association, 64 tracks x 64 detections per frame, one 1080p stream
            -> frames/s through the batched frame loop (mot_td_step) and through the oracle's loop on one host core
  config 5: N independent 512x512 association problems (cost matrix + Munkres), device-resident
            -> matrices/s on the GPU vs assignmentoptimal (oracle) on the host cores; assignments compared bit-exactly

    python bench_assoc.py [--matrices 256] [--dim 512] [--cpu-matrices 8]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "multiple-object-tracking_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np


def config5(args):
    import torch
    import mot_b200 as M
    import oraclelib
    from synth import random_boxes, jittered_detections
    n, dim = args.matrices, args.dim
    rng = np.random.default_rng(0x5EED0500)
    trk = np.zeros((n, dim), M.BBOX_DTYPE); det = np.zeros((n, dim), M.BBOX_DTYPE)
    for m in range(n):
        trk[m] = random_boxes(rng, dim, 1920, 1080)
        det[m] = random_boxes(rng, dim, 1920, 1080)          # independent boxes: the hard case, no greedy shortcut
    ctx = M.Context(1920, 1080, max_tracks=4, kind=M.TRACKER_KALMAN)
    dev = torch.device("cuda", 0)
    d_T = torch.full((n,), dim, dtype=torch.int32, device=dev); d_D = d_T.clone()
    d_trk = torch.from_numpy(trk.view(np.uint8).reshape(n, dim * 24)).to(dev); d_det = torch.from_numpy(det.view(np.uint8).reshape(n, dim * 24)).to(dev)
    d_dist = torch.zeros((n, dim * dim), dtype=torch.float64, device=dev)
    d_assign = torch.zeros((n, dim), dtype=torch.int32, device=dev); d_cost = torch.zeros(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
    out = {}
    for mode, name in ((M.COST_IOU_CLAMPED, "iou_clamped"), (M.COST_REF_CENTROID, "ref_centroid")):
        def run():
            rc = M.lib().mot_associate_batch_dev(ctx.h, n, d_T.data_ptr(), d_D.data_ptr(), d_trk.data_ptr(), dim, d_det.data_ptr(), dim, mode,
                                                 d_dist.data_ptr(), dim * dim, d_assign.data_ptr(), dim, d_cost.data_ptr(), dim)
            assert rc == 0, M.lib().mot_last_error()
        with torch.cuda.stream(stream):
            run(); stream.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream); run(); e1.record(stream); stream.synchronize()
        ms = e0.elapsed_time(e1)
        assign = d_assign.cpu().numpy(); dist = d_dist.cpu().numpy()
        orc = oraclelib.Oracle(oraclelib.best())
        ncpu = min(args.cpu_matrices, n)
        t0 = time.perf_counter()
        same = True
        for m in range(ncpu):
            a, _ = orc.assign(dist[m].reshape(dim, dim).T)
            same &= bool(np.array_equal(a, assign[m]))
        cpu_s = (time.perf_counter() - t0) / ncpu
        out[name] = {"gpu_matrices_per_s": n / (ms * 1e-3), "gpu_ms_total": ms, "cpu_s_per_matrix_one_core": cpu_s,
                     "speedup_vs_one_core": cpu_s * n / (ms * 1e-3), "bit_exact_on_checked": same, "checked": ncpu}
    ctx.close()
    return {"config": "C5: %d x (%dx%d) association problems, device resident" % (n, dim, dim), **out}


def config2(args):
    import mot_b200 as M
    import oraclelib
    from synth import Scene
    W, H, frames = 1920, 1080, 300
    sc = Scene(0x5EED0200, W, H, 64, tsize=48, win=96)
    dets = []
    for f in range(frames):
        sc.step(); dets.append(sc.windows(jitter=2))
    out = {}
    for mode, name in ((0, "ref_centroid"), (1, "iou_clamped")):
        ctx = M.Context(W, H, max_tracks=256, kind=M.TRACKER_KALMAN)
        td = ctx.td(0, cap=128, cost_mode=mode)
        td.step(None, dets[0])
        t0 = time.perf_counter()
        for f in range(1, frames):
            td.step(None, dets[f])
        g = (frames - 1) / (time.perf_counter() - t0)
        orc = oraclelib.Oracle(oraclelib.best())
        ref = orc.td_new("kal", W, H, 128, mode)
        ref.step(None, dets[0])
        t0 = time.perf_counter()
        for f in range(1, frames):
            ref.step(None, dets[f])
        c = (frames - 1) / (time.perf_counter() - t0)
        same = all(np.array_equal(td.tracks()[k], ref.tracks()[k]) for k in ("tid", "boxes", "age"))
        out[name] = {"gpu_frames_per_s_host_loop": g, "cpu_frames_per_s_one_core": c, "identical_final_track_table": bool(same)}
        td.close(); ref.close(); ctx.close()
    # the same loop over 64 independent streams stepped in lock step: one predict / associate / update launch per frame for all
    ns, fr2 = 64, 60
    scs = [Scene(0x5EED0200 + 7 * s_, W, H, 64, tsize=48, win=96) for s_ in range(ns)]
    dd = []
    for f in range(fr2):
        row = []
        for sc_ in scs:
            sc_.step(); row.append(sc_.windows(jitter=2))
        dd.append(row)
    ctx = M.Context(W, H, max_tracks=64 * 128, kind=M.TRACKER_KALMAN)
    tds = [ctx.td(0, cap=128, cost_mode=0) for _ in range(ns)]
    M.step_multi(tds, None, dd[0])
    t0 = time.perf_counter()
    for f in range(1, fr2):
        M.step_multi(tds, None, dd[f])
    out["lockstep_64_streams"] = {"gpu_stream_frames_per_s": ns * (fr2 - 1) / (time.perf_counter() - t0)}
    for t in tds:
        t.close()
    ctx.close()
    # device-resident loop (mot_tdd_*): detections resident too, one launch per frame, one sync at the very end
    import torch
    for ns2, key in ((1, "device_resident_1_stream"), (64, "device_resident_64_streams")):
        ctx = M.Context(W, H, max_tracks=ns2 * 128, kind=M.TRACKER_KALMAN)
        loop = M.DeviceLoop(ctx, ns2, cap=128, max_det=64, cost_mode=0)
        dev = torch.device("cuda", 0)
        dd_dev = []
        for f in range(fr2):
            buf = np.zeros((ns2, 64), M.BBOX_DTYPE)
            for s_ in range(ns2):
                buf[s_, :len(dd[f][s_])] = dd[f][s_]
            dd_dev.append(torch.from_numpy(buf.view(np.uint8).reshape(ns2, 64 * 24)).to(dev))
        nd_dev = torch.full((ns2,), 64, dtype=torch.int32, device=dev)
        loop.step_dev(dd_dev[0].data_ptr(), nd_dev.data_ptr()); ctx.sync()
        reps = 5
        t0 = time.perf_counter()
        for rep in range(reps):
            for f in range(1, fr2):
                loop.step_dev(dd_dev[f].data_ptr(), nd_dev.data_ptr())
        ctx.sync()
        out[key] = {"gpu_stream_frames_per_s": ns2 * reps * (fr2 - 1) / (time.perf_counter() - t0)}
        # host detections every frame (mot_tdd_step): pinned staging in two alternating sets, read by the one-launch frame kernel itself
        t0 = time.perf_counter()
        for rep in range(reps):
            for f in range(1, fr2):
                loop.step([dd[f][s_] for s_ in range(ns2)])
        ctx.sync()
        out[key]["host_dets_stream_frames_per_s"] = ns2 * reps * (fr2 - 1) / (time.perf_counter() - t0)
        loop.close(); ctx.close()
    return {"config": "C2: Kalman + Hungarian, 64 tracks x 64 detections, one 1080p stream, %d frames (host-array frame loop, one sync per stage)" % frames, **out}


def use_mkl_fft_for_the_cpu_side():
    """The compiled reference links FFTW; the oracle build shims it with a slow but exact double DFT by default and with MKL's DFTI
    (the fastest FFT on the box, out of libtorch_cpu.so) on request: timing comparisons use the latter, like bench.py does."""
    try:
        import torch
        lib = os.path.join(os.path.dirname(torch.__file__), "lib", "libtorch_cpu.so")
        if os.path.exists(lib):
            os.environ.setdefault("REF_FFT_PROVIDER", "mkl"); os.environ.setdefault("REF_FFT_MKL_LIB", lib)
            os.environ.setdefault("MKL_NUM_THREADS", "1"); os.environ.setdefault("OMP_NUM_THREADS", "1")
    except Exception:
        pass


def config1(args):
    """C1: single-target KCF on one synthetic 640x480 300-frame sequence with a seeded 128x128 box: upload + predict + update per
    frame through the host-array ABI (three synchronising calls per frame: latency, not throughput) vs the CPU restatement."""
    import ctypes as C
    import mot_b200 as M
    import oraclelib
    from synth import Scene, boxes_array, BBox
    from gpu_common import crop_gray
    W, H, F = 640, 480, 300
    sc = Scene(0x5EED0100, W, H, 1, tsize=51, win=128, vmax=2.0)
    sc.pos[:] = [[320.0, 240.0]]
    frames = []
    for f in range(F):
        frames.append(sc.render()); sc.step()
    import torch
    pin = torch.from_numpy(np.stack(frames)).pin_memory().numpy()      # frames in pinned host memory (pageable: the driver stages every upload)
    frames = [pin[f] for f in range(F)]
    b = boxes_array(1); b["l"], b["t"], b["r"], b["b"], b["type"], b["score"] = 256, 176, 383, 303, 1, 1.0
    ctx = M.Context(W, H, max_tracks=2, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frames[0]); h = ctx.new(b); ctx.update(h, [0], b)
    g = b.copy()
    t0 = time.perf_counter()
    for f in range(1, F):
        ctx.upload(0, frames[f]); g = ctx.predict(h, [0], g, clamp=1); ctx.update(h, [0], g)
    gpu = (F - 1) / (time.perf_counter() - t0)
    ctx.close()
    # the same sequence with predict + update as ONE call (mot_track_batch): two synchronising calls per frame instead of three
    ctx = M.Context(W, H, max_tracks=2, n_frame_slots=2, kind=M.TRACKER_KCF)
    ctx.upload(0, frames[0]); h = ctx.new(b); ctx.update(h, [0], b)
    g2 = b.copy()
    t0 = time.perf_counter()
    for f in range(1, F):
        ctx.upload(f & 1, frames[f]); g2 = ctx.track(h, [f & 1], g2, clamp=1)
    gpu2 = (F - 1) / (time.perf_counter() - t0)
    ctx.close()
    orc = oraclelib.Oracle(oraclelib.best())
    fft = orc.kcf.ref_fft_provider().decode() if orc.kind == "ref" else "dft64 (restatement)"
    ob = BBox(256, 176, 303, 383, 1, 1.0); oh = orc.kcf_new(ob)
    orc.kcf_update(oh, crop_gray(orc, frames[0], ob, 128, 128), ob)
    t0 = time.perf_counter()
    for f in range(1, F):
        orc.kcf_predict(oh, crop_gray(orc, frames[f], ob, 128, 128), ob)
        ob.l, ob.r = min(max(0, ob.l), W - 1), min(max(0, ob.r), W - 1); ob.t, ob.b = min(max(0, ob.t), H - 1), min(max(0, ob.b), H - 1)
        orc.kcf_update(oh, crop_gray(orc, frames[f], ob, 128, 128), ob)
    cpu = (F - 1) / (time.perf_counter() - t0)
    same = tuple(int(g[0][k]) for k in "ltbr") == ob.tup() and tuple(int(g2[0][k]) for k in "ltbr") == ob.tup()
    orc.kcf_delete(oh)
    return {"config": "C1: single-target KCF, 640x480, %d frames, one 128x128 window" % F, "gpu_frames_per_s_host_api": gpu, "gpu_frames_per_s_track_call": gpu2,
            "cpu_frames_per_s_one_core": cpu, "cpu_fft": fft, "identical_final_box": bool(same)}


def config3(args):
    """C3: 256 concurrent KCF tracks in one 1080p stream, the whole frame loop (predict, association, update, lifecycle):
    host-side loop, device-resident loop, and the CPU restatement on one core; then the device-resident loop on 64 streams x 128.
    The first WARM frames are not timed (first launches of every kernel, lazy allocations)."""
    import mot_b200 as M
    import oraclelib
    from synth import Scene
    W, H, F, WARM, FCPU = 1920, 1080, 28, 4, 8
    sc = Scene(0x5EED0300, W, H, 256, tsize=56, win=128)
    frames, dets = [], []
    for f in range(F):
        sc.step(); frames.append(sc.render()); dets.append(sc.windows(jitter=2))
    # the timed loops read their frames from PINNED host memory (like bench.py's e2e leg); the same loop fed from pageable arrays,
    # where every upload is staged by the driver (a 6.2 MB host memcpy per frame), is reported next to it
    import torch
    pin = torch.from_numpy(np.stack(frames)).pin_memory().numpy()
    pageable, frames = frames, [pin[f] for f in range(F)]
    out = {}
    ctx = M.Context(W, H, max_tracks=512, n_frame_slots=1, kind=M.TRACKER_KCF)
    td = ctx.td(0, cap=256, cost_mode=0)
    host_tab = None
    for f in range(F):
        if f == WARM:
            t0 = time.perf_counter()
        td.step(frames[f], dets[f])
        if f == FCPU - 1:
            host_tab = td.tracks()
    out["gpu_frames_per_s_host_loop"] = (F - WARM) / (time.perf_counter() - t0)
    td.close(); ctx.close()
    ctx = M.Context(W, H, max_tracks=256, n_frame_slots=2, kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(ctx, 1, cap=256, max_det=256, cost_mode=0)
    loop.kcf_windows([(128, 128)])
    dev_tab = None
    for f in range(F):
        if f == WARM:
            ctx.sync(); t0 = time.perf_counter()
        ctx.upload(f & 1, frames[f]); loop.frame_base(f & 1); loop.step([dets[f]])       # two slots: frame f uploads under the kernels of f-1
        if f == FCPU - 1:
            dev_tab = loop.tracks(0)
    ctx.sync()
    out["gpu_frames_per_s_device_loop"] = (F - WARM) / (time.perf_counter() - t0)
    loop.close(); ctx.close()
    ctx = M.Context(W, H, max_tracks=256, n_frame_slots=2, kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(ctx, 1, cap=256, max_det=256, cost_mode=0)
    loop.kcf_windows([(128, 128)])
    for f in range(F):
        if f == WARM:
            ctx.sync(); t0 = time.perf_counter()
        ctx.upload(f & 1, pageable[f]); loop.frame_base(f & 1); loop.step([dets[f]])
    ctx.sync()
    out["gpu_frames_per_s_device_loop_pageable_frames"] = (F - WARM) / (time.perf_counter() - t0)
    loop.close(); ctx.close()
    orc = oraclelib.Oracle(oraclelib.best())
    ref = orc.td_new("kcf", W, H, 256, 0)
    ref.step(frames[0], dets[0])
    t0 = time.perf_counter()
    for f in range(1, FCPU):
        ref.step(frames[f], dets[f])
    out["cpu_frames_per_s_one_core"] = (FCPU - 1) / (time.perf_counter() - t0)
    rt = ref.tracks()
    out["identical_track_table_after_%d_frames" % FCPU] = bool(all(np.array_equal(host_tab[k], rt[k]) and np.array_equal(dev_tab[k], rt[k]) for k in ("tid", "boxes", "age")))
    ref.close()
    # 64 streams x 128 tracks, frames resident (uploaded once): the loop itself, no PCIe in the timed region
    ns = 64
    sc2 = Scene(0x5EED0400, W, H, 128, tsize=96, win=128)
    fr = sc2.render(); d2 = sc2.windows(jitter=0)
    ctx = M.Context(W, H, max_tracks=ns * 128, n_frame_slots=ns, kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(ctx, ns, cap=128, max_det=128, cost_mode=0)
    loop.kcf_windows([(128, 128)])
    for s_ in range(ns):
        ctx.upload(s_, fr)
    dl = [d2] * ns
    loop.step(dl); loop.step(dl); ctx.sync()
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        loop.step(dl)
    ctx.sync()
    dt = (time.perf_counter() - t0) / K
    out["device_loop_64_streams_x_128_tracks"] = {"ms_per_frame_step": dt * 1e3, "stream_frames_per_s": ns / dt, "track_updates_per_s": ns * 128 / dt}
    loop.close(); ctx.close()
    return {"config": "C3: multi-target KCF, 256 tracks (128x128 px) in one 1080p stream, whole frame loop, %d frames timed after %d warm-up frames" % (F - WARM, WARM), **out}


if __name__ == "__main__":
    use_mkl_fft_for_the_cpu_side()
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrices", type=int, default=1024)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--cpu-matrices", type=int, default=2)
    args = ap.parse_args()
    print(json.dumps(config1(args)))
    print(json.dumps(config3(args)))
    print(json.dumps(config2(args)))
    print(json.dumps(config5(args)))
