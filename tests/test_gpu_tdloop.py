"""GPU parity: the batched frame loop (host/td_loop.cpp over the C ABI) vs the oracle's loop, both tracker kinds."""
import numpy as np
import pytest

from synth import Scene
from gpu_common import require_gpu, mot

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tracker,mode", [("kal", 0), ("kal", 1), ("kcf", 0)])
def test_frame_loop_trace(oracle, tracker, mode):
    require_gpu()
    M = mot()
    W, H = 1280, 720
    n = 24 if tracker == "kal" else 10
    sc = Scene(211 + mode, W, H, n, tsize=40, win=64)
    ctx = M.Context(W, H, max_tracks=128, n_frame_slots=1, kind=M.TRACKER_KCF if tracker == "kcf" else M.TRACKER_KALMAN)
    td = ctx.td(0, cap=64, cost_mode=mode)
    ref = oracle.td_new(tracker, W, H, 64, mode)
    drng = np.random.default_rng(4)
    for f in range(60 if tracker == "kal" else 25):
        sc.step()
        frame = sc.render()
        dets = sc.windows(jitter=2)
        keep = drng.random(len(dets)) > 0.1
        dets = np.ascontiguousarray(dets[keep])
        if f % 7 == 3:          # a false positive now and then (spawns a short-lived track)
            fp = dets[:1].copy(); fp["l"] += 300; fp["r"] += 300; fp["l"] %= (W - 80); fp["r"] = fp["l"] + 63
            dets = np.ascontiguousarray(np.concatenate([dets, fp]))
        td.step(frame, dets); ref.step(frame, dets)
        a, b = td.tracks(), ref.tracks()
        for k in a:
            assert np.array_equal(a[k], b[k]), (f, k)
        pa, aa = td.last(); pb, ab = ref.last()
        assert np.array_equal(pa, pb), f
        assert np.array_equal(aa, ab), f
    td.close(); ref.close(); ctx.close()


def test_frame_loop_with_detections_of_any_size(oracle):
    """Detections whose sizes change from frame to frame: templates of arbitrary size (any-size path) and the
    reference's scrambled resize on every update whose box differs from the template."""
    require_gpu()
    M = mot()
    W, H = 1280, 720
    sc = Scene(77, W, H, 8, tsize=40, win=64)
    ctx = M.Context(W, H, max_tracks=128, n_frame_slots=1, kind=M.TRACKER_KCF)
    td = ctx.td(0, cap=64, cost_mode=0)
    ref = oracle.td_new("kcf", W, H, 64, 0)
    rng = np.random.default_rng(8)
    grow = rng.integers(-6, 30, size=(8, 2))
    for f in range(14):
        sc.step()
        frame = sc.render()
        dets = sc.windows(jitter=1)
        jig = rng.integers(-2, 3, size=(8, 2))
        dets["r"] = np.clip(dets["r"] + grow[:, 0] + jig[:, 0], dets["l"] + 16, W - 1)
        dets["b"] = np.clip(dets["b"] + grow[:, 1] + jig[:, 1], dets["t"] + 16, H - 1)
        td.step(frame, dets); ref.step(frame, dets)
        a, b = td.tracks(), ref.tracks()
        for k in a:
            assert np.array_equal(a[k], b[k]), (f, k)
        pa, aa = td.last(); pb, ab = ref.last()
        assert np.array_equal(pa, pb) and np.array_equal(aa, ab), f
    td.close(); ref.close(); ctx.close()


def test_multi_stream_lockstep_equals_single(oracle):
    require_gpu()
    M = mot()
    W, H = 640, 480
    ns = 3
    scs = [Scene(900 + s, W, H, 5, tsize=32, win=64) for s in range(ns)]
    ctx = M.Context(W, H, max_tracks=128, n_frame_slots=ns, kind=M.TRACKER_KCF)
    tds = [ctx.td(s, cap=32) for s in range(ns)]
    refs = [oracle.td_new("kcf", W, H, 32, 0) for _ in range(ns)]
    for f in range(10):
        frames, dets = [], []
        for sc in scs:
            sc.step(); frames.append(sc.render()); dets.append(sc.windows(jitter=1))
        M.step_multi(tds, frames, dets)
        for s in range(ns):
            refs[s].step(frames[s], dets[s])
            a, b = tds[s].tracks(), refs[s].tracks()
            for k in a:
                assert np.array_equal(a[k], b[k]), (f, s, k)
    for t in tds:
        t.close()
    for r in refs:
        r.close()
    ctx.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_device_resident_loop_trace(oracle, mode):
    """mot_tdd_*: track tables, bookkeeping, stable compaction and spawn on the device; five launches per frame, no host
    sync.  Trace-identical to the oracle's loop for several streams at once, through births, misses and deaths."""
    require_gpu()
    M = mot()
    W, H, ns, cap = 1280, 720, 5, 48
    scs = [Scene(400 + 3 * s + mode, W, H, 20, tsize=40, win=64) for s in range(ns)]
    ctx = M.Context(W, H, max_tracks=ns * cap, kind=M.TRACKER_KALMAN)
    loop = M.DeviceLoop(ctx, ns, cap=cap, max_det=64, cost_mode=mode)
    refs = [oracle.td_new("kal", W, H, cap, mode) for _ in range(ns)]
    drng = np.random.default_rng(17)
    for f in range(70):
        dets = []
        for s, sc in enumerate(scs):
            sc.step()
            d = sc.windows(jitter=2)
            keep = drng.random(len(d)) > (0.5 if 20 <= f < 45 and s % 2 == 0 else 0.1)     # a long drought kills tracks on even streams
            d = np.ascontiguousarray(d[keep])
            if f % 5 == 2:
                fp = d[:2].copy(); fp["l"] = (fp["l"] + 333) % (W - 80); fp["r"] = fp["l"] + 63
                d = np.ascontiguousarray(np.concatenate([d, fp]))
            if f == 30 and s == 1:
                d = d[:0]                                                           # a frame without detections
            dets.append(d)
        loop.step(dets)
        for s in range(ns):
            refs[s].step(None, dets[s])
        if f % 3 == 0 or f > 60:
            for s in range(ns):
                a, b = loop.tracks(s), refs[s].tracks()
                for k in a:
                    assert np.array_equal(a[k], b[k]), (f, s, k)
    loop.close()
    for r in refs:
        r.close()
    ctx.close()


def test_device_resident_loop_separate_launches_for_large_tables(oracle):
    """Streams with more than 256 tracks / detections keep the separate launches (1024-thread solver) instead of the one-launch frame
    kernel: same traces as the oracle's loop, and the same as the one-launch form on a table that both can run (MOT_TDD_UNFUSED)."""
    require_gpu()
    M = mot()
    W, H, cap = 1920, 1080, 320
    sc = Scene(4242, W, H, 280, tsize=24, win=32)
    ctx = M.Context(W, H, max_tracks=cap, kind=M.TRACKER_KALMAN)
    loop = M.DeviceLoop(ctx, 1, cap=cap, max_det=320, cost_mode=0)
    ref = oracle.td_new("kal", W, H, cap, 0)
    drng = np.random.default_rng(3)
    for f in range(12):
        sc.step()
        d = sc.windows(jitter=1)
        d = np.ascontiguousarray(d[drng.random(len(d)) > 0.05])
        loop.step([d]); ref.step(None, d)
        a, b = loop.tracks(0), ref.tracks()
        assert len(a["tid"]) > 256
        for k in a:
            assert np.array_equal(a[k], b[k]), (f, k)
    loop.close(); ref.close(); ctx.close()
    import os
    tabs = []
    for unfused in (False, True):
        if unfused:
            os.environ["MOT_TDD_UNFUSED"] = "1"
        try:
            sc = Scene(99, W, H, 40, tsize=24, win=32)
            ctx = M.Context(W, H, max_tracks=64, kind=M.TRACKER_KALMAN)
            loop = M.DeviceLoop(ctx, 1, cap=64, max_det=64, cost_mode=1)
            for f in range(15):
                sc.step(); loop.step([sc.windows(jitter=2)])
            tabs.append(loop.tracks(0))
            assert ctx.launches() >= (15 * 6 if unfused else 15) and (unfused or ctx.launches() < 15 * 3)
            loop.close(); ctx.close()
        finally:
            os.environ.pop("MOT_TDD_UNFUSED", None)
    for k in tabs[0]:
        assert np.array_equal(tabs[0][k], tabs[1][k]), k


def test_device_resident_kcf_loop_trace(oracle):
    """The KCF kind of mot_tdd_*: job lists per window class built on the device, fused predict / update over them,
    tracker_new + first update inside the loop.  Trace-identical to the oracle's KCF loop for three streams with different
    window classes (64x64 and 128x128 px detections in the same context) through births, misses and deaths; a detection
    whose window has no fused kernel is dropped and counted, and changes nothing else."""
    require_gpu()
    M = mot()
    W, H, ns, cap = 1280, 720, 3, 24
    wins = [64, 128, 64]
    scs = [Scene(900 + 7 * s, W, H, 6 if wins[s] == 128 else 9, tsize=wins[s] * 5 // 8, win=wins[s]) for s in range(ns)]
    ctx = M.Context(W, H, max_tracks=(ns + 1) * cap, n_frame_slots=2 * (ns + 1), kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(ctx, ns + 1, cap=cap, max_det=32, cost_mode=0)       # stream ns stays empty until the last step
    loop.kcf_windows([(64, 64), (128, 128)])
    refs = [oracle.td_new("kcf", W, H, cap, 0) for _ in range(ns)]
    drng = np.random.default_rng(23)
    for f in range(26):
        dets, frames = [], []
        for s, sc in enumerate(scs):
            sc.step()
            frames.append(sc.render())
            d = sc.windows(jitter=2)
            keep = drng.random(len(d)) > (0.6 if 8 <= f < 20 and s == 0 else 0.1)         # a drought on stream 0: tracks die
            d = np.ascontiguousarray(d[keep])
            if f % 6 == 2 and len(d):                                                      # a false positive: a short-lived track
                fp = d[:1].copy(); fp["l"] = (fp["l"] + 311) % (W - 140); fp["r"] = fp["l"] + wins[s] - 1
                d = np.ascontiguousarray(np.concatenate([d, fp]))
            dets.append(d)
        base = (f & 1) * (ns + 1)                          # two sets of frame slots: the step of frame f-1 may still be reading the other one
        for s in range(ns):
            ctx.upload(base + s, frames[s])
        ctx.upload(base + ns, frames[0])
        loop.frame_base(base)
        loop.step(dets + [dets[0][:0]])
        for s in range(ns):
            refs[s].step(frames[s], dets[s])
        if f % 2 == 0 or f > 20:
            for s in range(ns):
                a, b = loop.tracks(s), refs[s].tracks()
                for k in a:
                    assert np.array_equal(a[k], b[k]), (f, s, k)
    assert all(loop.dropped(s) == 0 for s in range(ns + 1))
    # a detection of a size without a fused kernel on the empty stream: not spawned, counted; the other streams carry on
    odd = dets[0][:1].copy(); odd["l"] = 20; odd["r"] = 20 + 99; odd["t"] = 30; odd["b"] = 30 + 59
    before = loop.tracks(1)
    loop.step([dets[0], dets[1][:0], dets[2], odd])
    assert loop.dropped(ns) == 1 and loop.dropped(0) == 0 and len(loop.tracks(ns)["tid"]) == 0
    after = loop.tracks(1)
    assert np.array_equal(after["tid"], before["tid"]) and np.array_equal(after["inv"], before["inv"] + 1)
    loop.close()
    for r in refs:
        r.close()
    ctx.close()


def test_device_resident_kcf_loop_with_detections_of_any_size(oracle):
    """The device-resident KCF loop with detections of ARBITRARY sizes that change from frame to frame: every tracker is born on the
    device whatever its window (fixed-size fused classes and the any-size kernel's job lists side by side), its constants come
    from the per-N tables, and every update whose box differs from the template goes through the reference's scrambled resize.
    Trace-identical to the oracle's loop; nothing is dropped."""
    require_gpu()
    M = mot()
    W, H, ns, cap = 1280, 720, 2, 32
    scs = [Scene(77 + 5 * s, W, H, 8, tsize=40, win=64) for s in range(ns)]
    ctx = M.Context(W, H, max_tracks=ns * cap, n_frame_slots=2 * ns, kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(ctx, ns, cap=cap, max_det=32, cost_mode=0)
    refs = [oracle.td_new("kcf", W, H, cap, 0) for _ in range(ns)]
    rng = np.random.default_rng(8)
    grow = [rng.integers(-6, 90, size=(8, 2)) for _ in range(ns)]            # windows from 58 to 154 px: fused 64/128 classes, small and large any-size ones
    for f in range(14):
        dets, frames = [], []
        for s, sc in enumerate(scs):
            sc.step()
            frames.append(sc.render())
            d = sc.windows(jitter=1)
            jig = rng.integers(-2, 3, size=(8, 2))
            d["r"] = np.clip(d["r"] + grow[s][:, 0] + jig[:, 0], d["l"] + 16, W - 1)
            d["b"] = np.clip(d["b"] + grow[s][:, 1] + jig[:, 1], d["t"] + 16, H - 1)
            if f in (4, 5) and s == 0:
                d = d[:5]                                                       # misses
            dets.append(np.ascontiguousarray(d))
        base = (f & 1) * ns
        for s in range(ns):
            ctx.upload(base + s, frames[s])
        loop.frame_base(base)
        loop.step(dets)
        for s in range(ns):
            refs[s].step(frames[s], dets[s])
            a, b = loop.tracks(s), refs[s].tracks()
            for k in a:
                assert np.array_equal(a[k], b[k]), (f, s, k)
    assert all(loop.dropped(s) == 0 for s in range(ns))
    sizes = {(int(b["b"] - b["t"] + 1) // 4, int(b["r"] - b["l"] + 1) // 4) for d in dets for b in d}
    assert len(sizes) >= 6, "the test must really mix window sizes"
    loop.close()
    for r in refs:
        r.close()
    ctx.close()


def test_device_resident_kcf_loop_spawns_strip_mode_windows(oracle):
    """Detections the size of real detector boxes (150..300 px per side: 1400..5600 cells, beyond one CTA's shared memory) are born on
    the device too: the strip-mode job list of the any-size kernel with the loop's own per-CTA scratch, models in an arena sized by
    mot_ctx_reserve_window.  Trace-identical to the oracle's loop; nothing is dropped.  Without the reservation the same detections
    whose spectra exceed the default slot are counted as dropped and nothing else happens."""
    require_gpu()
    M = mot()
    W, H, cap = 1280, 720, 8
    sc = Scene(314, W, H, 3, tsize=110, win=160)
    ctx = M.Context(W, H, max_tracks=cap, n_frame_slots=2, kind=M.TRACKER_KCF)
    ctx.reserve_window(320, 320)
    loop = M.DeviceLoop(ctx, 1, cap=cap, max_det=8, cost_mode=0)
    ref = oracle.td_new("kcf", W, H, cap, 0)
    rng = np.random.default_rng(5)
    grow = np.array([[0, 60], [139, 20], [40, 110]])                              # 160x220, 299x180, 200x270 px windows (40x55, 74x45, 50x67 cells)
    for f in range(6):
        sc.step()
        frame = sc.render()
        d = sc.windows(jitter=1)
        jig = rng.integers(-2, 3, size=(3, 2))
        d["r"] = np.clip(d["r"] + grow[:, 0] + jig[:, 0], d["l"] + 16, W - 1)
        d["b"] = np.clip(d["b"] + grow[:, 1] + jig[:, 1], d["t"] + 16, H - 1)
        if f == 3:
            d = d[:2]
        d = np.ascontiguousarray(d)
        ctx.upload(f & 1, frame)
        loop.frame_base(f & 1)
        loop.step([d])
        ref.step(frame, d)
        a, b = loop.tracks(0), ref.tracks()
        assert len(a["tid"]) >= 2
        for k in a:
            assert np.array_equal(a[k], b[k]), (f, k)
    assert loop.dropped(0) == 0
    from test_any_plan import plan
    plans = [plan((int(b["b"] - b["t"] + 1)) // 4, (int(b["r"] - b["l"] + 1)) // 4) for b in d]
    assert any(p["strips"] for p in plans), "the test must reach strip mode"
    loop.close(); ref.close(); ctx.close()
    # the default arena: 1152 bins per slot
    ctx = M.Context(W, H, max_tracks=cap, n_frame_slots=2, kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(ctx, 1, cap=cap, max_det=8, cost_mode=0)
    ctx.upload(0, frame)
    loop.step([d])
    big = sum(((int(b["r"] - b["l"] + 1)) // 4) * (((int(b["b"] - b["t"] + 1)) // 4) // 2 + 1) > 1152 for b in d)
    assert big >= 2 and loop.dropped(0) == big and len(loop.tracks(0)["tid"]) == len(d) - big
    loop.close(); ctx.close()


def test_detector_wire_format_gives_the_same_trace(oracle):
    """detections arrive as bbox_chain_t (top/cnntype.h:43-47), in batches of up to four frames as tensorRunB delivers
    them (top/td.cpp:178-204).  Feeding chains -- one by one, batched, and to the device-resident loop -- gives exactly the track tables
    that flat arrays give, which are the oracle's."""
    require_gpu()
    M = mot()
    assert M.CHAIN_DTYPE.itemsize == 4 + 128 * 24
    W, H = 1280, 720
    sc = Scene(41, W, H, 7, tsize=40, win=64)
    frames, dets = [], []
    for f in range(9):
        sc.step(); frames.append(sc.render()); dets.append(sc.windows(jitter=1))
    ctx = M.Context(W, H, max_tracks=4 * 32, n_frame_slots=4, kind=M.TRACKER_KCF)
    td_flat, td_chain, td_batch = ctx.td(0, cap=32), ctx.td(1, cap=32), ctx.td(2, cap=32)
    loop_ctx = M.Context(W, H, max_tracks=32, n_frame_slots=1, kind=M.TRACKER_KCF)
    loop = M.DeviceLoop(loop_ctx, 1, cap=32, max_det=128, cost_mode=0)
    ref = oracle.td_new("kcf", W, H, 32, 0)
    chains = [M.make_chain(d) for d in dets]
    for f in range(9):
        td_flat.step(frames[f], dets[f]); td_chain.step_chain(frames[f], chains[f]); ref.step(frames[f], dets[f])
        loop_ctx.upload(0, frames[f]); loop.step_chains([chains[f]]); loop_ctx.sync()
    for f0 in (0, 4, 8):                                      # batches of 4, 4 and 1 frames
        td_batch.step_chain_batch(frames[f0:f0 + 4], chains[f0:f0 + 4])
    want = ref.tracks()
    for got in (td_flat.tracks(), td_chain.tracks(), td_batch.tracks(), loop.tracks(0)):
        for k in got:
            assert np.array_equal(got[k], want[k]), k
    assert len(want["tid"]) >= 5
    loop.close(); loop_ctx.close()
    for t in (td_flat, td_chain, td_batch):
        t.close()
    ref.close(); ctx.close()
