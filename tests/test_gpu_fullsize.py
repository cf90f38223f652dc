"""BASELINE.json's full sizes on the GPU, checked through size-independent properties plus oracle samples.

  config 3/4  8192 KCF tracks (64 1080p streams x 128 tracks, 128x128 px windows) in one launch per stage:
              determinism (two contexts, same inputs -> same bytes), a repeated predict on the same crop moves the
              position by the same response shift again (kcf.cpp:419-426 updates pos, not the model), window sizes are
              preserved, sampled tracks equal the oracle's tracker driven with the same crops
  config 5    1024 association problems of 512x512: every result is a permutation, the returned cost is the sum of the
              chosen entries in row order (the reference's accumulation order), identical across two runs; samples vs oracle"""
import numpy as np
import pytest

from synth import Scene, boxes_array, random_boxes
from gpu_common import require_gpu, mot, box_of, crop_gray

pytestmark = pytest.mark.gpu

W, H = 1920, 1080


def _streams(ns, nt, seed):
    """ns streams x 2 frames (targets move by a few pixels) + the initial 128x128 windows; four scenes re-used across streams."""
    frames = np.zeros((ns, 2, H, W, 3), np.uint8)
    boxes = []
    base = [Scene(seed + k, W, H, nt, tsize=96, win=128) for k in range(4)]
    for s in range(ns):
        b = base[s % 4]
        f0 = b.render()
        sh = (1 + s % 3, -(1 + (s // 3) % 4))
        frames[s, 0] = np.roll(f0, (s % 5, s % 7), (0, 1))
        frames[s, 1] = np.roll(frames[s, 0], sh, (0, 1))
        w = b.windows().copy()
        w["l"] += s % 7; w["r"] += s % 7; w["t"] += s % 5; w["b"] += s % 5
        ok = (w["r"] < W) & (w["b"] < H)
        w["l"][~ok] = 100; w["r"][~ok] = 227; w["t"][~ok] = 100; w["b"][~ok] = 227
        boxes.append(w)
    return frames, boxes


def test_config4_full_size_properties(oracle):
    require_gpu()
    M = mot()
    NS, NT = 64, 128
    n = NS * NT
    frames, boxes = _streams(NS, NT, 4400)
    all_boxes = np.ascontiguousarray(np.concatenate(boxes))
    fs = np.repeat(np.arange(NS, dtype=np.int32), NT)

    def run():
        ctx = M.Context(W, H, max_tracks=n, n_frame_slots=NS, kind=M.TRACKER_KCF)
        for s in range(NS):
            ctx.upload(s, frames[s, 0])
        h = ctx.new(all_boxes)
        ctx.update(h, fs, all_boxes)                        # first update: model = features of frame 0
        for s in range(NS):
            ctx.upload(s, frames[s, 1])
        p1 = ctx.predict(h, fs, all_boxes, clamp=0)
        p2 = ctx.predict(h, fs, all_boxes, clamp=0)         # same crop again: same response, pos moves by the same shift again
        ctx.update(h, fs, p1)                               # second update (lerp 0.05) at the predicted boxes
        p3 = ctx.predict(h, fs, p1, clamp=1)
        ctx.close()
        return p1, p2, p3

    p1, p2, p3 = run()
    q1, q2, q3 = run()
    assert p1.tobytes() == q1.tobytes() and p2.tobytes() == q2.tobytes() and p3.tobytes() == q3.tobytes(), "two identical runs differ"
    # predict touches pos only: the second call sees the same crop and the same model, hence the same integer cell shift;
    # the shift is a whole number of pixels here (scale 1, 4-pixel cells), so pos advances by exactly the same amount
    for k in "ltbr":
        assert np.array_equal(p2[k] - p1[k], p1[k] - all_boxes[k]), k
    moved = (p1["l"] != all_boxes["l"]) | (p1["t"] != all_boxes["t"])
    assert moved.mean() > 0.5, "targets moved between the frames; most boxes must follow"
    # window size is preserved by the box shift (kcf.cpp:423-426 adds the same offset to both corners)
    assert np.all(p1["r"] - p1["l"] == 127) and np.all(p1["b"] - p1["t"] == 127)
    # sampled tracks against the oracle's tracker fed with the same crops
    rng = np.random.default_rng(7)
    ok = np.ones(n, bool)
    for q in (p1, p2, p3):                               # the oracle-side crop helper does not clamp: keep to boxes inside the frame
        ok &= (q["l"] >= 0) & (q["t"] >= 0) & (q["r"] < W) & (q["b"] < H)
    assert ok.mean() > 0.9
    for i in rng.choice(np.flatnonzero(ok), 24, replace=False):
        s = int(fs[i])
        ob = box_of(all_boxes[i]); oh = oracle.kcf_new(ob)
        oracle.kcf_update(oh, crop_gray(oracle, frames[s, 0], ob, 128, 128), ob)
        b1 = box_of(all_boxes[i])
        oracle.kcf_predict(oh, crop_gray(oracle, frames[s, 1], b1, 128, 128), b1)
        assert tuple(int(p1[i][k]) for k in "ltbr") == (b1.l, b1.t, b1.b, b1.r), i
        b2 = box_of(all_boxes[i])
        oracle.kcf_predict(oh, crop_gray(oracle, frames[s, 1], b2, 128, 128), b2)
        assert tuple(int(p2[i][k]) for k in "ltbr") == (b2.l, b2.t, b2.b, b2.r), i
        u = box_of(p1[i])
        oracle.kcf_update(oh, crop_gray(oracle, frames[s, 1], u, 128, 128), u)
        b3 = box_of(p1[i])
        oracle.kcf_predict(oh, crop_gray(oracle, frames[s, 1], b3, 128, 128), b3)
        want3 = (min(max(0, b3.l), W - 1), min(max(0, b3.t), H - 1), min(max(0, b3.b), H - 1), min(max(0, b3.r), W - 1))
        assert tuple(int(p3[i][k]) for k in "ltbr") == want3, i
        oracle.kcf_delete(oh)


def test_config5_full_size_properties(oracle):
    require_gpu()
    M = mot()
    NM, T = 1024, 512
    rng = np.random.default_rng(0xC5)
    trks = [random_boxes(rng, T, W, H) for _ in range(8)]
    dets = [random_boxes(rng, T, W, H) for _ in range(8)]
    # 1024 problems: every (tracks, detections) pairing of the 8 + 8 box sets, rolled by the problem index
    tl = [trks[m % 8] if m < 64 else np.roll(trks[m % 8], m) for m in range(NM)]
    dl = [dets[(m // 8) % 8] if m < 64 else np.roll(dets[(m // 8) % 8], 3 * m + 1) for m in range(NM)]
    ctx = M.Context(W, H, max_tracks=4, kind=M.TRACKER_KALMAN)
    CH = 256                                             # four calls of 256 problems keep the host copies of the matrices at 0.5 GB
    for c0 in range(0, NM, CH):
        a1, c1, d1 = ctx.associate(tl[c0:c0 + CH], dl[c0:c0 + CH], cost_mode=1, want_dist=True)
        a2, c2 = ctx.associate(tl[c0:c0 + CH], dl[c0:c0 + CH], cost_mode=1)
        rows = np.arange(T)
        for m in range(CH):
            a = np.asarray(a1[m])
            assert np.array_equal(a, a2[m]) and c1[m] == c2[m]
            assert np.array_equal(np.sort(a), rows), "square problem: the assignment is a permutation"
            acc = 0.0
            for v in d1[m][rows, a]:                     # hungarian.cpp:185-196 sums the chosen entries in row order
                acc += float(v)
            assert acc == c1[m], c0 + m
        for m in rng.choice(CH, 1, replace=False):
            a_or, c_or = oracle.assign(d1[m])
            assert np.array_equal(a1[m], a_or) and c1[m] == c_or
    ctx.close()
