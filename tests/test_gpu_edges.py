"""GPU parity on the edges of the path: empty and ragged batches, extreme sizes, frame borders, reversed boxes,
untrained trackers, slot reuse and capacity errors.  The reference has no tests; the cases follow its code paths
(top/td.cpp:460 skips the solver for empty sides; top/drawlib.c:203-215 swaps reversed corners; trackers/kcf.cpp:172-174
starts from a zero model; hungarian.cpp:65/104 picks row or column reduction by shape)."""
import numpy as np
import pytest

from synth import BBox, Scene, BBOX_DTYPE, boxes_array, random_boxes, jittered_detections
from gpu_common import require_gpu, mot, box_of, crop_gray

pytestmark = pytest.mark.gpu


def one_box(l, t, rows, cols, typ=1):
    b = boxes_array(1)
    b["l"], b["t"], b["r"], b["b"], b["type"], b["score"] = l, t, l + cols - 1, t + rows - 1, typ, 1.0
    return b


def test_empty_batches_are_noops():
    require_gpu()
    M = mot()
    for kind in (M.TRACKER_KCF, M.TRACKER_KALMAN):
        ctx = M.Context(640, 480, max_tracks=4, kind=kind)
        e = boxes_array(0)
        assert len(ctx.new(e)) == 0
        assert len(ctx.predict(np.zeros(0, np.int32), np.zeros(0, np.int32), e)) == 0
        ctx.update(np.zeros(0, np.int32), np.zeros(0, np.int32), e)
        ctx.delete(np.zeros(0, np.int32))
        ctx.close()


def test_association_empty_sides_and_ragged_batch(oracle):
    """One batched call with problems of different shapes, including T = 0 and D = 0 (solver skipped, td.cpp:460)."""
    require_gpu()
    M = mot()
    rng = np.random.default_rng(21)
    ctx = M.Context(1920, 1080, max_tracks=4, kind=M.TRACKER_KALMAN)
    shapes = [(0, 5), (4, 0), (1, 1), (3, 17), (40, 12), (25, 25), (0, 0), (2, 300)]
    trks = [random_boxes(rng, max(T, 0), 1920, 1080) for T, D in shapes]
    dets = [random_boxes(rng, max(D, 0), 1920, 1080) for T, D in shapes]
    assigns, costs, dists = ctx.associate(trks, dets, cost_mode=M.COST_IOU_CLAMPED, want_dist=True)
    for m, (T, D) in enumerate(shapes):
        assert len(assigns[m]) == min(T, D)
        if T and D:
            a_or, c_or = oracle.assign(dists[m])
            assert np.array_equal(assigns[m], a_or), shapes[m]
            assert costs[m] == c_or
        else:
            assert costs[m] == 0.0
    ctx.close()


def test_association_maximum_size_1024(oracle):
    require_gpu()
    M = mot()
    rng = np.random.default_rng(22)
    ctx = M.Context(1920, 1080, max_tracks=4, kind=M.TRACKER_KALMAN)
    trk = random_boxes(rng, 1024, 1920, 1080)
    det = jittered_detections(rng, trk, 1920, 1080, jitter=3)
    assigns, costs, dists = ctx.associate([trk], [det], cost_mode=M.COST_REF_CENTROID, want_dist=True)
    a_or, c_or = oracle.assign(dists[0])
    assert np.array_equal(assigns[0], a_or) and costs[0] == c_or
    with pytest.raises(M.MotError):
        ctx.assign([np.zeros((1025, 3))])
    ctx.close()


def test_crops_touching_the_frame_border_and_reversed_corners(oracle):
    """Predicted boxes are clamped into the frame (td.cpp:378-381), which changes their size and sends the next crop through
    the scrambled resize; reversed corners are swapped by rgb2Gray (drawlib.c:203-215)."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    frame = Scene(5, W, H, 6, tsize=40, win=64).render()
    ctx = M.Context(W, H, max_tracks=8, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    cases = [((0, 0, 63, 63), (64, 64)), ((W - 64, H - 64, W - 1, H - 1), (64, 64)), ((0, 400, 50, 479), (64, 64)),
             ((600, 10, 639, 100), (128, 64)), ((200, 150, 327, 277), (128, 128))]
    for (l, t, r, b), (rd, cd) in cases:
        bx = boxes_array(1); bx["l"], bx["t"], bx["r"], bx["b"] = l, t, r, b
        got = ctx.crop_gray_resize(0, bx, rd, cd).reshape(cd, rd).T
        assert np.array_equal(got, crop_gray(oracle, frame, box_of(bx[0]), rd, cd)), (l, t, r, b)
        rv = boxes_array(1); rv["l"], rv["t"], rv["r"], rv["b"] = r, b, l, t          # reversed corners
        got2 = ctx.crop_gray_resize(0, rv, rd, cd).reshape(cd, rd).T
        assert np.array_equal(got2, got)
    ctx.close()


def test_predict_before_any_update_does_not_move(oracle):
    """xf_md and alpha start at zero (kcf.cpp:172-174): the response is identically zero, the first maximum is cell (1,1)."""
    require_gpu()
    M = mot()
    frame = Scene(6, 640, 480, 3, tsize=40, win=64).render()
    ctx = M.Context(640, 480, max_tracks=4, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    for (rows, cols) in [(64, 64), (60, 100)]:
        b = one_box(100, 100, rows, cols)
        h = ctx.new(b)
        out = ctx.predict(h, [0], b)
        ob = box_of(b[0]); oh = oracle.kcf_new(ob)
        oracle.kcf_predict(oh, crop_gray(oracle, frame, ob, rows, cols), ob)
        assert tuple(int(out[0][k]) for k in "ltbr") == ob.tup() == (100, 100, 100 + rows - 1, 100 + cols - 1)
        oracle.kcf_delete(oh)
    ctx.close()


def test_slot_capacity_reuse_and_invalid_handles():
    require_gpu()
    M = mot()
    ctx = M.Context(640, 480, max_tracks=3, n_frame_slots=1, kind=M.TRACKER_KCF)
    b3 = np.concatenate([one_box(10 + 70 * i, 10, 64, 64) for i in range(3)])
    h = ctx.new(b3)
    assert sorted(h.tolist()) == [0, 1, 2]
    with pytest.raises(M.MotError):
        ctx.new(b3[:1])                               # out of slots: an error, not a silent drop
    ctx.delete(h[1:2])
    h2 = ctx.new(b3[:1])                              # the freed slot is reused
    assert h2[0] == h[1]
    with pytest.raises(M.MotError):
        ctx.delete(np.array([7], np.int32))
    with pytest.raises(M.MotError):
        ctx.predict(np.array([0], np.int32), np.array([0], np.int32), b3[:1])      # frame slot 0 is still empty
    ctx.close()


def test_reserved_window_arena_gives_the_same_filter(oracle):
    """mot_ctx_reserve_window only changes WHERE a model lives (a larger slot of the arena instead of an individually allocated block):
    a 200x260 px window (50x65 cells, 1650 half-spectrum bins, strip mode) tracks identically with and without the reservation and
    like the compiled reference; the call is refused once a tracker exists."""
    require_gpu()
    M = mot()
    W, H, rows, cols = 1280, 720, 260, 200
    sc = Scene(12, W, H, 1, tsize=150, win=200, vmax=3.0)
    sc.pos[:] = [[400.0, 330.0]]
    frames = []
    for _ in range(2):
        frames.append(sc.render()); sc.step()
    outs = []
    for reserve in (False, True):
        ctx = M.Context(W, H, max_tracks=2, n_frame_slots=1, kind=M.TRACKER_KCF)
        if reserve:
            ctx.reserve_window(320, 320)
        b = one_box(300, 200, rows, cols)
        ctx.upload(0, frames[0]); h = ctx.new(b); ctx.update(h, [0], b)
        if reserve:
            with pytest.raises(M.MotError):
                ctx.reserve_window(400, 400)             # a tracker exists
        ctx.upload(0, frames[1]); g = ctx.predict(h, [0], b.copy(), clamp=1); ctx.update(h, [0], g)
        outs.append((g.copy(), ctx.state(h[0], "alpha").copy()))
        ctx.close()
    assert outs[0][0].tobytes() == outs[1][0].tobytes() and np.array_equal(outs[0][1], outs[1][1])
    ob = box_of(outs[0][0][0])
    rb = box_of(one_box(300, 200, rows, cols)[0]); oh = oracle.kcf_new(rb)
    oracle.kcf_update(oh, crop_gray(oracle, frames[0], rb, rows, cols), rb)
    oracle.kcf_predict(oh, crop_gray(oracle, frames[1], rb, rows, cols), rb)
    rb.l, rb.r = min(max(0, rb.l), W - 1), min(max(0, rb.r), W - 1); rb.t, rb.b = min(max(0, rb.t), H - 1), min(max(0, rb.b), H - 1)
    assert ob.tup() == rb.tup()
    oracle.kcf_delete(oh)


def test_many_tracks_one_launch_equals_per_track_oracle(oracle):
    """A 300-track batch (more CTAs than SMs: the persistent kernel loops) against per-track oracle calls."""
    require_gpu()
    M = mot()
    W, H = 1280, 720
    sc = Scene(33, W, H, 300, tsize=24, win=64)
    frame = sc.render()
    ctx = M.Context(W, H, max_tracks=512, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    b = sc.windows()
    h = ctx.new(b)
    fs = np.zeros(len(b), np.int32)
    ctx.update(h, fs, b)
    frame2 = np.ascontiguousarray(np.roll(frame, (4, -4), (0, 1)))
    ctx.upload(0, frame2)
    out = ctx.predict(h, fs, b, clamp=1)
    rng = np.random.default_rng(0)
    for i in rng.choice(len(b), 40, replace=False):
        ob = box_of(b[i]); oh = oracle.kcf_new(ob)
        oracle.kcf_update(oh, crop_gray(oracle, frame, ob, 64, 64), ob)
        oracle.kcf_predict(oh, crop_gray(oracle, frame2, ob, 64, 64), ob)
        want = (min(max(0, ob.l), W - 1), min(max(0, ob.t), H - 1), min(max(0, ob.b), H - 1), min(max(0, ob.r), W - 1))
        assert tuple(int(out[i][k]) for k in "ltbr") == want, i
        oracle.kcf_delete(oh)
    ctx.close()


def test_reference_plugin_signatures_through_the_shim(oracle):
    """host/tracker_shim.cpp: the reference's own per-object entry points (tracker_new/predict/update/delete with a HOST gray
    patch, assignmentoptimal with a host matrix) running on the GPU, compared call by call with the oracle."""
    require_gpu()
    import ctypes as C
    M = mot()
    L = M.lib()
    L.mot_shim_tracker_new.restype = C.c_void_p
    for kind in (M.TRACKER_KCF, M.TRACKER_KALMAN):
        assert L.mot_shim_configure(kind, 640, 480, 16) == 0, L.mot_last_error()
        rng = np.random.default_rng(4)
        img = (rng.random((60, 80)).repeat(8, 0).repeat(8, 1) * 200).astype(np.float32)
        b = BBox(96, 64, 64 + 63, 96 + 63, 1, 1.0); ob = BBox.from_buffer_copy(b)
        h = C.c_void_p(L.mot_shim_tracker_new(C.byref(b)))
        assert h.value
        oh = oracle.kcf_new(ob) if kind == M.TRACKER_KCF else oracle.kal_new(ob)
        for step in range(4):
            g = np.asfortranarray(img[b.t:b.b + 1, b.l:b.r + 1])
            og = np.asfortranarray(img[ob.t:ob.b + 1, ob.l:ob.r + 1])
            L.mot_shim_tracker_update(h, g.ctypes.data_as(C.c_void_p), C.byref(b))
            (oracle.kcf_update(oh, og, ob) if kind == M.TRACKER_KCF else oracle.kal_update(oh, ob))
            img = np.roll(img, (4, -4), (0, 1))
            g = np.asfortranarray(img[b.t:b.b + 1, b.l:b.r + 1]); og = np.asfortranarray(img[ob.t:ob.b + 1, ob.l:ob.r + 1])
            L.mot_shim_tracker_predict(h, g.ctypes.data_as(C.c_void_p), C.byref(b))
            (oracle.kcf_predict(oh, og, ob) if kind == M.TRACKER_KCF else oracle.kal_predict(oh, ob))
            assert b.tup() == ob.tup(), (kind, step)
        L.mot_shim_tracker_delete(h)
        (oracle.kcf_delete(oh) if kind == M.TRACKER_KCF else oracle.kal_delete(oh))
    d = np.asfortranarray(np.round(rng.random((23, 31)), 2))
    a = np.full(23, -5, np.int32); cost = C.c_double()
    L.mot_shim_assignmentoptimal(a.ctypes.data_as(C.c_void_p), C.byref(cost), d.ctypes.data_as(C.c_void_p), 23, 31)
    a_or, c_or = oracle.assign(d)
    assert np.array_equal(a, a_or) and cost.value == c_or
