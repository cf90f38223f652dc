"""GPU parity: the fused KCF kernels (through the C ABI) vs the oracle, stage by stage and over sequences.

Tolerances are BASELINE.json's: fHOG features <= 1e-4 relative, integer peak / int bbox bit-exact (the top-2 response
margin is checked so that FFT rounding differences cannot flip the peak), everything upstream of the FFT bit-exact.
"""
import numpy as np
import pytest

from synth import BBox, Scene, BBOX_DTYPE, boxes_array
from gpu_common import require_gpu, mot, oracle_bins, rel_err, box_of, crop_gray

pytestmark = pytest.mark.gpu


def one_box(l, t, rows, cols, typ=1):
    b = boxes_array(1)
    b["l"], b["t"], b["r"], b["b"], b["type"], b["score"] = l, t, l + cols - 1, t + rows - 1, typ, 1.0
    return b


@pytest.mark.parametrize("rows,cols", [(128, 128), (64, 64), (64, 128), (130, 67), (35, 34), (35, 66), (66, 33), (129, 34), (34, 131),
                                       (120, 160), (100, 60), (37, 53), (150, 91), (8, 8), (300, 148), (200, 203), (90, 330)])   # first row: every fixed-size fused class (cell sides 8/16/32 in all nine combinations); second row: the any-size kernel -- mixed radices, primes (37 cells), the smallest window, and windows that run in strips (75x37, 50x50, 22x82 cells)
def test_stagewise_vs_oracle(oracle, rows, cols):
    require_gpu()
    M = mot()
    W, H = 640, 480
    sc = Scene(5 + rows + cols, W, H, 3, tsize=40, win=64)
    frame = sc.render()
    ctx = M.Context(W, H, max_tracks=8, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    b = one_box(200, 150, rows, cols)
    h = ctx.new(b)
    ob = box_of(b[0])
    oh = oracle.kcf_new(ob)
    ctx.enable_dumps(True)

    # ---- first update -----------------------------------------------------------------------------------
    ctx.update(h, [0], b)
    g_or = crop_gray(oracle, frame, ob, rows, cols)
    oracle.kcf_update(oh, g_or, ob)
    hr, wc = rows // 4, cols // 4
    gray = ctx.fetch("gray").reshape(cols, rows).T
    assert np.array_equal(gray, g_or), "gray patch must be bit-exact"
    Mo, Oo = oracle.gradmag(g_or)                                    # [x][y]
    m0 = ctx.fetch("m0").reshape(4 * wc, 4 * hr)
    assert np.array_equal(m0, Mo[:4 * wc, :4 * hr] * np.float32(0.0625)), "gradient magnitude must be bit-exact (SSE table emulation)"
    bins = ctx.fetch("bin", np.int32).reshape(4 * wc, 4 * hr)
    assert np.array_equal(bins, oracle_bins(Oo)[:4 * wc, :4 * hr]), "orientation bins must be identical"
    r1 = ctx.fetch("r1").reshape(18, wc, hr)
    assert np.array_equal(r1, oracle.gradhist18(Mo, Oo)), "cell histograms must be bit-exact (ordered gather)"
    feat = ctx.fetch("feat").reshape(31, wc, hr)
    f_or = oracle.kcf_get(oh, "xf_tm").reshape(31, wc, hr)
    assert rel_err(feat, f_or) <= 1e-4, "windowed fHOG features, tolerance 1e-4 relative (BASELINE.json)"
    assert np.array_equal(feat, f_or), "... and in fact bit-exact"
    S = wc * (hr // 2 + 1)
    spec = ctx.fetch("spec").reshape(31, S, 2)
    s_or = oracle.kcf_get(oh, "xf_fq").reshape(31, S, 2)
    assert rel_err(spec, s_or) < 2e-6
    assert rel_err(ctx.fetch("kf"), oracle.kcf_get(oh, "kf").reshape(S, 2)[:, 0]) < 5e-6
    assert rel_err(ctx.state(h[0], "alpha"), oracle.kcf_get(oh, "alpha")) < 2e-5
    assert rel_err(ctx.state(h[0], "xf_md").reshape(31, S, 2), oracle.kcf_get(oh, "xf_md").reshape(31, S, 2)) < 2e-6

    # ---- predict on a shifted scene, then a second (lerp) update ------------------------------------------------
    if hr * wc < 16:
        oracle.kcf_delete(oh); ctx.close()
        return                   # a 2x2-cell window has a degenerate (tied) response; every stage up to the model is compared above
    for step, (dy, dx) in enumerate([(4, -8), (-8, 4)]):
        frame = np.ascontiguousarray(np.roll(frame, (dy, dx), (0, 1)))
        ctx.upload(0, frame)
        out = ctx.predict(h, [0], b)
        g_or = crop_gray(oracle, frame, ob, rows, cols)
        oracle.kcf_predict(oh, g_or, ob)
        resp = ctx.fetch("resp"); r_or = oracle.kcf_get(oh, "response")
        top2 = np.sort(r_or)[-2:]
        margin = (top2[1] - top2[0]) / max(1e-30, abs(top2[1]))
        assert rel_err(resp, r_or) < 2e-5
        assert margin > 1e-3, "synthetic target must give an unambiguous peak (margin %g)" % margin
        assert int(resp.argmax()) == int(r_or.argmax()), "integer peak location must be identical"
        assert tuple(int(out[0][k]) for k in "ltbr") == ob.tup(), "predicted box must be bit-exact"
        b = out.copy()
        ctx.update(h, [0], b)
        g_or = crop_gray(oracle, frame, ob, rows, cols)
        oracle.kcf_update(oh, g_or, ob)
        assert rel_err(ctx.state(h[0], "alpha"), oracle.kcf_get(oh, "alpha")) < 5e-5
        assert rel_err(ctx.state(h[0], "xf_md").reshape(31, S, 2), oracle.kcf_get(oh, "xf_md").reshape(31, S, 2)) < 5e-6
    oracle.kcf_delete(oh)
    ctx.close()


def test_resize_path_matches_reference_scramble(oracle):
    """A crop whose size differs from the template goes through the reference's row/column-mixed resample (drawlib.c:542-637)."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    frame = Scene(77, W, H, 4, tsize=40, win=64).render()
    ctx = M.Context(W, H, max_tracks=4, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    for (rs, cs, rd, cd) in [(100, 90, 128, 128), (64, 64, 64, 64), (150, 141, 64, 128), (40, 33, 32, 32)]:
        b = one_box(100, 80, rs, cs)
        got = ctx.crop_gray_resize(0, b, rd, cd).reshape(cd, rd).T
        want = crop_gray(oracle, frame, box_of(b[0]), rd, cd)
        assert np.array_equal(got, want), (rs, cs, rd, cd)
    ctx.close()


def test_config1_single_target_sequence(oracle):
    """BASELINE config 1 (shortened): 640x480, one seeded 128x128 window, KCF only, update with the own predicted box."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    sc = Scene(0x5EED0100, W, H, 1, tsize=51, win=128, vmax=2.0)
    sc.pos[:] = [[320.0, 240.0]]
    frame = sc.render()
    b = one_box(256, 176, 128, 128)
    ctx = M.Context(W, H, max_tracks=2, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    h = ctx.new(b); ob = box_of(b[0]); oh = oracle.kcf_new(ob)
    ctx.update(h, [0], b); oracle.kcf_update(oh, crop_gray(oracle, frame, ob, 128, 128), ob)
    for f in range(60):
        sc.step(); frame = sc.render(); ctx.upload(0, frame)
        out = ctx.predict(h, [0], b, clamp=1)
        oracle.kcf_predict(oh, crop_gray(oracle, frame, ob, 128, 128), ob)
        ob.l, ob.r = min(max(0, ob.l), W - 1), min(max(0, ob.r), W - 1)
        ob.t, ob.b = min(max(0, ob.t), H - 1), min(max(0, ob.b), H - 1)
        assert tuple(int(out[0][k]) for k in "ltbr") == ob.tup(), "frame %d" % f
        b = out.copy()
        ctx.update(h, [0], b); oracle.kcf_update(oh, crop_gray(oracle, frame, ob, 128, 128), ob)
    # the tracker really follows the target
    assert abs((ob.l + ob.r) / 2 - sc.pos[0, 0]) < 12 and abs((ob.t + ob.b) / 2 - sc.pos[0, 1]) < 12
    oracle.kcf_delete(oh); ctx.close()


def test_mixed_sizes_batch_and_gray_entry(oracle):
    """One batched call over tracks of different window sizes == per-track oracle calls; plus the literal gray-patch plugin form."""
    require_gpu()
    M = mot()
    W, H = 1280, 720
    sc = Scene(31, W, H, 12, tsize=40, win=64)
    frame = sc.render()
    sizes = [(128, 128), (64, 64), (64, 128), (128, 64), (32, 32), (128, 128), (64, 64), (35, 130), (120, 100), (51, 77), (120, 100)]
    b = boxes_array(len(sizes))
    for i, (r, c) in enumerate(sizes):
        b[i] = one_box(40 + 100 * i, 100 + 30 * (i % 3), r, c, typ=i % 3)[0]
    ctx = M.Context(W, H, max_tracks=32, n_frame_slots=2, kind=M.TRACKER_KCF)
    ctx.upload(1, frame)
    h = ctx.new(b)
    fs = np.ones(len(sizes), np.int32)
    ctx.update(h, fs, b)
    obs = [box_of(x) for x in b]; ohs = [oracle.kcf_new(o) for o in obs]
    for o, oh, (r, c) in zip(obs, ohs, sizes):
        oracle.kcf_update(oh, crop_gray(oracle, frame, o, r, c), o)
    frame2 = np.ascontiguousarray(np.roll(frame, (4, 8), (0, 1)))
    ctx.upload(1, frame2)
    out = ctx.predict(h, fs, b)
    for i, (o, oh, (r, c)) in enumerate(zip(obs, ohs, sizes)):
        oracle.kcf_predict(oh, crop_gray(oracle, frame2, o, r, c), o)
        assert tuple(int(out[i][k]) for k in "ltbr") == o.tup(), i
    # gray-patch entry on track 0: same answer as the fused frame path
    ctx2 = M.Context(W, H, max_tracks=4, n_frame_slots=1, kind=M.TRACKER_KCF)
    h2 = ctx2.new(b[:1])
    o = box_of(b[0]); g = crop_gray(oracle, frame, o, 128, 128)
    ctx2.update_gray(h2[0], g, b[:1])
    g2 = crop_gray(oracle, frame2, o, 128, 128)
    got = ctx2.predict_gray(h2[0], g2, b[:1])
    assert tuple(int(got[0][k]) for k in "ltbr") == tuple(int(out[0][k]) for k in "ltbr")
    for oh in ohs:
        oracle.kcf_delete(oh)
    ctx.close(); ctx2.close()


def test_unsupported_shape_fails_loudly():
    require_gpu()
    M = mot()
    ctx = M.Context(640, 480, max_tracks=4, kind=M.TRACKER_KCF)
    with pytest.raises(M.MotError):
        ctx.new(one_box(10, 10, 7, 40))          # fewer than 2 cells along one side
    with pytest.raises(M.MotError):
        ctx.new(one_box(0, 0, 500, 100))         # taller than the frame
    ctx.close()


def test_config1_mixed_radix_window_sequence(oracle):
    """BASELINE config 1, mixed-radix variant: a 120x160-px window = 30x40 cells, any-size path."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    sc = Scene(0x5EED0101, W, H, 1, tsize=51, win=128, vmax=2.0)
    sc.pos[:] = [[320.0, 240.0]]
    frame = sc.render()
    b = one_box(240, 180, 120, 160)
    ctx = M.Context(W, H, max_tracks=2, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    h = ctx.new(b); ob = box_of(b[0]); oh = oracle.kcf_new(ob)
    ctx.update(h, [0], b); oracle.kcf_update(oh, crop_gray(oracle, frame, ob, 120, 160), ob)
    for f in range(25):
        sc.step(); frame = sc.render(); ctx.upload(0, frame)
        out = ctx.predict(h, [0], b, clamp=1)
        oracle.kcf_predict(oh, crop_gray(oracle, frame, ob, 120, 160), ob)
        assert tuple(int(out[0][k]) for k in "ltbr") == ob.tup(), "frame %d" % f
        b = out.copy()
        ctx.update(h, [0], b); oracle.kcf_update(oh, crop_gray(oracle, frame, ob, 120, 160), ob)
    oracle.kcf_delete(oh); ctx.close()


@pytest.mark.parametrize("rows,cols", [(128, 128), (64, 64), (64, 128), (130, 67), (35, 34), (35, 66), (66, 33), (129, 34), (34, 131),
                                       (120, 160), (100, 60), (37, 53), (300, 148), (200, 203)])
def test_model_state_with_dumps_off(oracle, rows, cols):
    """The PRODUCTION instantiation (stage dumps compiled out / never enabled): xf_md and alpha read back with mot_debug_state
    after six predict + update rounds against the compiled reference's model -- all nine fused classes and any-size windows."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    sc = Scene(900 + rows + cols, W, H, 3, tsize=40, win=64)
    frame = sc.render()
    ctx = M.Context(W, H, max_tracks=8, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, frame)
    b = one_box(180, 140, rows, cols)
    h = ctx.new(b); ob = box_of(b[0]); oh = oracle.kcf_new(ob)
    ctx.update(h, [0], b); oracle.kcf_update(oh, crop_gray(oracle, frame, ob, rows, cols), ob)
    hr, wc = rows // 4, cols // 4
    S = wc * (hr // 2 + 1)
    for step in range(6):
        sc.step(); frame = sc.render(); ctx.upload(0, frame)
        out = ctx.predict(h, [0], b, clamp=1)
        oracle.kcf_predict(oh, crop_gray(oracle, frame, ob, rows, cols), ob)
        ob.l, ob.r = min(max(0, ob.l), W - 1), min(max(0, ob.r), W - 1)
        ob.t, ob.b = min(max(0, ob.t), H - 1), min(max(0, ob.b), H - 1)
        assert tuple(int(out[0][k]) for k in "ltbr") == ob.tup(), "step %d" % step
        b = out.copy()
        ctx.update(h, [0], b); oracle.kcf_update(oh, crop_gray(oracle, frame, ob, rows, cols), ob)
        assert rel_err(ctx.state(h[0], "alpha"), oracle.kcf_get(oh, "alpha")) < 5e-5, step
        assert rel_err(ctx.state(h[0], "xf_md").reshape(31, S, 2), oracle.kcf_get(oh, "xf_md").reshape(31, S, 2)) < 5e-6, step
    oracle.kcf_delete(oh); ctx.close()


def test_track_call_equals_predict_then_update(oracle):
    """mot_track_batch = predict (+ clamp) then update with the predicted box, in one call: same boxes and same model as the two calls,
    over tracks of different window sizes (fixed-size and any-size kernels in one batch)."""
    require_gpu()
    M = mot()
    W, H = 1280, 720
    sc = Scene(55, W, H, 8, tsize=40, win=64)
    sizes = [(128, 128), (64, 64), (100, 60), (51, 77), (120, 160), (64, 128)]
    b = boxes_array(len(sizes))
    for i, (r, c) in enumerate(sizes):
        b[i] = one_box(60 + 190 * i, 120 + 40 * (i % 3), r, c, typ=i % 3)[0]
    ca = M.Context(W, H, max_tracks=16, n_frame_slots=1, kind=M.TRACKER_KCF)
    cb = M.Context(W, H, max_tracks=16, n_frame_slots=1, kind=M.TRACKER_KCF)
    frame = sc.render()
    fs = np.zeros(len(sizes), np.int32)
    ha, hb = None, None
    for ctx in (ca, cb):
        ctx.upload(0, frame)
    ha = ca.new(b); hb = cb.new(b)
    ca.update(ha, fs, b); cb.update(hb, fs, b)
    ba, bb = b.copy(), b.copy()
    for step in range(5):
        sc.step(); frame = sc.render()
        ca.upload(0, frame); cb.upload(0, frame)
        ba = ca.predict(ha, fs, ba, clamp=1); ca.update(ha, fs, ba)
        bb = cb.track(hb, fs, bb, clamp=1)
        assert ba.tobytes() == bb.tobytes(), step
    for i in range(len(sizes)):
        assert np.array_equal(ca.state(ha[i], "xf_md"), cb.state(hb[i], "xf_md")) and np.array_equal(ca.state(ha[i], "alpha"), cb.state(hb[i], "alpha")), i
    ca.close(); cb.close()
