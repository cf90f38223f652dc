"""Pins the C restatement (oracle/port) against the UNMODIFIED compiled reference (oracle/_ref).

The reference ships no tests, golden vectors or fixtures, so the strongest
available anchor is the reference's own code run here on seeded inputs.  CPU only.
"""
import ctypes as C
import numpy as np
import pytest

from synth import BBox, Scene, random_boxes, jittered_detections, BBOX_DTYPE


def textured(rng, h, w):
    base = rng.random((h // 8 + 2, w // 8 + 2)).repeat(8, 0).repeat(8, 1)[:h, :w] * 180
    return (base + rng.random((h, w)) * 20).astype(np.float32)


@pytest.mark.parametrize("h,w", [(128, 128), (64, 64), (100, 60), (120, 160), (37, 53), (8, 8)])
def test_fhog_port_equals_reference(port, ref, h, w):
    rng = np.random.default_rng(h * 1000 + w)
    I = textured(rng, h, w)
    Mr, Or = ref.gradmag(I)
    Mp, Op = port.gradmag(I)
    assert np.array_equal(Mr, Mp), "M must be bit-exact (same rsqrtps/rcpps instructions)"
    assert np.array_equal(Or, Op), "O must be bit-exact (same libm acosf table)"
    assert np.array_equal(ref.gradhist18(Mr, Or), port.gradhist18(Mp, Op))
    Hr, Hp = ref.fhog(I), port.fhog(I)
    assert np.array_equal(Hr, Hp)
    assert not Hr[31].any(), "channel 31 is all zero (libhog/fhog.h:27-31)"


def test_fhog_flat_patch(port, ref):
    I = np.full((32, 32), 77.0, np.float32)
    assert np.array_equal(ref.fhog(I), port.fhog(I))


@pytest.mark.parametrize("rows,cols", [(128, 128), (64, 96), (60, 100)])
def test_kcf_port_vs_reference(port, ref, rows, cols):
    rng = np.random.default_rng(7)
    H, W = 240, 320
    img = textured(rng, H, W)
    box = BBox(80, 60, 60 + rows - 1, 80 + cols - 1, 1, 1.0)

    def patch(im, b):
        return im[b.t:b.b + 1, b.l:b.r + 1]

    hr_, hp_ = ref.kcf_new(box), port.kcf_new(box)
    for name in ("labels", "cos_win"):
        assert np.array_equal(ref.kcf_get(hr_, name), port.kcf_get(hp_, name)), name
    np.testing.assert_allclose(port.kcf_get(hp_, "yf"), ref.kcf_get(hr_, "yf"), rtol=0, atol=1e-6)
    br, bp = BBox.from_buffer_copy(box), BBox.from_buffer_copy(box)
    ref.kcf_update(hr_, patch(img, br), br); port.kcf_update(hp_, patch(img, bp), bp)
    for step, (dy, dx) in enumerate([(4, -8), (-4, 4), (8, 8), (0, -4)]):
        img = np.roll(img, (dy, dx), (0, 1))
        ref.kcf_predict(hr_, patch(img, br), br); port.kcf_predict(hp_, patch(img, bp), bp)
        assert br.tup() == bp.tup(), "predicted box, step %d" % step
        resp_r, resp_p = ref.kcf_get(hr_, "response"), port.kcf_get(hp_, "response")
        assert int(resp_r.argmax()) == int(resp_p.argmax())
        np.testing.assert_allclose(resp_p, resp_r, rtol=0, atol=2e-6 * np.abs(resp_r).max())
        assert np.array_equal(ref.kcf_get(hr_, "xf_tm"), port.kcf_get(hp_, "xf_tm"))
        ref.kcf_update(hr_, patch(img, br), br); port.kcf_update(hp_, patch(img, bp), bp)
        for name in ("alpha", "xf_md", "kf"):
            a, b = ref.kcf_get(hr_, name), port.kcf_get(hp_, name)
            np.testing.assert_allclose(b, a, rtol=0, atol=1e-6 * np.abs(a).max(), err_msg=name)
    ref.kcf_delete(hr_); port.kcf_delete(hp_)


def test_kalman_port_vs_reference(port, ref):
    rng = np.random.default_rng(3)
    b0 = BBox(100, 50, 120, 180, 0, 1.0)
    hr_, hp_ = ref.kal_new(b0), port.kal_new(b0)
    pos = np.array([100.0, 50.0])
    for i in range(300):
        br, bp = BBox(), BBox()
        ref.kal_predict(hr_, br); port.kal_predict(hp_, bp)
        assert br.tup() == bp.tup()
        pos += np.array([1.7, -0.6]) + rng.normal(0, 0.5, 2)
        l, t = int(pos[0]), int(pos[1])
        z = BBox(l, t, t + 70, l + 80, 0, 1.0)
        ref.kal_update(hr_, z); port.kal_update(hp_, z)
        xr, Pr, Kr = ref.kal_state(hr_); xp, Pp, Kp = port.kal_state(hp_)
        np.testing.assert_allclose(xp, xr, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(Pp, Pr, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(Kp, Kr, rtol=1e-11, atol=1e-14)
    ref.kal_delete(hr_); port.kal_delete(hp_)


def _cost(port, trk, det, mode, W):
    T, D = len(trk), len(det)
    nr, nc = (T, D) if T < D else (D, T)
    d = np.zeros(nr * nc, np.float64)
    port.kcf.port_cost_matrix(d.ctypes.data_as(C.c_void_p), trk.ctypes.data_as(C.c_void_p), T,
                              det.ctypes.data_as(C.c_void_p), D, mode, C.c_double(1.0 / W))
    return d.reshape(nc, nr).T      # column-major nr x nc


@pytest.mark.parametrize("n,m,mode", [(64, 64, 0), (64, 64, 1), (58, 64, 1), (64, 50, 0), (128, 128, 1), (1, 5, 0), (7, 1, 1), (200, 160, 0)])
def test_hungarian_port_equals_reference(port, ref, n, m, mode):
    rng = np.random.default_rng(n * 100 + m + mode)
    for rep in range(3):
        trk = random_boxes(rng, n, 1920, 1080)
        det = jittered_detections(rng, trk, 1920, 1080, drop=0.0)[:m] if m <= n else \
            np.concatenate([jittered_detections(rng, trk, 1920, 1080), random_boxes(rng, m - n, 1920, 1080)])
        d = _cost(port, trk, det, mode, 1920)
        ar, cr = ref.assign(d); ap, cp = port.assign(d)
        assert np.array_equal(ar, ap)
        assert cr == cp


def test_hungarian_ties_and_degenerate(port, ref):
    rng = np.random.default_rng(11)
    cases = [np.zeros((6, 6)), np.ones((5, 9)), np.ones((9, 5)),
             rng.integers(0, 3, size=(40, 40)).astype(float), rng.integers(0, 4, size=(30, 45)).astype(float),
             rng.integers(0, 4, size=(45, 30)).astype(float), np.round(rng.random((64, 64)), 1)]
    for d in cases:
        ar, cr = ref.assign(d); ap, cp = port.assign(d)
        assert np.array_equal(ar, ap), d.shape
        assert cr == cp


def test_gray_and_resize_vs_original_at_1280(port, ref):
    """rgb2Gray / bilinearInterpolationGray exactly as shipped (stride 3840) vs the stride-parameterised port."""
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, size=(720, 1280, 3), dtype=np.uint8)
    for (l, t, r, b) in [(10, 20, 137, 147), (500, 300, 563, 395), (0, 0, 99, 59)]:
        rows, cols = b - t + 1, r - l + 1
        g_ref = np.zeros(rows * cols, np.float32); g_port = np.zeros(rows * cols, np.float32)
        ref.draw.rgb2Gray(g_ref.ctypes.data_as(C.c_void_p), frame.ctypes.data_as(C.c_void_p), l, t, r, b)
        port.kcf.port_rgb2gray(g_port.ctypes.data_as(C.c_void_p), frame.ctypes.data_as(C.c_void_p), 3840, l, t, r, b)
        assert np.array_equal(g_ref, g_port)
        for (rd, cd) in [(rows, cols), (128, 128), (rows - 7, cols + 5)]:
            o_ref = np.zeros(rd * cd, np.float32); o_port = np.zeros(rd * cd, np.float32)
            ref.draw.bilinearInterpolationGray(o_ref.ctypes.data_as(C.c_void_p), g_ref.ctypes.data_as(C.c_void_p), rows, cols, rd, cd)
            port.kcf.port_resize_gray(o_port.ctypes.data_as(C.c_void_p), g_port.ctypes.data_as(C.c_void_p), rows, cols, rd, cd)
            assert np.array_equal(o_ref, o_port)
            if (rd, cd) == (rows, cols):
                assert np.array_equal(o_ref, g_ref), "equal sizes are an exact copy"


@pytest.mark.parametrize("tracker", ["kal", "kcf"])
def test_frame_loop_port_vs_reference(port, ref, tracker):
    W, H = 640, 480
    n = 6 if tracker == "kcf" else 24
    sc = Scene(101, W, H, n, tsize=32, win=64)
    a, b = ref.td_new(tracker, W, H, 64, 0), port.td_new(tracker, W, H, 64, 0)
    drng = np.random.default_rng(9)
    for f in range(40 if tracker == "kal" else 12):
        sc.step()
        frame = sc.render()
        dets = sc.windows(jitter=2)
        keep = drng.random(len(dets)) > 0.1
        dets = np.ascontiguousarray(dets[keep])
        a.step(frame, dets); b.step(frame, dets)
        ta, tb = a.tracks(), b.tracks()
        for k in ta:
            assert np.array_equal(ta[k], tb[k]), (f, k)
        pa, aa = a.last(); pb, ab = b.last()
        assert np.array_equal(pa, pb) and np.array_equal(aa, ab), f
    a.close(); b.close()
