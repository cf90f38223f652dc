"""Golden vectors: outputs of the UNMODIFIED reference (oracle/_ref, run in the authoring container by
tests/golden/make_golden.py) committed as fixtures.  CPU: the C restatement must reproduce them.  GPU: so must the
CUDA path through the C ABI.  These are the only checks that still pin parity when neither /root/reference nor the
prebuilt oracle/_ref is around."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402
from synth import BBox, Scene, boxes_array, BBOX_DTYPE  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(HERE, "golden", "golden_v1.npz")))


@pytest.fixture(scope="module")
def inputs():
    return G.golden_inputs()


EXACT_PREFIXES = ("fhog_", "gray_", "resize_", "kcf_boxes", "kcf_peaks", "kcf_feat0", "kal_boxes", "assoc_", "ties_", "td_")


def test_port_reproduces_reference_golden(port, golden, inputs):
    got = G.compute(port, inputs)
    assert set(got) == set(golden)
    for k in sorted(golden):
        if k.startswith(EXACT_PREFIXES):
            assert np.array_equal(got[k], golden[k]), k
        elif k == "kal_state":
            np.testing.assert_allclose(got[k], golden[k], rtol=1e-10, atol=1e-10, err_msg=k)
        else:
            np.testing.assert_allclose(got[k], golden[k], rtol=0, atol=2e-6 * np.abs(golden[k]).max(), err_msg=k)


def test_reference_still_reproduces_golden(ref, golden, inputs):
    """Guards the fixtures themselves (runs only where oracle/_ref exists)."""
    got = G.compute(ref, inputs)
    for k in sorted(golden):
        if k.startswith(EXACT_PREFIXES) or k == "kal_state":
            assert np.array_equal(got[k], golden[k]), k


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_golden(golden, inputs):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mot_b200 as M
    g = inputs

    # ---- crop + gray + resize (top/drawlib.c:192-240, 542-637): bit-exact -------------------------------------
    ctx = M.Context(320, 240, max_tracks=8, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.upload(0, g["frame"])
    for i, ((l, t, r, b), (rd, cd)) in enumerate(g["crops"]):
        bx = boxes_array(1); bx["l"], bx["t"], bx["r"], bx["b"] = l, t, r, b
        assert np.array_equal(ctx.crop_gray_resize(0, bx, rd, cd), golden["resize_%d" % i]), i

    # ---- KCF through the literal gray-patch plugin form: features bit-exact, boxes and peaks bit-exact ----------
    l, t, rows, cols = g["kcf_box"]
    img = g["kcf_img"].copy()
    bx = boxes_array(1); bx["l"], bx["t"], bx["r"], bx["b"], bx["type"], bx["score"] = l, t, l + cols - 1, t + rows - 1, 1, 1.0
    h = ctx.new(bx)[0]
    patch = lambda im, b_: im[int(b_["t"][0]):int(b_["b"][0]) + 1, int(b_["l"][0]):int(b_["r"][0]) + 1]
    ctx.enable_dumps(True)
    ctx.update_gray(h, patch(img, bx), bx)
    assert np.array_equal(ctx.fetch("feat"), golden["kcf_feat0"]), "windowed fHOG features vs the reference: bit-exact"
    np.testing.assert_allclose(ctx.state(h, "alpha"), golden["kcf_alpha0"], rtol=0, atol=2e-5 * np.abs(golden["kcf_alpha0"]).max())
    boxes, peaks = [], []
    for (dy, dx) in g["kcf_shifts"]:
        img = np.roll(img, (dy, dx), (0, 1))
        bx = ctx.predict_gray(h, patch(img, bx), bx)
        boxes.append(tuple(int(bx[k][0]) for k in "ltbr")); peaks.append(int(ctx.fetch("resp").argmax()))
        ctx.update_gray(h, patch(img, bx), bx)
    assert np.array_equal(np.array(boxes, np.int32), golden["kcf_boxes"])
    assert np.array_equal(np.array(peaks, np.int32), golden["kcf_peaks"])
    np.testing.assert_allclose(ctx.state(h, "alpha"), golden["kcf_alpha_end"], rtol=0, atol=1e-4 * np.abs(golden["kcf_alpha_end"]).max())
    ctx.close()

    # ---- Kalman: int boxes bit-exact, state <= 1e-5 relative ------------------------------------------------------
    ctx = M.Context(1920, 1080, max_tracks=8, kind=M.TRACKER_KALMAN)
    b0 = boxes_array(1); b0["l"], b0["t"], b0["b"], b0["r"], b0["score"] = 100, 50, 120, 180, 1.0
    h = ctx.new(b0)
    rng = np.random.default_rng(3); pos = np.array([100.0, 50.0])
    for i in range(len(golden["kal_boxes"])):
        out = ctx.predict(h, None, b0)
        assert tuple(int(out[k][0]) for k in "ltbr") == tuple(golden["kal_boxes"][i]), i
        pos += np.array([1.7, -0.6]) + rng.normal(0, 0.5, 2)
        l_, t_ = int(pos[0]), int(pos[1])
        z = boxes_array(1); z["l"], z["t"], z["b"], z["r"], z["score"] = l_, t_, t_ + 70, l_ + 80, 1.0
        ctx.update(h, None, z)
        st = np.concatenate([ctx.state(h[0], "x"), ctx.state(h[0], "P").ravel()])
        ref_st = golden["kal_state"][i]
        assert np.abs(st - ref_st).max() <= 1e-5 * np.abs(ref_st).max(), i

    # ---- association: cost matrices and assignments bit-exact, ties included ----------------------------------------
    trks = [a[0] for a in g["assoc"]]; dets = [a[1] for a in g["assoc"]]
    for mode in (0, 1):
        idx = [i for i, a in enumerate(g["assoc"]) if a[2] == mode]
        assigns, costs, dists = ctx.associate([trks[i] for i in idx], [dets[i] for i in idx], cost_mode=mode, want_dist=True)
        for j, i in enumerate(idx):
            assert np.array_equal(dists[j], golden["assoc_dist_%d" % i]), i
            assert np.array_equal(assigns[j], golden["assoc_assign_%d" % i]), i
            assert costs[j] == golden["assoc_cost_%d" % i][0], i
    assigns, costs = ctx.assign(g["ties"])
    for i in range(len(g["ties"])):
        assert np.array_equal(assigns[i], golden["ties_assign_%d" % i]), i
        assert costs[i] == golden["ties_cost_%d" % i][0], i
    ctx.close()

    # ---- frame loops -----------------------------------------------------------------------------------------------------
    class Gpu:
        def td_new(self, tracker, W, H, cap, mode):
            c = M.Context(W, H, max_tracks=128, n_frame_slots=1, kind=M.TRACKER_KCF if tracker == "kcf" else M.TRACKER_KALMAN)
            td = c.td(0, cap=cap, cost_mode=mode); td._ctx = c
            return td
    assert np.array_equal(G.td_trace(Gpu(), "kal", 40), golden["td_kal"])
    assert np.array_equal(G.td_trace(Gpu(), "kcf", 10), golden["td_kcf"])
