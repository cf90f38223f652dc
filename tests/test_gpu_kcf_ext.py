"""GPU: the north-star extensions of the correlation filter (Gaussian kernel, sub-pixel peak, padding, target-sized labels) against the
NumPy restatement tests/kcf_ext_numpy.py.  The reference has none of them: parity is UNPINNED by the reference, the restatement is the
specification.  Tolerances: response-derived sub-pixel position within 0.01 px (north-star), alpha / model within 1e-4 relative, integer
boxes equal unless the restatement's untruncated coordinate lies within 0.01 px of an integer."""
import numpy as np
import pytest

from synth import Scene, boxes_array
from gpu_common import require_gpu, mot, rel_err, box_of, crop_gray
from kcf_ext_numpy import KcfExt, pad_box, unpad_box

pytestmark = pytest.mark.gpu


def window_fhog(oracle, frame, wbox, rows, cols):
    from synth import BBox
    g = crop_gray(oracle, frame, BBox(wbox[0], wbox[1], wbox[2], wbox[3], 0, 0.0), rows, cols)
    return oracle.fhog(np.ascontiguousarray(g))


@pytest.mark.parametrize("name,opts,size", [
    ("gaussian", dict(gaussian=True, sigma=0.5), (96, 80)),
    ("gaussian+subpixel", dict(gaussian=True, sigma=0.5, subpixel=True), (128, 128)),
    ("subpixel (linear kernel)", dict(subpixel=True), (100, 60)),
    ("padding 2.5 + gaussian + subpixel + target-sized label (north-star config 1)", dict(gaussian=True, sigma=0.5, subpixel=True, padding=2.5, osf=0.1), (51, 51)),
    ("padding 1.5 (linear kernel)", dict(padding=1.5), (64, 48)),
])
def test_extensions_vs_numpy_restatement(oracle, name, opts, size):
    require_gpu()
    M = mot()
    W, H = 640, 480
    th, tw = size
    sc = Scene(0xE0 + th + tw, W, H, 1, tsize=max(24, min(th, tw) * 5 // 8), win=max(th, tw), vmax=2.5)
    sc.pos[:] = [[320.0, 240.0]]
    sc.vel[:] = [[3.1, -2.3]]                       # px / frame: the window moves by whole cells every other frame, by fractions with sub-pixel mode
    frame = sc.render()
    ctx = M.Context(W, H, max_tracks=4, n_frame_slots=1, kind=M.TRACKER_KCF)
    ctx.set_kcf_options(**opts)
    pad = opts.get("padding", 0.0)
    b = boxes_array(1)
    l0, t0 = 320 - tw // 2, 240 - th // 2
    b["l"], b["t"], b["r"], b["b"], b["type"], b["score"] = l0, t0, l0 + tw - 1, t0 + th - 1, 1, 1.0
    target = (int(b[0]["l"]), int(b[0]["t"]), int(b[0]["b"]), int(b[0]["r"]))
    window = pad_box(*target, pad) if pad > 1 else target
    rows, cols = window[2] - window[1] + 1, window[3] - window[0] + 1
    ref = KcfExt(rows, cols, gaussian=opts.get("gaussian", False), sigma=opts.get("sigma", 0.5), subpixel=opts.get("subpixel", False), osf=opts.get("osf", 0.0))
    ctx.upload(0, frame)
    h = ctx.new(b)
    ctx.update(h, [0], b)
    ref.update(window_fhog(oracle, frame, window, rows, cols), window, target)
    S = ref.wc * (ref.hr // 2 + 1)

    def check_state(tag):
        a_re = ctx.state(h[0], "alpha")[:S]
        want = ref.alphaf[:, : ref.hr // 2 + 1].reshape(-1)
        # alpha = yf / (kf + lambda): with the Gaussian kernel on HOG features kf is the transform of a nearly flat kernel, small next
        # to its DC bin, so float32 alpha is conditioned ~10x worse than the linear filter's
        tol_a = 1e-3 if opts.get("gaussian") else 1e-4
        assert rel_err(a_re, np.real(want)) < tol_a, (name, tag, "Re alpha", rel_err(a_re, np.real(want)))
        if opts.get("gaussian"):
            a_im = ctx.state(h[0], "alpha_im")[:S]
            assert np.abs(a_im - np.imag(want)).max() <= tol_a * np.abs(want).max(), (name, tag, "Im alpha")
        md = ctx.state(h[0], "xf_md").reshape(31, S, 2)
        wm = ref.model[:, :, : ref.hr // 2 + 1].reshape(31, S)
        assert rel_err(md[..., 0], np.real(wm)) < 1e-4 and np.abs(md[..., 1] - np.imag(wm)).max() <= 1e-4 * np.abs(wm).max(), (name, tag, "model")

    check_state("first update")
    moved = 0
    for f in range(8):
        sc.step(); frame = sc.render(); ctx.upload(0, frame)
        prev_window = ref.pos
        out = ctx.predict(h, [0], b, clamp=0)
        newpos, (sdv, sdh), resp, fl = ref.predict(window_fhog(oracle, frame, prev_window, rows, cols))
        top2 = np.sort(resp.ravel())[-2:]
        assert (top2[1] - top2[0]) / abs(top2[1]) > 1e-4, "ambiguous peak in the synthetic scene"
        if opts.get("subpixel"):
            got = ctx.state(h[0], "subpixel")
            assert abs(got[0] - sdv) * 4 <= 0.01 and abs(got[1] - sdh) * 4 <= 0.01, (name, f, got, (sdv, sdh))
        want_t = unpad_box(*newpos, ref.tw, ref.th) if pad > 1 else newpos
        got_t = (int(out[0]["l"]), int(out[0]["t"]), int(out[0]["b"]), int(out[0]["r"]))
        if got_t != want_t:
            # only a coordinate whose untruncated value is within 0.01 px of an integer may land on the other side
            near = [abs(v - round(v)) <= 0.01 for v in fl]
            assert any(near) and max(abs(g - w) for g, w in zip(got_t, want_t)) <= 1, (name, f, got_t, want_t, fl)
        moved += got_t != target
        # both sides continue from the CUDA path's box (identical crops from here on)
        target = got_t
        window = pad_box(*target, pad) if pad > 1 else target
        ref.pos = window
        b = out.copy()
        ctx.update(h, [0], b)
        ref.update(window_fhog(oracle, frame, window, rows, cols), window, target)
        if f in (0, 7):
            check_state("frame %d" % f)
    assert moved >= 3, "the target must actually move in this test"
    # the tracker follows the target
    cx, cy = (target[0] + target[3]) / 2, (target[1] + target[2]) / 2
    assert abs(cx - sc.pos[0, 0]) < 10 and abs(cy - sc.pos[0, 1]) < 10, (name, (cx, cy), sc.pos[0])
    ctx.close()


def test_extension_defaults_are_the_reference_filter(oracle):
    """All options zero -> nothing changes: same boxes as a context that never heard of the options."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    sc = Scene(7, W, H, 2, tsize=40, win=64)
    frame = sc.render()
    b = sc.windows()
    ca = M.Context(W, H, max_tracks=4, n_frame_slots=1, kind=M.TRACKER_KCF)
    cb = M.Context(W, H, max_tracks=4, n_frame_slots=1, kind=M.TRACKER_KCF)
    cb.set_kcf_options()
    fs = np.zeros(len(b), np.int32)
    for c in (ca, cb):
        c.upload(0, frame)
    ha, hb = ca.new(b), cb.new(b)
    ca.update(ha, fs, b); cb.update(hb, fs, b)
    sc.step(); frame = sc.render()
    for c in (ca, cb):
        c.upload(0, frame)
    assert ca.predict(ha, fs, b).tobytes() == cb.predict(hb, fs, b).tobytes()
    ca.close(); cb.close()
