"""Oracle for the YOLO3 post-processing that feeds the tracker (detectors/yolo3.cpp:141-356, 487-527).

Round-1 state of this row: the oracle only.  The reference's four functions are compiled from their own text into
oracle/_ref/libref_yolo.so (oracle/tools/extract_yolo.py + oracle/capi/yolo_capi.cpp) and the C restatement
(oracle/port/port_yolo.c) is pinned against them bit for bit -- candidates, truncated corners, the exchange sort, the
per-class NMS with its never-cleared suppression flags, clipping and output order.  No CUDA path exists for it yet."""
import ctypes as C

import numpy as np
import pytest

import oraclelib
from synth import BBOX_DTYPE

ANCHORS = np.array([55, 69, 75, 234, 133, 240, 136, 129, 142, 363, 203, 290, 228, 184, 285, 359, 341, 260], np.int32)   # yolo3.cpp:557


def synth_outputs(rng, th, tw, nc, n_obj, dup=3):
    """Three output maps (grid/32, /16, /8) of background logits with n_obj objects injected, each with a few near-duplicates
    (neighbouring anchors / cells, a second class now and then) so that NMS, the sort and the class loop all have work."""
    per = 5 + nc
    outs = []
    for k in range(3):
        gh, gw = (th // 32) << k, (tw // 32) << k
        o = rng.normal(0.0, 1.0, (gh, gw, 3, per)).astype(np.float32)
        o[..., 4] -= 7.0                                   # objectness far below any threshold
        o[..., 5:] -= 3.0
        outs.append(o)
    for _ in range(n_obj):
        k = int(rng.integers(0, 3)); o = outs[k]
        gh, gw = o.shape[:2]
        r, c = int(rng.integers(0, gh)), int(rng.integers(0, gw))
        cls = int(rng.integers(0, nc))
        for d in range(int(rng.integers(1, dup + 1))):
            rr = min(gh - 1, max(0, r + int(rng.integers(-1, 2)))) if d else r
            cc = min(gw - 1, max(0, c + int(rng.integers(-1, 2)))) if d else c
            a = int(rng.integers(0, 3))
            o[rr, cc, a, 0:2] = rng.normal(0, 1, 2)
            o[rr, cc, a, 2:4] = rng.normal(0, 0.4, 2)
            o[rr, cc, a, 4] = rng.uniform(1.0, 7.0)
            o[rr, cc, a, 5 + cls] = rng.uniform(0.5, 6.0)
            if nc > 1 and rng.random() < 0.3:
                o[rr, cc, a, 5 + int(rng.integers(0, nc))] = rng.uniform(0.5, 6.0)
    return [np.ascontiguousarray(o.reshape(-1)) for o in outs]


def run(lib, name, outs, obj, nms, th, tw, ih, iw, nc, max_out=4096):
    out = np.zeros(max_out, BBOX_DTYPE)
    f = getattr(lib, name)
    f.restype = C.c_int
    n = f(outs[0].ctypes.data_as(C.c_void_p), outs[1].ctypes.data_as(C.c_void_p), outs[2].ctypes.data_as(C.c_void_p),
          ANCHORS.ctypes.data_as(C.c_void_p), C.c_float(obj), C.c_float(nms), th, tw, ih, iw, nc, out.ctypes.data_as(C.c_void_p), max_out)
    return out[:n]


@pytest.mark.skipif(not __import__("os").path.exists(__import__("os").path.join(oraclelib.ODIR, "_ref", "libref_yolo.so")), reason="oracle/_ref/libref_yolo.so not built")
@pytest.mark.parametrize("th,tw,ih,iw,nc,nobj,obj,nms", [
    (416, 416, 720, 1280, 80, 60, 0.5, 0.45), (480, 480, 480, 640, 1, 40, 0.5, 0.45), (416, 416, 1080, 1920, 3, 120, 0.3, 0.3),
    (320, 608, 720, 1280, 20, 80, 0.25, 0.6), (416, 416, 1280, 720, 5, 200, 0.1, 0.45), (416, 416, 720, 1280, 2, 1, 0.5, 0.45)])
def test_port_equals_compiled_reference(th, tw, ih, iw, nc, nobj, obj, nms):
    import os
    ref = C.CDLL(os.path.join(oraclelib.ODIR, "_ref", "libref_yolo.so"))
    port = oraclelib.Oracle("port").kcf
    rng = np.random.default_rng(th * 7 + iw + nc)
    total = 0
    for rep in range(4):
        outs = synth_outputs(rng, th, tw, nc, nobj)
        a = run(ref, "ref_yolo_post", outs, obj, nms, th, tw, ih, iw, nc)
        b = run(port, "port_yolo_post", outs, obj, nms, th, tw, ih, iw, nc)
        assert len(a) == len(b), (rep, len(a), len(b))
        assert a.tobytes() == b.tobytes(), rep
        total += len(a)
    assert total > 0, "the synthetic maps must produce detections"


@pytest.mark.skipif(not __import__("os").path.exists(__import__("os").path.join(oraclelib.ODIR, "_ref", "libref_yolo.so")), reason="oracle/_ref/libref_yolo.so not built")
def test_port_equals_compiled_reference_on_overflowing_sizes():
    """Raw t_w / t_h large enough that expf overflows: the corner conversions are out of the int range, where the x86 (int) of the
    reference build yields INT_MIN.  The restatement (same compiler, same conversion) must agree with the compiled reference."""
    import os
    ref = C.CDLL(os.path.join(oraclelib.ODIR, "_ref", "libref_yolo.so"))
    port = oraclelib.Oracle("port").kcf
    rng = np.random.default_rng(99)
    th, tw, ih, iw, nc = 416, 416, 720, 1280, 3
    outs = synth_outputs(rng, th, tw, nc, 30)
    o = outs[1].reshape((th // 32) * 2, (tw // 32) * 2, 3, 5 + nc)
    for (r_, c_, a_, tw_, th_) in ((1, 1, 0, 95.0, 0.1), (2, 3, 1, 0.2, 120.0), (3, 2, 2, 60.0, 60.0), (4, 4, 0, 30.0, -3.0)):
        o[r_, c_, a_, 2] = tw_; o[r_, c_, a_, 3] = th_; o[r_, c_, a_, 4] = 6.0; o[r_, c_, a_, 5] = 5.0
    a = run(ref, "ref_yolo_post", outs, 0.5, 0.45, th, tw, ih, iw, nc)
    b = run(port, "port_yolo_post", outs, 0.5, 0.45, th, tw, ih, iw, nc)
    assert a.tobytes() == b.tobytes() and len(a) > 0


def test_no_candidates_gives_no_detections():
    port = oraclelib.Oracle("port").kcf
    outs = [np.full(((416 // 32) << k) ** 2 * 3 * 85, -20.0, np.float32) for k in range(3)]
    assert len(run(port, "port_yolo_post", outs, 0.5, 0.45, 416, 416, 720, 1280, 80)) == 0


def test_port_equals_golden_fixture():
    """The same pin where neither /root/reference nor oracle/_ref exists: detections of the compiled reference, committed."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yolo_v1.npz"))
    port = oraclelib.Oracle("port").kcf
    for tag in ("a", "b"):
        th, tw, ih, iw, nc = [int(v) for v in g[tag + "_cfg"]]
        obj, nms = [float(v) for v in g[tag + "_thr"]]
        outs = [np.ascontiguousarray(g["%s_out%d" % (tag, k)]) for k in range(3)]
        det = run(port, "port_yolo_post", outs, obj, nms, th, tw, ih, iw, nc)
        want = g[tag + "_det"]
        assert len(det) == len(want)
        for k in ("l", "t", "b", "r", "type"):
            assert np.array_equal(det[k], want[k]), k
        # the scores go through expf: identical with the same libm, else equal to the last bits
        assert np.allclose(det["score"], want["score"], rtol=2e-6, atol=0)


def test_expf_recipe_reproduces_the_c_library():
    """What a bit-exact GPU version of the decode step needs: the C library's expf restated in plain double operations
    (oracle/port/port_expf.c) equals it on every 97th float bit pattern with |x| <= 87 (23 million arguments)."""
    port = oraclelib.Oracle("port").kcf
    port.port_expf_mismatches.restype = C.c_long
    tested = C.c_long(0)
    bad = port.port_expf_mismatches(C.c_uint32(97), C.byref(tested))
    assert tested.value > 20_000_000
    assert bad == 0, "%d of %d differ: this C library evaluates expf differently (glibc < 2.27?)" % (bad, tested.value)


@pytest.mark.gpu
@pytest.mark.parametrize("th,tw,ih,iw,nc,nobj,obj,nms", [
    (416, 416, 720, 1280, 80, 60, 0.5, 0.45), (480, 480, 480, 640, 1, 40, 0.5, 0.45), (416, 416, 1080, 1920, 3, 120, 0.3, 0.3),
    (320, 608, 720, 1280, 20, 80, 0.25, 0.6), (416, 416, 1280, 720, 5, 200, 0.1, 0.45)])
def test_gpu_yolo_post_equals_oracle(th, tw, ih, iw, nc, nobj, obj, nms):
    """mot_yolo_post (csrc/yolo_post.cu) against the oracle: same boxes, classes, order and score BITS (expf reproduced in FP64)."""
    from gpu_common import require_gpu, mot
    require_gpu()
    M = mot()
    ctx = M.Context(iw, ih, max_tracks=4, kind=M.TRACKER_KALMAN)
    import os
    # the strongest oracle present: the reference's own functions compiled from their text (oracle/_ref/libref_yolo.so travels to
    # the GPU box), else the C restatement that a CPU test pins to them bit for bit
    ref_so = os.path.join(oraclelib.ODIR, "_ref", "libref_yolo.so")
    port, fn = (C.CDLL(ref_so), "ref_yolo_post") if os.path.exists(ref_so) else (oraclelib.Oracle("port").kcf, "port_yolo_post")
    rng = np.random.default_rng(th * 7 + iw + nc)
    total = 0
    for rep in range(4):
        outs = synth_outputs(rng, th, tw, nc, nobj)
        if rep == 3:
            # extreme raw sizes: expf overflows to +inf (and inf * 0 style NaNs downstream), so the corner conversions leave the
            # int range -- the reference's x86 (int) gives INT_MIN there, which decides whether the box survives the clipping
            o = outs[1].reshape((th // 32) * 2, (tw // 32) * 2, 3, 5 + nc)
            for (r_, c_, a_, tw_, th_) in ((1, 1, 0, 95.0, 0.1), (2, 3, 1, 0.2, 120.0), (3, 2, 2, 60.0, 60.0), (4, 4, 0, 30.0, -3.0)):
                o[r_, c_, a_, 2] = tw_; o[r_, c_, a_, 3] = th_; o[r_, c_, a_, 4] = 6.0; o[r_, c_, a_, 5] = 5.0
        if rep == 2:                                     # exact score ties: the literal exchange sort path
            o = outs[0].reshape(th // 32, tw // 32, 3, 5 + nc)
            o[1, 1, 0, :] = o[2, 2, 1, :] = o[3, 1, 2, :]
            o[1, 1, 0, 4] = o[2, 2, 1, 4] = o[3, 1, 2, 4] = 5.0
            o[1, 1, 0, 5] = o[2, 2, 1, 5] = o[3, 1, 2, 5] = 4.0
        want = run(port, fn, outs, obj, nms, th, tw, ih, iw, nc)
        got = ctx.yolo_post(outs, ANCHORS, obj, nms, th, tw, ih, iw, nc)
        assert len(got) == len(want), (rep, len(got), len(want))
        assert got.tobytes() == want.tobytes(), rep
        total += len(got)
    assert total > 0
    ctx.close()
