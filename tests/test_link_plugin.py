"""The drop-in boundary really links: a client written against the REFERENCE's declarations (top/cnntype.h:36-47 and the
prototypes of top/td.cpp:229-261) is compiled and linked against libmot_b200.so with -Wl,--no-undefined, and -- on the GPU --
run through the same call sequence the tracking thread makes, against the compiled reference plugin."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "link", "plugin_client.cpp")
REF_HDR = "/root/reference/top/cnntype.h"
EXPECTED = [l.strip() for l in open(os.path.join(HERE, "link", "expected_symbols.txt")) if l.strip() and not l.startswith("#")]


def build_client(out, with_reference_header):
    import mot_b200
    mot_b200.build()
    libdir = os.path.dirname(mot_b200.LIB_PATH)
    cmd = ["g++", "-O1", "-std=c++14", "-fPIC", "-shared", SRC, "-o", out, "-Wl,--no-undefined", "-L" + libdir, "-lmot_b200", "-Wl,-rpath," + libdir]
    if with_reference_header:
        cmd.insert(1, "-DREF_CNNTYPE_H=\"%s\"" % REF_HDR)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "client does not link against libmot_b200.so:\n" + r.stderr[-3000:]
    return out


def undefined_plugin_symbols(so):
    out = subprocess.run(["nm", "-D", "--undefined-only", so], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if l.strip()]
    return sorted(s for s in syms if "tracker_" in s or "assignmentoptimal" in s or s in ("rgb2Gray", "bilinearInterpolationGray"))


def test_client_written_against_the_reference_links(tmp_path):
    so = build_client(str(tmp_path / "libclient.so"), with_reference_header=False)
    assert undefined_plugin_symbols(so) == sorted(EXPECTED)
    import mot_b200
    defined = subprocess.run(["nm", "-D", "--defined-only", mot_b200.LIB_PATH], capture_output=True, text=True).stdout
    have = {l.split()[-1] for l in defined.splitlines() if l.strip()}
    for s in EXPECTED:
        assert s in have, "libmot_b200.so does not define %s" % s


@pytest.mark.skipif(not os.path.exists(REF_HDR), reason="/root/reference absent")
def test_expected_symbols_come_from_the_reference_header(tmp_path):
    """Same client, but bbox_t comes from the reference's own top/cnntype.h: the names it needs are the committed ones."""
    so = build_client(str(tmp_path / "libclient_ref.so"), with_reference_header=True)
    assert undefined_plugin_symbols(so) == sorted(EXPECTED)
    # and the prototypes in the client are the reference's, token for token (top/td.cpp:229-234)
    ref = open("/root/reference/top/td.cpp", encoding="latin1").read().split("\n")[228:234]
    mine = open(SRC).read()
    for line in ref:
        if line.strip():
            assert line.strip() in mine, line


@pytest.mark.gpu
def test_client_sequence_matches_the_reference_plugin(tmp_path, oracle):
    """The calls of top/td.cpp for one KCF track (spawn, first update, then crop / resize / predict / clamp / update per frame),
    made by the client through the reference's own symbols into libmot_b200.so, against the oracle's plugin fed by the
    oracle's crop helpers: every box equal.  Frames are the reference's 1280x720 (the shim's default configuration)."""
    from gpu_common import require_gpu, crop_gray
    from synth import Scene, BBox
    require_gpu()
    os.environ["MOT_TRACKER"] = "kcf"
    so = build_client(str(tmp_path / "libclient.so"), with_reference_header=False)
    L = C.CDLL(so)
    assert L.client_sizes() == 24 * 1000 + (4 + 128 * 24) % 1000
    W, H, NF = 1280, 720, 6
    sc = Scene(21, W, H, 3, tsize=44, win=96)
    frames = []
    for k in range(NF):
        frames.append(sc.render()); sc.step()
    fr = np.ascontiguousarray(np.stack(frames))
    for (l, t, w, h) in ((300, 200, 128, 128), (1180, 600, 90, 110), (40, 30, 61, 77)):      # pow-2, near the border (clamp -> resize path), odd size
        first = BBox(l, t, t + h - 1, l + w - 1, 1, 0.5)
        out = (BBox * NF)()
        assert L.client_kcf_sequence(fr.ctypes.data_as(C.c_void_p), NF, C.byref(first), out) == 0
        # the oracle's side of the same sequence
        ob = BBox(first.l, first.t, first.b, first.r, 1, 0.5)
        oh = oracle.kcf_new(ob)
        rows, cols = h, w
        oracle.kcf_update(oh, crop_gray(oracle, fr[0], ob, rows, cols), ob)
        for k in range(1, NF):
            oracle.kcf_predict(oh, crop_gray(oracle, fr[k], ob, rows, cols), ob)
            ob.l = min(max(0, ob.l), W - 1); ob.r = min(max(0, ob.r), W - 1); ob.t = min(max(0, ob.t), H - 1); ob.b = min(max(0, ob.b), H - 1)
            oracle.kcf_update(oh, crop_gray(oracle, fr[k], ob, rows, cols), ob)
            assert (out[k].l, out[k].t, out[k].b, out[k].r) == (ob.l, ob.t, ob.b, ob.r), (k, (l, t, w, h))
        oracle.kcf_delete(oh)
    # assignmentoptimal through the same door
    rng = np.random.default_rng(4)
    d = np.asfortranarray(rng.integers(0, 20, (24, 31)).astype(np.float64))
    a = np.full(24, -2, np.int32); cost = C.c_double(0)
    L.client_assign(a.ctypes.data_as(C.c_void_p), C.byref(cost), d.ctypes.data_as(C.c_void_p), 24, 31)
    a_or, c_or = oracle.assign(d)
    assert np.array_equal(a, a_or) and cost.value == c_or
