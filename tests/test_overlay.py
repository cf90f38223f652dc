"""Overlay step (top/td.cpp:647-733 + drawRect, top/drawlib.c:97-151).

CPU part: the restatement (oracle/port/port_overlay.c) against the compiled original at its hard-coded 1280-pixel stride,
the palette and hash against the fixture extracted from the reference (tests/golden/overlay_v1.npz), and the library's
host-side colour helper against both.  GPU part: the device kernel through the C ABI, byte for byte against the oracle,
on overlapping, reversed, degenerate and border-touching boxes, several frames per call."""
import ctypes as C
import os

import numpy as np
import pytest

import oraclelib
from synth import boxes_array
from gpu_common import require_gpu, mot

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "overlay_v1.npz"))


def port():
    return oraclelib.Oracle("port").kcf


def overlay_boxes(rng, n, W, H, kind="mixed"):
    b = boxes_array(n)
    l = rng.integers(0, W, n); t = rng.integers(0, H, n)
    w = rng.integers(0, 300, n); h = rng.integers(0, 300, n)
    b["l"], b["t"] = l, t
    b["r"], b["b"] = np.minimum(l + w, W - 1), np.minimum(t + h, H - 1)        # the loop clamps boxes to the frame (td.cpp:378-381)
    if kind == "mixed":
        rev = rng.random(n) < 0.15                                             # reversed corners: drawRect swaps them
        b["l"][rev], b["r"][rev] = b["r"][rev].copy(), b["l"][rev].copy()
        tiny = rng.random(n) < 0.15                                            # thinner than the three nested rectangles
        b["r"][tiny] = b["l"][tiny] + rng.integers(0, 4, int(tiny.sum()))
        b["b"][tiny] = b["t"][tiny] + rng.integers(0, 4, int(tiny.sum()))
    return b


def port_overlay(frame, boxes, rgb, thickness=3):
    out = frame.copy()
    L = port()
    L.port_overlay(out.ctypes.data_as(C.c_void_p), C.c_int(out.strides[0]), C.c_long(out.nbytes), C.c_int(len(boxes)),
                   boxes.ctypes.data_as(C.c_void_p), np.ascontiguousarray(rgb, np.uint32).ctypes.data_as(C.c_void_p), C.c_int(thickness))
    return out


def test_palette_and_hash_match_reference_fixture():
    L = port()
    L.port_colormap.restype = C.c_uint32; L.port_hashcolor.restype = C.c_uint32; L.port_track_color.restype = C.c_uint32
    L.port_hashcolor.argtypes = [C.c_uint32]; L.port_track_color.argtypes = [C.c_uint32]
    cmap = np.array([L.port_colormap(i) for i in range(256)], np.uint32)
    assert np.array_equal(cmap, GOLD["colormap"])
    tids, hashes = GOLD["tids"], GOLD["hashes"]
    assert all(L.port_hashcolor(int(t)) == int(h) for t, h in zip(tids, hashes))
    import mot_b200
    # the colour of a track is fixed by the reference's spawn statement (top/td.cpp:619-620), which hashes the id counter
    # AFTER `tid = tracker_id++`: the fixture holds (tid, colour index) pairs produced by those two lines themselves
    sp_tid, sp_col = GOLD["spawn_tid"], GOLD["spawn_color"]
    assert len(sp_tid) >= 4096 and int(sp_tid[0]) == 0
    for t, ci in zip(sp_tid[::5], sp_col[::5]):
        want = int(GOLD["colormap"][int(ci)])
        assert L.port_track_color(int(t)) == want
        assert mot_b200.track_color(int(t)) == want                            # the shipped helper (host code, no GPU needed)


@pytest.mark.skipif(not oraclelib.have_ref(), reason="oracle/_ref not built")
def test_port_drawrect_equals_compiled_reference():
    ref = oraclelib.Oracle("ref").draw
    rng = np.random.default_rng(11)
    W, H = 1280, 720                                                           # the reference's compile-time frame (stride 3840)
    PAD = 8                                                                    # rows of slack on both sides: the original has no bounds check
    big = rng.integers(0, 256, (H + 2 * PAD, W, 3), dtype=np.uint8)
    a = big[PAD:PAD + H]                                                       # view: drawRect's stray writes land in the slack rows
    b = a.copy()
    boxes = overlay_boxes(rng, 300, W, H)
    rgb = rng.integers(0, 1 << 24, len(boxes), dtype=np.uint32)
    for bx, col in zip(boxes, rgb):
        for k in range(3):
            ref.drawRect(C.c_void_p(a.ctypes.data), C.c_int(int(bx["l"]) + k), C.c_int(int(bx["t"]) + k), C.c_int(int(bx["r"]) - k),
                         C.c_int(int(bx["b"]) - k), C.c_uint32(int(col)))
    b = port_overlay(b, boxes, rgb, 3)
    assert np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(1280, 720), (1920, 1080)])
def test_gpu_overlay_equals_oracle(W, H):
    require_gpu()
    M = mot()
    rng = np.random.default_rng(W)
    nslots = 3
    ctx = M.Context(W, H, max_tracks=8, n_frame_slots=nslots, kind=M.TRACKER_KALMAN)
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(nslots)]
    for s, f in enumerate(frames):
        ctx.upload(s, f)
    n = 700
    boxes = overlay_boxes(rng, n, W, H)
    slots = rng.integers(0, nslots, n).astype(np.int32)
    tids = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    rgb = np.array([M.track_color(int(t)) for t in tids], np.uint32)
    ctx.overlay(slots, boxes, rgb, 3)
    for s in range(nslots):
        sel = slots == s
        want = port_overlay(frames[s], boxes[sel], rgb[sel], 3)
        got = ctx.download(s)
        assert np.array_equal(got, want), "slot %d: %d bytes differ" % (s, int((got != want).sum()))
    # thickness 1 on a fresh frame, and the empty call
    ctx.upload(0, frames[0])
    ctx.overlay(slots[:50] * 0, boxes[:50], rgb[:50], 1)
    assert np.array_equal(ctx.download(0), port_overlay(frames[0], boxes[:50], rgb[:50], 1))
    ctx.overlay(np.zeros(0, np.int32), boxes[:0], rgb[:0], 3)
    ctx.close()


@pytest.mark.gpu
def test_gpu_overlay_border_and_wraparound():
    """Boxes on the frame border: the shrunken rectangles swap corners and x can pass the row end, which drawRect's linear
    addressing turns into the next row; bytes beyond the buffer are dropped (the original would write out of bounds)."""
    require_gpu()
    M = mot()
    W, H = 640, 480
    ctx = M.Context(W, H, max_tracks=8, n_frame_slots=1, kind=M.TRACKER_KALMAN)
    f = np.zeros((H, W, 3), np.uint8)
    ctx.upload(0, f)
    b = boxes_array(6)
    b["l"] = [W - 1, 0, W - 2, 0, 5, 630]
    b["r"] = [W - 1, 0, W - 1, W - 1, 5, 639]
    b["t"] = [10, 0, H - 1, H - 1, 7, 470]
    b["b"] = [20, 0, H - 1, H - 1, 300, 479]
    rgb = np.array([0xFF0000, 0x00FF00, 0x0000FF, 0x123456, 0xABCDEF, 0x777777], np.uint32)
    ctx.overlay(np.zeros(6, np.int32), b, rgb, 3)
    assert np.array_equal(ctx.download(0), port_overlay(f, b, rgb, 3))
    ctx.close()


@pytest.mark.gpu
def test_frame_loop_overlay_equals_reference_loop(oracle):
    """The annotated frame of the batched loop (mot_td_step + mot_td_overlay) equals the oracle loop's track table drawn
    with the restated overlay: same tracks, same tids, hence the same colours and the same pixels."""
    require_gpu()
    M = mot()
    from synth import Scene
    W, H = 1280, 720
    sc = Scene(5, W, H, 20, tsize=40, win=64)
    ctx = M.Context(W, H, max_tracks=128, n_frame_slots=1, kind=M.TRACKER_KALMAN)
    td = ctx.td(0, cap=64, cost_mode=0)
    ref = oracle.td_new("kal", W, H, 64, 0)
    L = port(); L.port_track_color.restype = C.c_uint32; L.port_track_color.argtypes = [C.c_uint32]
    for f in range(12):
        sc.step()
        frame = sc.render()
        dets = sc.windows(jitter=2)
        td.step(frame, dets); ref.step(frame, dets)
        if f % 4 == 3:
            td.overlay()
            t = ref.tracks()
            rgb = np.array([L.port_track_color(int(x)) for x in t["tid"]], np.uint32)
            assert np.array_equal(ctx.download(0), port_overlay(frame, t["boxes"], rgb, 3)), f
    td.close(); ref.close(); ctx.close()
