"""Seeded synthetic scenes for the parity tests and bench.py.

The reference has no test data at all: webcam frames in, YOLO boxes in.
Both are replaced by seeded synthetic streams of the named shapes: BGR u8 frames (low-frequency
background + noise + textured square targets), constant-velocity target motion reflecting at the
borders, and detections in the reference's bbox_t layout (top/cnntype.h:36-41, field order l,t,b,r).
Everything is generated on the host with numpy's PCG64 so the oracle and the CUDA path see the same bytes.
"""
import ctypes as C
import numpy as np


class BBox(C.Structure):
    """top/cnntype.h:36-41 -- inclusive pixel coordinates, field order l, t, b, r."""
    _fields_ = [("l", C.c_int), ("t", C.c_int), ("b", C.c_int), ("r", C.c_int),
                ("type", C.c_int), ("score", C.c_float)]

    def tup(self):
        return (self.l, self.t, self.b, self.r)


BBOX_DTYPE = np.dtype([("l", "<i4"), ("t", "<i4"), ("b", "<i4"), ("r", "<i4"), ("type", "<i4"), ("score", "<f4")])
assert BBOX_DTYPE.itemsize == C.sizeof(BBox) == 24


def boxes_array(n):
    return np.zeros(n, dtype=BBOX_DTYPE)


def make_background(rng, H, W):
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    bg = np.full((H, W), 120.0, np.float32)
    for amp in (40.0, 25.0, 15.0):
        fx, fy = rng.uniform(-0.01, 0.01, 2)
        ph = rng.uniform(0, 2 * np.pi)
        bg += amp * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
    img = np.stack([bg * 0.9, bg, bg * 1.05], axis=-1)
    img += rng.integers(-8, 9, size=img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_texture(rng, size):
    """8x8 random block texture up-sampled to size x size, contrast >= 64 LSB, 3 channels."""
    blocks = rng.integers(0, 2, size=(8, 8, 1)) * rng.integers(64, 160) + rng.integers(20, 80, size=(8, 8, 3))
    rep = -(-size // 8)
    tex = np.repeat(np.repeat(blocks, rep, axis=0), rep, axis=1)[:size, :size]
    return np.clip(tex, 0, 255).astype(np.uint8)


class Scene:
    """One stream: n targets of tsize px moving at constant velocity inside a W x H frame.

    Each target is tracked through a win x win window centred on it (the harness-side "padding":
    the reference has no padding concept, the filter window is the bbox handed to tracker_new,
    trackers/kcf.cpp:148-156)."""

    def __init__(self, seed, W, H, n, tsize=48, win=128, vmax=3.0, n_classes=3, grid=True):
        self.rng = np.random.default_rng(seed)
        self.W, self.H, self.n, self.tsize, self.win = W, H, n, tsize, win
        self.bg = make_background(self.rng, H, W)
        self.tex = [make_texture(self.rng, tsize) for _ in range(n)]
        half = win // 2
        self.lo = np.array([half + 4, half + 4], np.float64)
        self.hi = np.array([W - half - 4, H - half - 4], np.float64)
        if grid:
            step = tsize + 8
            nx = max(1, int((self.hi[0] - self.lo[0]) // step) + 1)
            ny = max(1, int((self.hi[1] - self.lo[1]) // step) + 1)
            assert nx * ny >= n, "frame too small for %d targets" % n
            slots = self.rng.permutation(nx * ny)[:n]
            self.pos = np.stack([self.lo[0] + (slots % nx) * step, self.lo[1] + (slots // nx) * step], axis=1).astype(np.float64)
        else:
            self.pos = self.rng.uniform(self.lo, self.hi, size=(n, 2))
        self.vel = self.rng.uniform(-vmax, vmax, size=(n, 2))
        self.cls = self.rng.integers(0, n_classes, size=n).astype(np.int32)

    def step(self):
        self.pos += self.vel + self.rng.normal(0.0, 0.5, size=self.pos.shape)
        for a in range(2):
            lo_hit = self.pos[:, a] < self.lo[a]
            hi_hit = self.pos[:, a] > self.hi[a]
            self.pos[lo_hit, a] = 2 * self.lo[a] - self.pos[lo_hit, a]
            self.pos[hi_hit, a] = 2 * self.hi[a] - self.pos[hi_hit, a]
            self.vel[lo_hit | hi_hit, a] *= -1
        self.pos = np.clip(self.pos, self.lo, self.hi)

    def render(self):
        img = self.bg.copy()
        h = self.tsize // 2
        for i in range(self.n):
            cx, cy = int(round(self.pos[i, 0])), int(round(self.pos[i, 1]))
            img[cy - h:cy - h + self.tsize, cx - h:cx - h + self.tsize] = self.tex[i]
        return img

    def windows(self, jitter=0):
        """win x win boxes centred on the targets (pure translation jitter keeps size == template size)."""
        b = boxes_array(self.n)
        half = self.win // 2
        cx = np.rint(self.pos[:, 0]).astype(np.int32)
        cy = np.rint(self.pos[:, 1]).astype(np.int32)
        if jitter:
            cx = cx + self.rng.integers(-jitter, jitter + 1, size=self.n).astype(np.int32)
            cy = cy + self.rng.integers(-jitter, jitter + 1, size=self.n).astype(np.int32)
        b["l"] = np.clip(cx - half, 0, self.W - self.win)
        b["t"] = np.clip(cy - half, 0, self.H - self.win)
        b["r"] = b["l"] + self.win - 1
        b["b"] = b["t"] + self.win - 1
        b["type"] = self.cls
        b["score"] = 1.0
        return b


def random_boxes(rng, n, W, H, smin=16, smax=128, n_classes=3):
    """Uniform boxes on a W x H canvas (config C5 of SURVEY.md section 8d)."""
    b = boxes_array(n)
    w = rng.integers(smin, smax + 1, size=n)
    h = rng.integers(smin, smax + 1, size=n)
    b["l"] = rng.integers(0, W - w)
    b["t"] = rng.integers(0, H - h)
    b["r"] = b["l"] + w - 1
    b["b"] = b["t"] + h - 1
    b["type"] = rng.integers(0, n_classes, size=n)
    b["score"] = 1.0
    return b


def jittered_detections(rng, truth, W, H, jitter=2, drop=0.0, n_false=0, shuffle=True):
    """Detections = ground-truth boxes + integer jitter, some dropped, plus false positives."""
    keep = rng.random(len(truth)) >= drop
    d = truth[keep].copy()
    for f in ("l", "r"):
        d[f] = np.clip(d[f] + rng.integers(-jitter, jitter + 1, size=len(d)), 0, W - 1)
    for f in ("t", "b"):
        d[f] = np.clip(d[f] + rng.integers(-jitter, jitter + 1, size=len(d)), 0, H - 1)
    if n_false:
        d = np.concatenate([d, random_boxes(rng, n_false, W, H)])
    if shuffle:
        d = d[rng.permutation(len(d))]
    return np.ascontiguousarray(d)
