"""ctypes doors onto the oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Two backends with the same Python surface:
  Oracle("ref")   the UNMODIFIED reference compiled by path from /root/reference into oracle/_ref/*.so
                  (oracle/Makefile `ref`; prebuilt files travel to the GPU box, the sources do not)
  Oracle("port")  the C restatement oracle/port/*.c -> oracle/libport.so (always buildable)
`best()` returns "ref" when the prebuilt objects are present, else "port".
"""
import ctypes as C
import os
import subprocess
import numpy as np

from synth import BBox, BBOX_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

fptr = C.POINTER(C.c_float)
dptr = C.POINTER(C.c_double)
iptr = C.POINTER(C.c_int)


def _fp(a):
    return a.ctypes.data_as(fptr)


def build_port():
    subprocess.check_call(["make", "-s", "-C", ODIR, "port"])


def build_ref():
    subprocess.check_call(["make", "-s", "-C", ODIR, "ref"])


def have_ref():
    return all(os.path.exists(os.path.join(ODIR, "_ref", n)) for n in
               ("libref_kcf.so", "libref_kalman.so", "libref_hung.so", "libref_draw.so"))


def best():
    return "ref" if have_ref() else "port"


class Oracle:
    def __init__(self, kind="port"):
        self.kind = kind
        if kind == "port":
            if not os.path.exists(os.path.join(ODIR, "libport.so")):
                build_port()
            lib = C.CDLL(os.path.join(ODIR, "libport.so"))
            self.kcf = self.kal = self.hung = lib
            self.p = "port_"
        else:
            assert have_ref(), "oracle/_ref not built (make -C oracle ref needs /root/reference)"
            self.kcf = C.CDLL(os.path.join(ODIR, "_ref", "libref_kcf.so"))
            self.kal = C.CDLL(os.path.join(ODIR, "_ref", "libref_kalman.so"))
            self.hung = C.CDLL(os.path.join(ODIR, "_ref", "libref_hung.so"))
            self.draw = C.CDLL(os.path.join(ODIR, "_ref", "libref_draw.so"))
            self.p = "ref_"
            self.kcf.ref_fft_provider.restype = C.c_char_p
        for lib, names in ((self.kcf, ("kcf_new", "kcf_td_new")), (self.kal, ("kal_new", "kal_td_new"))):
            for n in names:
                getattr(lib, self.p + n).restype = C.c_void_p

    # ---- fHOG -------------------------------------------------------------------------------
    def fhog(self, I):
        """I: (h, w) float32 gray patch (any layout); returns H as (32, wb, hb) float32 (hb fastest)."""
        h, w = I.shape
        If = np.asfortranarray(I, dtype=np.float32)
        H = np.zeros((32, w // 4, h // 4), np.float32)
        getattr(self.kcf, self.p + "fhog_extract")(_fp(If), h, w, _fp(H))
        return H

    def gradmag(self, I):
        h, w = I.shape
        If = np.asfortranarray(I, dtype=np.float32)
        M = np.zeros((w, h), np.float32)
        O = np.zeros((w, h), np.float32)
        getattr(self.kcf, self.p + "gradmag")(_fp(If), _fp(M), _fp(O), h, w)
        return M, O          # indexed [x][y]

    def gradhist18(self, M, O):
        w, h = M.shape
        R1 = np.zeros((18, w // 4, h // 4), np.float32)
        getattr(self.kcf, self.p + "gradhist18")(_fp(M), _fp(O), _fp(R1), h, w)
        return R1

    # ---- KCF plugin -------------------------------------------------------------------------
    def kcf_new(self, box):
        return C.c_void_p(getattr(self.kcf, self.p + "kcf_new")(C.byref(box)))

    def kcf_predict(self, h, gray, box):
        g = np.asfortranarray(gray, dtype=np.float32)
        getattr(self.kcf, self.p + "kcf_predict")(h, _fp(g), C.byref(box))

    def kcf_update(self, h, gray, box):
        g = np.asfortranarray(gray, dtype=np.float32)
        getattr(self.kcf, self.p + "kcf_update")(h, _fp(g), C.byref(box))

    def kcf_delete(self, h):
        getattr(self.kcf, self.p + "kcf_delete")(h)

    def kcf_dims(self, h):
        d = (C.c_int * 6)()
        getattr(self.kcf, self.p + "kcf_dims")(h, d)
        return dict(rows=d[0], cols=d[1], hr=d[2], wc=d[3], chan=d[4], S=d[5])

    GET = dict(xf_tm=0, xf_fq=1, xf_md=2, yf=3, kf=4, zf=5, alpha=6, response=7, labels=8, cos_win=9)

    def kcf_get(self, h, what):
        d = self.kcf_dims(h)
        nb, S = d["hr"] * d["wc"], d["S"]
        n = {0: 31 * nb, 1: 62 * S, 2: 62 * S, 3: 2 * S, 4: 2 * S, 5: 2 * S, 6: S, 7: nb, 8: nb, 9: nb}[self.GET[what]]
        out = np.zeros(n, np.float32)
        f = getattr(self.kcf, self.p + "kcf_get")
        f.restype = C.c_long
        got = f(h, self.GET[what], _fp(out))
        assert got == n
        return out

    # ---- Kalman plugin ----------------------------------------------------------------------
    def kal_new(self, box):
        return C.c_void_p(getattr(self.kal, self.p + "kal_new")(C.byref(box)))

    def kal_predict(self, h, box):
        getattr(self.kal, self.p + "kal_predict")(h, None, C.byref(box))

    def kal_update(self, h, box):
        getattr(self.kal, self.p + "kal_update")(h, None, C.byref(box))

    def kal_delete(self, h):
        getattr(self.kal, self.p + "kal_delete")(h)

    def kal_state(self, h):
        x = np.zeros(6); P = np.zeros(36); K = np.zeros(24)
        getattr(self.kal, self.p + "kal_state")(h, x.ctypes.data_as(dptr), P.ctypes.data_as(dptr), K.ctypes.data_as(dptr))
        return x, P.reshape(6, 6).T.copy(), K.reshape(4, 6).T.copy()

    # ---- association ------------------------------------------------------------------------
    def assign(self, dist):
        """dist: (nrows, ncols) float64; returns (assignment[nrows] int32, cost)."""
        nr, nc = dist.shape
        d = np.asfortranarray(dist, dtype=np.float64)
        a = np.full(nr, -2, np.int32)
        cost = C.c_double(0)
        getattr(self.hung, self.p + "assignmentoptimal")(a.ctypes.data_as(iptr), C.byref(cost), d.ctypes.data_as(dptr), nr, nc)
        return a, cost.value

    # ---- frame loop -------------------------------------------------------------------------
    def td_new(self, tracker, W, H, cap=256, cost_mode=0):
        lib, pre = (self.kcf, self.p + "kcf_td_") if tracker == "kcf" else (self.kal, self.p + "kal_td_")
        return TdLoop(lib, pre, W, H, cap, cost_mode)


class TdLoop:
    def __init__(self, lib, pre, W, H, cap, cost_mode):
        self.lib, self.pre, self.cap = lib, pre, cap
        self.h = C.c_void_p(getattr(lib, pre + "new")(W, H, cap, cost_mode))

    def step(self, bgr, dets):
        d = np.ascontiguousarray(dets)
        assert d.dtype == BBOX_DTYPE
        p = bgr.ctypes.data_as(C.c_void_p) if bgr is not None else None
        getattr(self.lib, self.pre + "step")(self.h, p, d.ctypes.data_as(C.c_void_p), len(d))

    def tracks(self):
        f = getattr(self.lib, self.pre + "ntracks"); n = f(self.h)
        tid = np.zeros(n, np.uint32); boxes = np.zeros(n, BBOX_DTYPE)
        age = np.zeros(n, np.int32); vis = np.zeros(n, np.int32); inv = np.zeros(n, np.int32)
        getattr(self.lib, self.pre + "get")(self.h, tid.ctypes.data_as(C.c_void_p), boxes.ctypes.data_as(C.c_void_p),
                                            age.ctypes.data_as(iptr), vis.ctypes.data_as(iptr), inv.ctypes.data_as(iptr))
        return dict(tid=tid, boxes=boxes, age=age, vis=vis, inv=inv)

    def last(self):
        pred = np.zeros(self.cap, BBOX_DTYPE); asg = np.zeros(self.cap, np.int32)
        n = getattr(self.lib, self.pre + "last")(self.h, pred.ctypes.data_as(C.c_void_p), asg.ctypes.data_as(iptr))
        return pred[:n], asg[:n]

    def close(self):
        if self.h:
            getattr(self.lib, self.pre + "free")(self.h)
            self.h = None
