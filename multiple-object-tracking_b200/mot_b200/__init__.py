"""ctypes binding of libmot_b200.so (include/mot_b200.h) for the tests and bench.py.

Thin on purpose: the product is the CUDA library and its C ABI; the host side of the drop-in is C++
(host/td_loop.cpp, host/tracker_shim.cpp).  This module never computes anything itself and there is no
fallback: if the shared library is missing, or a call fails, it raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("MOT_B200_LIB") or os.path.join(PKG_DIR, "libmot_b200.so")   # override: A/B builds side by side on one GPU box

TRACKER_KALMAN, TRACKER_KCF = 0, 1
COST_REF_CENTROID, COST_IOU_CLAMPED = 0, 1

BBOX_DTYPE = np.dtype([("l", "<i4"), ("t", "<i4"), ("b", "<i4"), ("r", "<i4"), ("type", "<i4"), ("score", "<f4")])
CHAIN_DTYPE = np.dtype([("nbox", "<i4"), ("bbox", BBOX_DTYPE, (128,))])        # bbox_chain_t, top/cnntype.h:43-47 (3076 bytes)


def make_chain(dets):
    """A detector-style bbox_chain_t holding `dets` (at most 128 boxes)."""
    c = np.zeros(1, CHAIN_DTYPE)
    c["nbox"] = len(dets)
    c["bbox"][0, :len(dets)] = dets
    return c

EXPORTS = [
    "mot_ctx_create", "mot_ctx_destroy", "mot_ctx_set_kcf_options", "mot_ctx_reserve_window", "mot_last_error", "mot_ctx_set_stream", "mot_sync", "mot_ctx_kind", "mot_launch_count",
    "mot_frame_upload", "mot_frame_bind_device", "mot_frame_download", "mot_overlay_batch", "mot_track_color", "mot_yolo_post", "mot_tracker_new_batch", "mot_tracker_delete_batch", "mot_tracker_spawnable",
    "mot_predict_batch", "mot_update_batch", "mot_track_batch", "mot_predict_batch_dev", "mot_update_batch_dev", "mot_predict_gray", "mot_update_gray",
    "mot_crop_gray_resize", "mot_rgb2gray_host", "mot_resize_gray_host", "mot_associate_batch", "mot_assign_batch", "mot_associate_batch_dev",
    "mot_td_create", "mot_td_destroy", "mot_td_step", "mot_td_step_chain", "mot_td_step_chain_batch", "mot_td_step_multi", "mot_td_ntracks", "mot_td_dropped", "mot_td_get", "mot_td_last", "mot_td_overlay",
    "mot_tdd_create", "mot_tdd_destroy", "mot_tdd_step_dev", "mot_tdd_step", "mot_tdd_step_chains", "mot_tdd_read", "mot_tdd_kcf_windows", "mot_tdd_dropped", "mot_tdd_frame_base",
    "mot_debug_enable_dumps", "mot_debug_fetch", "mot_debug_state", "mot_debug_tables", "mot_debug_any_plan",
]


def build(verbose=False):
    """Compile libmot_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", PKG_DIR, "-j%d" % max(4, os.cpu_count() or 4)], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libmot_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


_lib = None


def track_color(tid):
    """colormap[hashcolor(tid + 1) & 255] of top/td.cpp:295-305, 619-620, 652-699 (host-only helper, no GPU needed)."""
    L = lib()
    L.mot_track_color.restype = C.c_uint32
    L.mot_track_color.argtypes = [C.c_uint32]
    return int(L.mot_track_color(int(tid) & 0xFFFFFFFF))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run __graft_entry__.build() (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.mot_last_error.restype = C.c_char_p
        L.mot_launch_count.restype = C.c_long
        L.mot_debug_fetch.restype = C.c_long
        L.mot_debug_state.restype = C.c_long
        L.mot_debug_tables.restype = C.c_long
        for name, types in {
            "mot_associate_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int,
                                    C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p],
            "mot_associate_batch_dev": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int,
                                        C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_int],
            "mot_assign_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p],
            "mot_predict_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int],
            "mot_update_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
            "mot_track_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int],
            "mot_predict_batch_dev": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int],
            "mot_update_batch_dev": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
            "mot_frame_upload": [C.c_void_p, C.c_int, C.c_void_p, C.c_int],
            "mot_frame_bind_device": [C.c_void_p, C.c_int, C.c_void_p, C.c_int],
            "mot_tracker_new_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
            "mot_tracker_delete_batch": [C.c_void_p, C.c_int, C.c_void_p],
            "mot_predict_gray": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
            "mot_update_gray": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
            "mot_crop_gray_resize": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
            "mot_ctx_set_stream": [C.c_void_p, C.c_void_p],
            "mot_sync": [C.c_void_p], "mot_ctx_destroy": [C.c_void_p], "mot_launch_count": [C.c_void_p], "mot_ctx_kind": [C.c_void_p],
            "mot_debug_enable_dumps": [C.c_void_p, C.c_int],
            "mot_debug_fetch": [C.c_void_p, C.c_int, C.c_void_p, C.c_long],
            "mot_debug_state": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_long],
            "mot_debug_tables": [C.c_int, C.c_void_p, C.c_long, C.c_void_p],
            "mot_td_create": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int],
            "mot_td_destroy": [C.c_void_p],
            "mot_td_step": [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int],
            "mot_td_step_chain": [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p],
            "mot_td_step_chain_batch": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p],
            "mot_tdd_step_chains": [C.c_void_p, C.c_void_p],
            "mot_td_step_multi": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p],
            "mot_td_ntracks": [C.c_void_p],
            "mot_td_get": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "mot_td_last": [C.c_void_p, C.c_void_p, C.c_void_p],
            "mot_tdd_create": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int],
            "mot_tdd_destroy": [C.c_void_p],
            "mot_tdd_step_dev": [C.c_void_p, C.c_void_p, C.c_void_p],
            "mot_tdd_step": [C.c_void_p, C.c_void_p, C.c_void_p],
            "mot_tdd_read": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        }.items():
            getattr(L, name).argtypes = types
        _lib = L
    return _lib


class MotError(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise MotError("libmot_b200 error %d: %s" % (rc, lib().mot_last_error().decode()))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _boxes(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == BBOX_DTYPE, a.dtype
    return a


STAGES = dict(gray=0, m0=1, bin=2, r1=3, norm=4, feat=5, spec=6, zf=7, resp=8, kf=9, peak=10)


class Context:
    def __init__(self, W, H, max_tracks=256, n_frame_slots=1, kind=TRACKER_KCF, device=0):
        self.W, self.H, self.kind = W, H, kind
        h = C.c_void_p()
        _chk(lib().mot_ctx_create(C.byref(h), device, W, H, max_tracks, n_frame_slots, kind))
        self.h = h

    def close(self):
        if self.h:
            lib().mot_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # context
    def set_kcf_options(self, gaussian=False, sigma=0.5, subpixel=False, padding=0.0, osf=0.0):
        """North-star extensions (mot_ctx_set_kcf_options); call before the first tracker is created."""
        class Opt(C.Structure):
            _fields_ = [("gaussian_kernel", C.c_int), ("kernel_sigma", C.c_float), ("subpixel_peak", C.c_int), ("padding", C.c_float), ("output_sigma_factor", C.c_float)]
        o = Opt(1 if gaussian else 0, sigma, 1 if subpixel else 0, padding, osf)
        _chk(lib().mot_ctx_set_kcf_options(self.h, C.byref(o)))

    def reserve_window(self, max_rows, max_cols):
        """Size the model arena for windows up to max_rows x max_cols px (mot_ctx_reserve_window); before the first tracker."""
        _chk(lib().mot_ctx_reserve_window(self.h, int(max_rows), int(max_cols)))

    def set_stream(self, cuda_stream_ptr):
        _chk(lib().mot_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        _chk(lib().mot_sync(self.h))

    def launches(self):
        return lib().mot_launch_count(self.h)

    # frames
    def upload(self, slot, frame):
        assert frame.dtype == np.uint8 and frame.flags.c_contiguous
        _chk(lib().mot_frame_upload(self.h, slot, _p(frame), frame.strides[0]))

    def bind_device(self, slot, dev_ptr, stride):
        _chk(lib().mot_frame_bind_device(self.h, slot, C.c_void_p(dev_ptr), stride))

    def download(self, slot):
        out = np.zeros((self.H, self.W, 3), np.uint8)
        _chk(lib().mot_frame_download(self.h, slot, _p(out), out.strides[0]))
        return out

    def yolo_post(self, outs, anchors, obj_thresh, nms_thresh, th, tw, ih, iw, nc, max_out=4096):
        """detectors/yolo3.cpp:141-356 + 487-527 on the device: three output maps of one image -> detections (bbox array)."""
        o = [np.ascontiguousarray(a, np.float32) for a in outs]; an = np.ascontiguousarray(anchors, np.int32)
        out = np.zeros(max_out, BBOX_DTYPE)
        f = lib().mot_yolo_post
        f.restype = C.c_int
        n = f(self.h, _p(o[0]), _p(o[1]), _p(o[2]), _p(an), C.c_float(obj_thresh), C.c_float(nms_thresh), th, tw, ih, iw, nc, _p(out), max_out)
        if n < 0:
            _chk(n)
        return out[:n]

    def overlay(self, frame_slots, boxes, rgb, thickness=3):
        """Tracking rectangles of top/td.cpp:647-733 drawn into the frames on the device, entries in order."""
        fs = _i32(frame_slots); b = _boxes(boxes); col = np.ascontiguousarray(rgb, np.uint32)
        assert len(fs) == len(b) == len(col)
        _chk(lib().mot_overlay_batch(self.h, len(fs), _p(fs), _p(b), _p(col), thickness))

    # trackers
    def new(self, boxes):
        b = _boxes(boxes); out = np.zeros(len(b), np.int32)
        _chk(lib().mot_tracker_new_batch(self.h, len(b), _p(b), _p(out)))
        return out

    def delete(self, handles):
        hs = _i32(handles)
        _chk(lib().mot_tracker_delete_batch(self.h, len(hs), _p(hs)))

    def predict(self, handles, frame_slots, boxes, clamp=0):
        hs = _i32(handles); fs = None if frame_slots is None else _i32(frame_slots); b = _boxes(boxes).copy()
        _chk(lib().mot_predict_batch(self.h, len(hs), _p(hs), _p(fs), _p(b), clamp))
        return b

    def track(self, handles, frame_slots, boxes, clamp=1):
        """predict (+ clamp) and update with the predicted box in one call (mot_track_batch)."""
        hs = _i32(handles); fs = None if frame_slots is None else _i32(frame_slots); b = _boxes(boxes).copy()
        _chk(lib().mot_track_batch(self.h, len(hs), _p(hs), _p(fs), _p(b), clamp))
        return b

    def update(self, handles, frame_slots, boxes):
        hs = _i32(handles); fs = None if frame_slots is None else _i32(frame_slots); b = _boxes(boxes)
        _chk(lib().mot_update_batch(self.h, len(hs), _p(hs), _p(fs), _p(b)))

    def predict_dev(self, n, d_handles, d_frames, d_boxes, clamp=0):
        _chk(lib().mot_predict_batch_dev(self.h, n, C.c_void_p(d_handles), C.c_void_p(d_frames), C.c_void_p(d_boxes), clamp))

    def update_dev(self, n, d_handles, d_frames, d_boxes):
        _chk(lib().mot_update_batch_dev(self.h, n, C.c_void_p(d_handles), C.c_void_p(d_frames), C.c_void_p(d_boxes)))

    def predict_gray(self, handle, gray, box):
        g = np.asfortranarray(gray, dtype=np.float32); b = _boxes(box).copy()
        _chk(lib().mot_predict_gray(self.h, int(handle), _p(g), _p(b)))
        return b

    def update_gray(self, handle, gray, box):
        g = np.asfortranarray(gray, dtype=np.float32); b = _boxes(box)
        _chk(lib().mot_update_gray(self.h, int(handle), _p(g), _p(b)))

    def crop_gray_resize(self, slot, box, rows_d, cols_d):
        out = np.zeros(rows_d * cols_d, np.float32); b = _boxes(box)
        _chk(lib().mot_crop_gray_resize(self.h, slot, _p(b), rows_d, cols_d, _p(out)))
        return out

    # association
    def associate(self, trk_list, det_list, cost_mode=COST_REF_CENTROID, want_dist=False):
        n = len(trk_list)
        T = np.array([len(t) for t in trk_list], np.int32); D = np.array([len(d) for d in det_list], np.int32)
        mt, mdd = max(1, int(T.max())), max(1, int(D.max())); md = max(mt, mdd)
        trk = np.zeros((n, mt), BBOX_DTYPE); det = np.zeros((n, mdd), BBOX_DTYPE)
        for m in range(n):
            trk[m, :T[m]] = trk_list[m]; det[m, :D[m]] = det_list[m]
        assign = np.full((n, md), -9, np.int32); cost = np.zeros(n, np.float64)
        dist = np.zeros((n, md * md), np.float64) if want_dist else None
        _chk(lib().mot_associate_batch(self.h, n, _p(T), _p(D), _p(trk), mt, _p(det), mdd, cost_mode, _p(dist), md * md, _p(assign), md, _p(cost)))
        nr = np.minimum(T, D)
        outs = [assign[m, :nr[m]].copy() for m in range(n)]
        if want_dist:
            return outs, cost, [dist[m, :T[m] * D[m]].reshape(max(T[m], D[m]), nr[m]).T.copy() for m in range(n)]
        return outs, cost

    def assign(self, mats):
        """mats: list of (nrows, ncols) float64 arrays; returns (list of assignment vectors, costs)."""
        n = len(mats)
        nr = np.array([m.shape[0] for m in mats], np.int32); nc = np.array([m.shape[1] for m in mats], np.int32)
        md = int(max(nr.max(), nc.max()))
        buf = np.zeros((n, md * md), np.float64)
        for i, m in enumerate(mats):
            buf[i, :m.size] = np.asfortranarray(m, dtype=np.float64).ravel(order="F")
        assign = np.full((n, md), -9, np.int32); cost = np.zeros(n, np.float64)
        _chk(lib().mot_assign_batch(self.h, n, _p(nr), _p(nc), _p(buf), md * md, _p(assign), md, _p(cost)))
        return [assign[i, :nr[i]].copy() for i in range(n)], cost

    # test hooks
    def enable_dumps(self, on=True):
        _chk(lib().mot_debug_enable_dumps(self.h, 1 if on else 0))

    def fetch(self, stage, dtype=np.float32, max_items=1 << 20):
        out = np.zeros(max_items, dtype)
        n = lib().mot_debug_fetch(self.h, STAGES[stage], _p(out), out.nbytes)
        if n < 0:
            _chk(int(n))
        return out[: n // out.itemsize].copy()

    def state(self, handle, which):
        code = dict(xf_md=0, alpha=1, x=2, P=3, subpixel=4, alpha_im=5)[which]
        out = np.zeros(1 << 18, np.float64 if code in (2, 3) else np.float32)
        n = lib().mot_debug_state(self.h, int(handle), code, _p(out), out.nbytes)
        if n < 0:
            _chk(int(n))
        out = out[: n // out.itemsize].copy()
        return out.reshape(6, 6).T.copy() if which == "P" else out

    def td(self, frame_slot=0, cap=256, cost_mode=COST_REF_CENTROID):
        return TdLoop(self, frame_slot, cap, cost_mode)


class DeviceLoop:
    """mot_tdd_*: the frame loop with the track tables resident on the device (Kalman or KCF contexts; KCF: stream s reads frame slot s)."""

    def __init__(self, ctx, n_streams, cap=256, max_det=128, cost_mode=COST_REF_CENTROID):
        self.ctx, self.n, self.cap = ctx, n_streams, cap
        h = C.c_void_p()
        _chk(lib().mot_tdd_create(C.byref(h), ctx.h, n_streams, cap, max_det, cost_mode))
        self.h = h

    def step(self, dets):
        dl = [_boxes(d) for d in dets]
        arr = (C.c_void_p * self.n)(*[d.ctypes.data if len(d) else None for d in dl])
        nd = np.array([len(d) for d in dl], np.int32)
        _chk(lib().mot_tdd_step(self.h, arr, _p(nd)))

    def step_chains(self, chains):
        arr = (C.c_void_p * self.n)(*[c.ctypes.data if c is not None else None for c in chains])
        _chk(lib().mot_tdd_step_chains(self.h, arr))

    def step_dev(self, d_dets, d_ndet):
        _chk(lib().mot_tdd_step_dev(self.h, C.c_void_p(d_dets), C.c_void_p(d_ndet)))

    def frame_base(self, base):
        """KCF kind: stream s reads frame slot base + s from the next step on (alternate two bases to overlap uploads)."""
        _chk(lib().mot_tdd_frame_base(self.h, base))

    def kcf_windows(self, sizes):
        r = np.array([a for a, _ in sizes], np.int32); c = np.array([b for _, b in sizes], np.int32)
        _chk(lib().mot_tdd_kcf_windows(self.h, len(r), _p(r), _p(c)))

    def dropped(self, s):
        n = lib().mot_tdd_dropped(self.h, s)
        if n < 0:
            _chk(n)
        return n

    def tracks(self, s):
        tid = np.zeros(self.cap, np.uint32); boxes = np.zeros(self.cap, BBOX_DTYPE)
        age = np.zeros(self.cap, np.int32); vis = np.zeros(self.cap, np.int32); inv = np.zeros(self.cap, np.int32)
        n = lib().mot_tdd_read(self.h, s, _p(tid), _p(boxes), _p(age), _p(vis), _p(inv))
        if n < 0:
            _chk(n)
        return dict(tid=tid[:n], boxes=boxes[:n], age=age[:n], vis=vis[:n], inv=inv[:n])

    def close(self):
        if self.h:
            lib().mot_tdd_destroy(self.h)
            self.h = None


class TdLoop:
    def __init__(self, ctx, frame_slot, cap, cost_mode):
        self.ctx, self.cap = ctx, cap
        h = C.c_void_p()
        _chk(lib().mot_td_create(C.byref(h), ctx.h, frame_slot, cap, cost_mode))
        self.h = h

    def step(self, frame, dets):
        d = _boxes(dets)
        if frame is None:
            _chk(lib().mot_td_step(self.h, None, 0, _p(d), len(d)))
        else:
            assert frame.dtype == np.uint8 and frame.flags.c_contiguous
            _chk(lib().mot_td_step(self.h, _p(frame), frame.strides[0], _p(d), len(d)))

    def step_chain(self, frame, chain):
        _chk(lib().mot_td_step_chain(self.h, _p(frame) if frame is not None else None, frame.strides[0] if frame is not None else 0, _p(chain)))

    def step_chain_batch(self, frames, chains):
        n = len(chains)
        arr_f = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        arr_c = (C.c_void_p * n)(*[c.ctypes.data for c in chains])
        _chk(lib().mot_td_step_chain_batch(self.h, n, arr_f, frames[0].strides[0], arr_c))

    def tracks(self):
        n = lib().mot_td_ntracks(self.h)
        tid = np.zeros(n, np.uint32); boxes = np.zeros(n, BBOX_DTYPE)
        age = np.zeros(n, np.int32); vis = np.zeros(n, np.int32); inv = np.zeros(n, np.int32)
        lib().mot_td_get(self.h, _p(tid), _p(boxes), _p(age), _p(vis), _p(inv))
        return dict(tid=tid, boxes=boxes, age=age, vis=vis, inv=inv)

    def last(self):
        pred = np.zeros(self.cap, BBOX_DTYPE); asg = np.zeros(self.cap, np.int32)
        n = lib().mot_td_last(self.h, _p(pred), _p(asg))
        return pred[:n], asg[:n]

    def overlay(self):
        """Draw the current tracks into the loop's frame slot (top/td.cpp:647-733); ctx.download(slot) reads it back."""
        _chk(lib().mot_td_overlay(self.h))

    def close(self):
        if self.h:
            lib().mot_td_destroy(self.h)
            self.h = None


def step_multi(tds, frames, dets):
    """Lock-step over several streams (mot_td_step_multi).  frames: list of HxWx3 uint8 or None."""
    n = len(tds)
    arr_td = (C.c_void_p * n)(*[t.h for t in tds])
    dl = [_boxes(d) for d in dets]
    arr_d = (C.c_void_p * n)(*[d.ctypes.data for d in dl])
    nd = np.array([len(d) for d in dl], np.int32)
    if frames is None:
        _chk(lib().mot_td_step_multi(arr_td, n, None, 0, arr_d, _p(nd)))
    else:
        arr_f = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        _chk(lib().mot_td_step_multi(arr_td, n, arr_f, frames[0].strides[0], arr_d, _p(nd)))
