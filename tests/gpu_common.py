"""Shared helpers for the -m gpu parity tests (CUDA path through the C ABI vs the oracle)."""
import numpy as np
import pytest


def require_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def mot():
    import mot_b200
    return mot_b200


def oracle_bins(O):
    """gradQuantize's nearest-bin rule (libhog/gradientMex.cpp:130-131, :143-144) applied to the oracle's O."""
    PI = np.float32(3.14159265)
    oMult = np.float32(18) / (np.float32(2) * PI)
    o = (O.astype(np.float32) * oMult).astype(np.float32)
    o0 = np.trunc((o + np.float32(.5)).astype(np.float32)).astype(np.int32)
    o0[o0 >= 18] = 0
    return o0


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def box_of(arr_elem):
    from synth import BBox
    return BBox(int(arr_elem["l"]), int(arr_elem["t"]), int(arr_elem["b"]), int(arr_elem["r"]), int(arr_elem["type"]), float(arr_elem["score"]))


def crop_gray(oracle, frame, b, rows, cols):
    """Oracle-side crop + gray + resize (port_rgb2gray / port_resize_gray are linked into every oracle .so)."""
    import ctypes as C
    rs, cs = b.b - b.t + 1, b.r - b.l + 1
    crop = np.zeros(rs * cs, np.float32); out = np.zeros(rows * cols, np.float32)
    oracle.kcf.port_rgb2gray(crop.ctypes.data_as(C.c_void_p), frame.ctypes.data_as(C.c_void_p), frame.strides[0], b.l, b.t, b.r, b.b)
    oracle.kcf.port_resize_gray(out.ctypes.data_as(C.c_void_p), crop.ctypes.data_as(C.c_void_p), rs, cs, rows, cols)
    return out.reshape(cols, rows).T       # (rows, cols) view of the column-major patch
