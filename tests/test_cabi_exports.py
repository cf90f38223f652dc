"""CPU-side checks of the product library: it builds, loads, exports every symbol include/mot_b200.h declares,
and refuses to run without a CUDA device (no CPU fallback).  No compute calls are made here."""
import os
import re

import pytest


def test_library_builds_and_exports_every_declared_symbol():
    import mot_b200
    mot_b200.build()
    L = mot_b200.lib()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "mot_b200.h")).read()
    declared = set(re.findall(r"\b(mot_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(L, name), "libmot_b200.so does not export %s" % name
    assert declared == set(mot_b200.EXPORTS), declared ^ set(mot_b200.EXPORTS)


def test_reference_plugin_symbols_are_exported():
    """The symbols top/td.cpp:229-261 declares, under the names a TU written against top/cnntype.h needs (bbox_t is
    `struct _bbox_pos_s`, so tracker_new mangles to _Z11tracker_newP11_bbox_pos_s); tests/test_link_plugin.py proves the
    list by linking such a TU with -Wl,--no-undefined."""
    import subprocess
    import mot_b200
    out = subprocess.run(["nm", "-D", "--defined-only", mot_b200.LIB_PATH], capture_output=True, text=True).stdout
    have = {l.split()[-1] for l in out.splitlines() if l.strip()}
    here = os.path.dirname(os.path.abspath(__file__))
    expected = [l.strip() for l in open(os.path.join(here, "link", "expected_symbols.txt")) if l.strip() and not l.startswith("#")]
    assert len(expected) == 7
    for name in expected:
        assert name in have, name


def test_no_cpu_fallback_without_a_device():
    import torch
    import mot_b200
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(mot_b200.MotError):
        mot_b200.Context(640, 480)


def test_sse_tables_match_the_oracle_acos_table(port):
    """The host-harvested tables are part of the product; pin the acos table against the oracle's (bit-exact)."""
    import ctypes as C
    import numpy as np
    import mot_b200
    info = (C.c_int * 4)()
    n = mot_b200.lib().mot_debug_tables(2, None, 0, info)
    assert n == 20020
    a = np.zeros(n, np.float32)
    assert mot_b200.lib().mot_debug_tables(2, a.ctypes.data_as(C.c_void_p), n, info) == n
    b = np.zeros(20020, np.float32)
    port.kcf.port_acos_table(b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)
    assert 6 <= info[0] <= 14 and 6 <= info[1] <= 14


def test_gray_conversion_identity_for_all_colours():
    """The kernel computes gray = RN_f32((144B + 587G + 299R) / 1000) without FP64 (csrc/kcf_fused.cuh: bgr_gray).
    Prove, for every 8-bit colour, that this is the reference's (float)(0.144*B + 0.587*G + 0.299*R) evaluated in
    double (top/drawlib.c:234)."""
    import numpy as np
    B, G, R = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    B = B.ravel().astype(np.float64); G = G.ravel().astype(np.float64); R = R.ravel().astype(np.float64)
    ref = ((0.144 * B + 0.587 * G) + 0.299 * R).astype(np.float32)
    N = (144 * B + 587 * G + 299 * R).astype(np.int64)
    nf = N.astype(np.float32)
    assert np.array_equal(nf.astype(np.int64), N)
    assert np.array_equal(ref, nf / np.float32(1000.0)), "IEEE single division of the exact integer numerator"
    # the kernel's division-free form: q0 = nf*rcp; rem = fma(-q0, 1000, nf); q = fma(rem, rcp, q0)
    rcp = np.float32(1.0) / np.float32(1000.0)
    q0 = (nf * rcp).astype(np.float32)
    rem = nf.astype(np.float64) - q0.astype(np.float64) * 1000.0          # exact in double
    assert np.array_equal(rem, rem.astype(np.float32).astype(np.float64)), "the remainder is exactly representable"
    q = (q0.astype(np.float64) + rem * np.float64(rcp))
    # distance of the exact fma argument to the nearest float rounding boundary dwarfs one double rounding
    qf = q.astype(np.float32)
    assert np.array_equal(ref, qf)


def _table(which, dtype=None):
    import ctypes as C
    import numpy as np
    import mot_b200
    info = (C.c_int * 4)()
    n = mot_b200.lib().mot_debug_tables(which, None, 0, info)
    assert n > 0
    a = np.zeros(n, np.float32)
    assert mot_b200.lib().mot_debug_tables(which, a.ctypes.data_as(C.c_void_p), n, info) == n
    return (a if dtype is None else a.view(dtype)), [int(v) for v in info]


def test_harvested_tables_reproduce_the_sse_instructions(port):
    """The fused kernel never executes rsqrtps / rcpps: it looks them up in tables harvested from the host CPU.  Pin those
    tables against the instructions themselves (through the oracle's C wrappers) on random inputs over many binades, and
    check the three properties the kernel leans on: rcp results leave the five low mantissa bits zero (the orientation bin
    is stored there), the fused table holds {rsqrt, rcp(rsqrt)/16}, and MIN(rsqrt(M2), 1e10f) saturates exactly for
    bits(M2) <= u_cap."""
    import ctypes as C
    import numpy as np
    rs, info = _table(0); rc, _ = _table(1); fused, _ = _table(3); consts, _ = _table(6, np.uint32)
    kb_rs, kb_rc = info[0], info[1]
    rng = np.random.default_rng(5)
    x = (rng.uniform(1.0, 2.0, 200000) * np.exp2(rng.integers(-60, 60, 200000))).astype(np.float32)

    def sse(name, v):
        out = np.zeros_like(v)
        getattr(port.kcf, name)(v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(len(v)))
        return out

    def emu_rsqrt(v):                                   # table[parity(e)][top mantissa bits] * 2^-(e >> 1), like the kernel
        u = v.view(np.uint32).astype(np.int64)
        e = (u >> 23) - 127
        key = (u & 0x7FFFFF) >> (23 - kb_rs)
        t = rs[((e & 1) << kb_rs) + key].view(np.uint32).astype(np.int64)
        return (t - ((e >> 1) << 23)).astype(np.uint32).view(np.float32)

    def emu_rcp(v):
        u = v.view(np.uint32).astype(np.int64)
        e = (u >> 23) - 127
        t = rc[(u & 0x7FFFFF) >> (23 - kb_rc)].view(np.uint32).astype(np.int64)
        return (t - (e << 23)).astype(np.uint32).view(np.float32)

    assert np.array_equal(emu_rsqrt(x).view(np.uint32), sse("port_sse_rsqrt", x).view(np.uint32))
    assert np.array_equal(emu_rcp(x).view(np.uint32), sse("port_sse_rcp", x).view(np.uint32))
    # fused table: {T, rcp(T) / 16}, and the low five mantissa bits of the second entry are free
    T = fused[0::2].copy(); R16 = fused[1::2].copy()
    assert np.array_equal(T, rs)
    assert np.array_equal(R16, (sse("port_sse_rcp", T) * np.float32(0.0625)).astype(np.float32))
    assert not np.any(R16.view(np.uint32) & 31) and not (int(consts[1]) & 31)
    # saturation threshold: exhaustive over the float patterns around it, plus zero / denormals / ordinary values
    u_cap = int(consts[0])
    u = np.concatenate([np.arange(max(0, u_cap - 200000), u_cap + 200000, dtype=np.uint32),
                        np.array([0, 1, 0x7FFFFF, 0x800000, 0x3F800000], np.uint32)])
    m = sse("port_sse_rsqrt", u.view(np.float32))
    saturates = ~(m < np.float32(1e10))
    assert np.array_equal(saturates, u <= u_cap)


def test_folded_orientation_table_equals_the_step_table():
    """bin2 (wrap 18 -> 0 folded in, both values stored) decodes to the same bin as the step table for every index."""
    import numpy as np
    b1, info = _table(4, np.uint32); b2, _ = _table(5, np.uint32)
    shift, nseg = info[2], info[3]
    ai = np.arange(20020, dtype=np.int64)
    for s in (0, 1):
        e1 = b1[s * nseg + (ai >> shift)].astype(np.int64)
        want = (e1 & 0xFF) - (ai >= (e1 >> 8)).astype(np.int64)
        want[want >= 18] = 0
        e2 = b2[s * nseg + (ai >> shift)].astype(np.int64)
        got = np.where(ai >= (e2 >> 10), (e2 >> 5) & 31, e2 & 31)
        assert np.array_equal(got, want)
