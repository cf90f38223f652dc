"""Host logic of the any-size KCF kernel (csrc/kcf_any.cuh, mot_capi.cu: any_plan), no GPU needed: the shared-memory plan of every cell
grid a 1080p frame can produce, and the radix plans of every transform length."""
import ctypes as C

import numpy as np


def plan(hr, wc):
    import mot_b200
    out = (C.c_int * 16)(); rad = (C.c_int * 14)()
    rc = mot_b200.lib().mot_debug_any_plan(hr, wc, out, rad)
    assert rc == 0
    keys = ["ok", "strips", "smem", "tc", "xw", "cs", "xbuf", "aF", "bF", "scratch", "ctas", "threads", "nf_r", "nf_c"]
    d = dict(zip(keys, list(out)))
    d["rad_r"] = [r for r in list(rad)[:7] if r]; d["rad_c"] = [r for r in list(rad)[7:] if r]
    return d


def test_radix_plans_factor_every_length():
    for n in range(2, 271):
        p = plan(n, 2)
        assert int(np.prod(p["rad_r"])) == n and len(p["rad_r"]) == p["nf_r"] <= 7, (n, p["rad_r"])
        # register butterflies exist for 2..10; any other radix must be a prime (evaluated from the definition)
        for r in p["rad_r"]:
            assert r <= 10 or all(r % q for q in range(2, int(r ** 0.5) + 1)), (n, r)
        # fewest passes: nothing that two allowed radices can do takes three
        if any(n % a == 0 and 2 <= n // a <= 10 for a in range(2, 11)):
            assert p["nf_r"] <= 2, (n, p["rad_r"])
    assert plan(30, 40)["rad_r"] == [6, 5] or plan(30, 40)["rad_r"] == [10, 3]
    assert plan(30, 40)["rad_c"] in ([8, 5], [10, 4])
    assert plan(37, 22)["rad_r"] == [37]


def test_shared_memory_plans():
    budget = 227 * 1024 - 2048
    for hr, wc in [(2, 2), (9, 13), (13, 9), (25, 15), (20, 12), (30, 40), (40, 30), (37, 22), (50, 22), (32, 46), (75, 37), (50, 50), (22, 82), (60, 60), (80, 70), (90, 90), (120, 50)]:
        p = plan(hr, wc)
        assert p["ok"], (hr, wc)
        assert 0 < p["smem"] <= budget and p["smem"] % 16 == 0, (hr, wc, p)
        assert 1 <= p["tc"] <= 31 and p["aF"] >= 4 * p["xbuf"], (hr, wc, p)
        assert p["threads"] * p["ctas"] <= 1024 and p["threads"] % 32 == 0
        if hr * wc <= 1200:
            assert not p["strips"] and p["scratch"] == 0, (hr, wc)
        if hr * wc >= 2000:
            assert p["strips"] and 1 <= p["cs"] <= wc and p["scratch"] >= 18 * hr * wc * 4, (hr, wc, p)
    # small windows share an SM: four CTAs for the smallest, two in between, one for the named-shape-sized ones
    assert plan(13, 9)["ctas"] == 4 and plan(25, 15)["ctas"] == 2 and plan(30, 40)["ctas"] == 1
    # a window no CTA can hold even in strips is refused (the unfused path serves it): the whole 1080p frame, or 100 x 100 cells
    assert not plan(270, 480)["ok"] and not plan(100, 100)["ok"]


def test_padding_round_trip_of_the_restatement():
    """kcf_pad_box / kcf_unpad_box (csrc/mot_internal.h) as restated in tests/kcf_ext_numpy.py: window -> target is the identity,
    the window is centred on the target (to within half a pixel) and has int(size * p) pixels per side."""
    from kcf_ext_numpy import pad_box, unpad_box
    rng = np.random.default_rng(1)
    for _ in range(2000):
        l, t = int(rng.integers(-50, 2000)), int(rng.integers(-50, 1100))
        w, h = int(rng.integers(8, 400)), int(rng.integers(8, 400))
        p = float(rng.choice([1.5, 2.0, 2.5, 3.0]))
        tgt = (l, t, t + h - 1, l + w - 1)
        wl, wt, wb, wr = pad_box(*tgt, p)
        assert wr - wl + 1 == int(np.float32(w) * np.float32(p)) and wb - wt + 1 == int(np.float32(h) * np.float32(p))
        assert abs((wl + wr) - (l + l + w - 1)) <= 1 and abs((wt + wb) - (t + t + h - 1)) <= 1
        assert unpad_box(wl, wt, wb, wr, w, h) == tgt
