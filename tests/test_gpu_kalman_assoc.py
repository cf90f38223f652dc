"""GPU parity: Kalman kernels, cost matrices and the Munkres kernel vs the oracle (BASELINE configs 2 and 5)."""
import ctypes as C
import numpy as np
import pytest

from synth import BBox, Scene, BBOX_DTYPE, boxes_array, random_boxes, jittered_detections
from gpu_common import require_gpu, mot, rel_err, box_of

pytestmark = pytest.mark.gpu


def test_kalman_state_vs_oracle(oracle):
    """Kalman x and P <= 1e-5 relative (BASELINE.json), int boxes equal; 64 tracks x 200 frames with sub-pixel motion."""
    require_gpu()
    M = mot()
    rng = np.random.default_rng(2)
    n = 64
    ctx = M.Context(1920, 1080, max_tracks=128, kind=M.TRACKER_KALMAN)
    truth = random_boxes(rng, n, 1920, 1080, 40, 120)
    h = ctx.new(truth)
    ohs = [oracle.kal_new(box_of(t)) for t in truth]
    pos = np.stack([truth["l"], truth["t"]], 1).astype(np.float64)
    vel = rng.uniform(-3, 3, size=(n, 2))
    w = (truth["r"] - truth["l"]).astype(np.int32); hh = (truth["b"] - truth["t"]).astype(np.int32)
    for f in range(200):
        out = ctx.predict(h, None, truth)
        for i, oh in enumerate(ohs):
            ob = BBox(); oracle.kal_predict(oh, ob)
            assert (ob.l, ob.t, ob.b, ob.r) == tuple(int(out[i][k]) for k in "ltbr"), (f, i)
        pos += vel + rng.normal(0, 0.5, size=pos.shape)
        z = truth.copy()
        z["l"] = pos[:, 0].astype(np.int32); z["t"] = pos[:, 1].astype(np.int32); z["r"] = z["l"] + w; z["b"] = z["t"] + hh
        ctx.update(h, None, z)
        for i, oh in enumerate(ohs):
            oracle.kal_update(oh, box_of(z[i]))
        if f % 20 == 0 or f == 199:
            for i, oh in enumerate(ohs):
                xo, Po, _ = oracle.kal_state(oh)
                assert rel_err(ctx.state(h[i], "x"), xo) <= 1e-5
                assert rel_err(ctx.state(h[i], "P"), Po) <= 1e-5
    for oh in ohs:
        oracle.kal_delete(oh)
    ctx.close()


def test_kalman_truncation_stress(oracle):
    """trackers/kalman.cpp:112-115 truncates the FP64 state to an int box: >= 10^6 predicts (2048 tracks x 512 frames; moving,
    jittered, and STATIC targets whose state converges onto integer measurements, the case where the last bits decide) must give
    the reference's int boxes.  kalman.cu is built with --fmad=false for this."""
    require_gpu()
    M = mot()
    rng = np.random.default_rng(11)
    n, nf = 2048, 512
    ctx = M.Context(1920, 1080, max_tracks=n, kind=M.TRACKER_KALMAN)
    init = random_boxes(rng, n, 1920, 1080, 40, 120)
    w = (init["r"] - init["l"]).astype(np.int32); hh = (init["b"] - init["t"]).astype(np.int32)
    pos = np.stack([init["l"], init["t"]], 1).astype(np.float64)
    vel = rng.uniform(-3, 3, size=(n, 2)); vel[: n // 4] = 0.0                      # a quarter of the targets never move ...
    noise = np.full(n, 0.5); noise[: n // 8] = 0.0                                  # ... and half of those are measured without jitter
    meas = np.zeros((nf, n), BBOX_DTYPE)
    for f in range(nf):
        pos += vel + rng.normal(0, 1.0, size=pos.shape) * noise[:, None]
        z = init.copy()
        z["l"] = pos[:, 0].astype(np.int32); z["t"] = pos[:, 1].astype(np.int32); z["r"] = z["l"] + w; z["b"] = z["t"] + hh
        meas[f] = z
    want = np.zeros((nf, n), BBOX_DTYPE)
    getattr(oracle.kal, oracle.p + "kal_run")(n, nf, init.ctypes.data_as(C.c_void_p), meas.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p))
    h = ctx.new(init)
    bad = 0
    for f in range(nf):
        out = ctx.predict(h, None, init)
        for k in "ltbr":
            bad += int(np.count_nonzero(out[k] != want[f][k]))
        ctx.update(h, None, meas[f])
    assert bad == 0, "%d of %d predicted coordinates differ from the reference's truncated state" % (bad, 4 * n * nf)
    ctx.close()


def _oracle_cost(oracle, trk, det, mode, W):
    T, D = len(trk), len(det)
    nr, nc = (T, D) if T < D else (D, T)
    d = np.zeros(nr * nc, np.float64)
    oracle.kcf.port_cost_matrix(d.ctypes.data_as(C.c_void_p), trk.ctypes.data_as(C.c_void_p), T,
                                det.ctypes.data_as(C.c_void_p), D, mode, C.c_double(1.0 / W))
    return d.reshape(nc, nr).T


@pytest.mark.parametrize("T,D,mode", [(64, 64, 0), (64, 64, 1), (58, 64, 1), (64, 50, 0), (1, 7, 0), (9, 1, 1), (128, 128, 1), (200, 160, 0), (256, 256, 1)])
def test_cost_and_assignment_vs_oracle(oracle, T, D, mode):
    """Cost matrices bit-exact vs the host restatement; assignment vectors bit-exact vs assignmentoptimal on the same matrix."""
    require_gpu()
    M = mot()
    rng = np.random.default_rng(T * 1000 + D * 7 + mode)
    ctx = M.Context(1920, 1080, max_tracks=4, kind=M.TRACKER_KALMAN)
    trks, dets = [], []
    for rep in range(6):
        trk = random_boxes(rng, T, 1920, 1080)
        if D <= T:
            det = jittered_detections(rng, trk, 1920, 1080)[:D]
        else:
            det = np.concatenate([jittered_detections(rng, trk, 1920, 1080), random_boxes(rng, D - T, 1920, 1080)])
        trks.append(trk); dets.append(np.ascontiguousarray(det))
    assigns, costs, dists = ctx.associate(trks, dets, cost_mode=mode, want_dist=True)
    for m in range(len(trks)):
        d_or = _oracle_cost(oracle, trks[m], dets[m], mode, 1920)
        assert np.array_equal(dists[m], d_or), "cost matrix %d must be bit-exact" % m
        a_or, c_or = oracle.assign(d_or)
        assert np.array_equal(assigns[m], a_or), "assignment %d must be bit-exact" % m
        assert costs[m] == c_or
    ctx.close()


def test_assignment_ties_degenerate_and_rectangular(oracle):
    require_gpu()
    M = mot()
    rng = np.random.default_rng(12)
    ctx = M.Context(640, 480, max_tracks=4, kind=M.TRACKER_KALMAN)
    mats = [np.zeros((6, 6)), np.ones((5, 9)), np.ones((9, 5)), np.zeros((1, 1)), np.full((3, 1), 2.5),
            rng.integers(0, 3, size=(40, 40)).astype(float), rng.integers(0, 4, size=(30, 45)).astype(float),
            rng.integers(0, 4, size=(45, 30)).astype(float), np.round(rng.random((64, 64)), 1),
            np.round(rng.random((100, 70)), 2), rng.integers(0, 2, size=(33, 65)).astype(float)]
    assigns, costs = ctx.assign(mats)
    for m, d in enumerate(mats):
        a_or, c_or = oracle.assign(d)
        assert np.array_equal(assigns[m], a_or), (m, d.shape)
        assert costs[m] == c_or
    ctx.close()


def test_config5_512_subset(oracle):
    """BASELINE config 5, small subset: 512x512 and 448x512 problems, both cost modes, bit-exact vs assignmentoptimal."""
    require_gpu()
    M = mot()
    rng = np.random.default_rng(0x5EED0500)
    ctx = M.Context(1920, 1080, max_tracks=4, kind=M.TRACKER_KALMAN)
    for (T, D, mode) in [(512, 512, 1), (448, 512, 0)]:
        trk = random_boxes(rng, T, 1920, 1080)
        det = jittered_detections(rng, trk, 1920, 1080)
        if D > T:
            det = np.concatenate([det, random_boxes(rng, D - T, 1920, 1080)])
        det = np.ascontiguousarray(det[:D])
        assigns, costs, dists = ctx.associate([trk, trk], [det, det[::-1].copy()], cost_mode=mode, want_dist=True)
        for m in range(2):
            a_or, c_or = oracle.assign(dists[m])
            assert np.array_equal(assigns[m], a_or)
            assert costs[m] == c_or
    ctx.close()


def test_nonfinite_costs_terminate():
    """The reference never returns on -inf costs; the kernel must, flagging the problem."""
    require_gpu()
    M = mot()
    ctx = M.Context(640, 480, max_tracks=4, kind=M.TRACKER_KALMAN)
    d = np.ones((8, 8)); d[3, 4] = -np.inf
    assigns, costs = ctx.assign([d, np.eye(4)])
    assert len(assigns[0]) == 8
    assert list(assigns[1]) == list(M.Context.assign(ctx, [np.eye(4)])[0][0])
    ctx.close()
