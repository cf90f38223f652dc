#!/usr/bin/env python
"""bench.py -- KCF track-updates/sec on B200 (BASELINE.json metric), beside the reference CPU path.

One "step" = one pass of the hot path over one batch: every stream gets a frame, every track gets one
tracker_predict (+ the td.cpp clamp) and one tracker_update with its own predicted box (the reference's unassigned
branch, top/td.cpp:550-582), crop + gray + resize included.  Workload = BASELINE config 4 per GPU:
64 independent 1080p streams x 128 KCF tracks with 128x128-px windows (32x32 cells).  Streams are independent, so
ranks share nothing: no data-path collective, weak scaling (every GPU gets its own 64 streams).

  python bench.py [--gpus N] [--steps K] [--warmup W]            the CUDA path (this repo), BASELINE config 4 (the headline)
  python bench.py --impl reference ...                            the reference's CPU path on the host cores
  python bench.py --config C1|C2|C3|C5                            the other BASELINE configs (one JSON line each, CPU oracle beside;
                                                                  single GPU; C4 is the default and the only multi-GPU one)
The default line also carries `strong` (config 4 as BASELINE words it: 64 streams IN TOTAL sharded over the N GPUs), `full_loop`
(the device-resident frame loop: the same tracks with association + lifecycle every frame) and, in `e2e`, the all-ranks-concurrent
H2D rate and the NUMA placement that explain the end-to-end number.

`value`     device-timed (CUDA events on the launching stream), frames + boxes + models resident in HBM
`e2e`       the same metric through the host-array C ABI: frames copied from pinned host memory every step,
            boxes H2D/D2H every step (mot_frame_upload + mot_predict_batch + mot_update_batch)
`roofline`  algorithmic HBM bytes of the dominant kernel / its measured launch time, vs MEASURED_PEAKS.json
`cpu_baseline`  the oracle (compiled reference if oracle/_ref is present, else the C port) on the host cores
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "multiple-object-tracking_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

W, H, WIN, TSIZE = 1920, 1080, 128, 48
HR = WC = WIN // 4
S = WC * (HR // 2 + 1)
# algorithmic HBM bytes per track (SURVEY.md 8d / BASELINE.md 3): ROI u8 BGR once per kernel, model read (predict) and
# read+write (update), alpha likewise, Re(yf) once, boxes in/out
B_PREDICT = WIN * WIN * 3 + 31 * S * 8 + S * 4 + 24 + 24
B_UPDATE = WIN * WIN * 3 + 2 * 31 * S * 8 + 2 * S * 4 + S * 4 + 24
B_PAIR = B_PREDICT + B_UPDATE + 0          # 511,816 with both box directions counted; BASELINE.md quotes 511,792


def make_streams(n_streams, n_tracks, ring, seed0):
    """Host-side synthetic streams: `ring` frames per stream (targets move between them) + the initial windows."""
    from synth import Scene
    frames = np.zeros((n_streams, ring, H, W, 3), np.uint8)
    boxes = []
    bgs = {}
    for s in range(n_streams):
        key = s % 4                                   # four distinct backgrounds / texture sets, re-used across streams
        if key not in bgs:
            bgs[key] = Scene(seed0 + key, W, H, n_tracks, tsize=TSIZE, win=WIN)
        base = bgs[key]
        sc = Scene.__new__(Scene)
        sc.__dict__.update(base.__dict__)
        sc.rng = np.random.default_rng(seed0 + 1000 + s)
        sc.pos = np.clip(base.pos + sc.rng.integers(-2, 3, size=base.pos.shape), base.lo, base.hi)
        sc.vel = sc.rng.uniform(-3, 3, size=base.vel.shape)
        boxes.append(sc.windows())
        for r in range(ring):
            frames[s, r] = sc.render()
            sc.step()
    return frames, boxes


def numa_info(gpu_index):
    """Where the GPU and this process's pinned memory live: the NUMA node of the GPU's PCIe function, the node(s) of the CPUs this
    process may run on, the number of nodes of the box."""
    out = {}
    try:
        import torch
        bus = torch.cuda.get_device_properties(gpu_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(bus, "pci_domain_id", 0), bus.pci_bus_id, bus.pci_device_id)
        out["gpu_pci"] = bdf
        out["gpu_numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception as e:
        out["gpu_numa_node"] = "unknown (%s)" % type(e).__name__
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        out["numa_nodes"] = len(nodes)
        mine = os.sched_getaffinity(0)
        on = []
        for d in nodes:
            cpus = set()
            for part in open("/sys/devices/system/node/%s/cpulist" % d).read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            if cpus & mine:
                on.append(int(d[4:]))
        out["process_cpu_nodes"] = on
    except Exception as e:
        out["numa_nodes"] = "unknown (%s)" % type(e).__name__
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ======================================================================================================= reference arm
def cpu_reference(n_threads, tracks_per_thread, warmup, steps, want_mkl=True):
    """The reference's CPU path for this metric: per host thread an independent set of KCF trackers driven through the
    reference frame loop (oracle td loop over the compiled tracker_predict/tracker_update + rgb2Gray/bilinear
    restatement), every track assigned to its own box each frame (so: predict + update per track per frame)."""
    if want_mkl:
        try:
            import torch
            lib = os.path.join(os.path.dirname(torch.__file__), "lib", "libtorch_cpu.so")
            if os.path.exists(lib):
                os.environ.setdefault("REF_FFT_PROVIDER", "mkl")
                os.environ.setdefault("REF_FFT_MKL_LIB", lib)
                os.environ.setdefault("MKL_NUM_THREADS", "1")
                os.environ.setdefault("OMP_NUM_THREADS", "1")
        except Exception:
            pass
    import oraclelib
    from synth import Scene
    oraclelib.build_port()
    kind = oraclelib.best()
    orc = oraclelib.Oracle(kind)
    provider = orc.kcf.ref_fft_provider().decode() if kind == "ref" else "dft64(port_fft, double precision)"
    sc = Scene(0xC4, W, H, tracks_per_thread, tsize=TSIZE, win=WIN)
    ring = []
    first = sc.windows()
    for _ in range(2):
        ring.append(sc.render()); sc.step()
    tds = [orc.td_new("kcf", W, H, tracks_per_thread + 8, 0) for _ in range(n_threads)]
    for td in tds:
        td.step(ring[0], first)                       # spawn + first update (tracker_new excluded from the timing)

    def one(td, k):
        dets = td.tracks()["boxes"]
        td.step(ring[k & 1], np.ascontiguousarray(dets))

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(n_threads) as ex:
        for k in range(warmup):
            list(ex.map(lambda td: one(td, k + 1), tds))
        t0 = time.perf_counter()
        for k in range(steps):
            list(ex.map(lambda td: one(td, warmup + k + 1), tds))
        dt = time.perf_counter() - t0
    done = n_threads * tracks_per_thread * steps
    for td in tds:
        td.close()
    return {"value": done / dt, "unit": "track-updates/s", "cores": n_threads,
            "kind": "reference" if kind == "ref" else "port",
            "sample": "%d threads x %d KCF tracks (128x128 px) x %d frames of one 1080p stream; FFT provider %s; %.1f s" %
                      (n_threads, tracks_per_thread, steps, provider, dt)}, dt


def cpu_single_core(frames=12):
    """The "host-core" figure of the north-star target: one thread, 32 KCF tracks, same loop."""
    cb, _ = cpu_reference(1, 32, 1, frames)
    return cb["value"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tpt = 32
    cb, dt = cpu_reference(cores, tpt, args.warmup, args.steps)
    line = {"metric": "KCF track-updates/sec", "value": cb["value"], "unit": "track-updates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "C4 bounded sample: %d host threads x %d KCF tracks (128x128 px windows), 1080p" % (cores, tpt)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "track-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ======================================================================================================= CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import mot_b200 as M

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    NS, NT, RING = args.streams, args.tracks, 2
    n = NS * NT
    frames_h, boxes0 = make_streams(NS, NT, RING, 0xC400 + 97 * rank)
    ctx = M.Context(W, H, max_tracks=n, n_frame_slots=NS * RING, kind=M.TRACKER_KCF, device=local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    # ---- resident state -------------------------------------------------------------------------------------------
    frames_pin = torch.from_numpy(frames_h.reshape(NS * RING, H, W, 3)).pin_memory()
    frames_d = frames_pin.to(dev, non_blocking=False)
    for i in range(NS * RING):
        ctx.bind_device(i, frames_d[i].data_ptr(), W * 3)
    all_boxes = np.ascontiguousarray(np.concatenate(boxes0))
    handles = ctx.new(all_boxes)
    stream_of = np.repeat(np.arange(NS, dtype=np.int32), NT)
    slots_ring = [np.ascontiguousarray(stream_of * RING + r) for r in range(RING)]
    ctx.update(handles, slots_ring[0], all_boxes)                         # first update (tracker_new's companion, td.cpp:629-641)
    d_handles = torch.from_numpy(handles).to(dev)
    d_frames = [torch.from_numpy(sr).to(dev) for sr in slots_ring]
    d_boxes = torch.from_numpy(all_boxes.view(np.uint8).reshape(n, 24)).to(dev)
    torch.cuda.synchronize()

    def step(k):
        fr = d_frames[(k + 1) % RING].data_ptr()
        ctx.predict_dev(n, d_handles.data_ptr(), fr, d_boxes.data_ptr(), clamp=1)
        ctx.update_dev(n, d_handles.data_ptr(), fr, d_boxes.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()

    with torch.cuda.stream(stream):
        clocks = ClockSampler(local) if rank == 0 else None      # sampled under load: warm-up + timed region
        for k in range(args.warmup):
            step(k)
        stream.synchronize()
        if clocks:                                               # nvidia-smi needs a moment to start: keep the GPU busy meanwhile
            t_end = time.time() + 0.6
            k = args.warmup
            while time.time() < t_end or (k - args.warmup) % RING:
                step(k); k += 1
                stream.synchronize()
        barrier(); torch.cuda.synchronize()
        l0 = ctx.launches()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
        evs[0].record(stream)
        for k in range(args.steps):
            fr = d_frames[(args.warmup + k + 1) % RING].data_ptr()
            ctx.predict_dev(n, d_handles.data_ptr(), fr, d_boxes.data_ptr(), clamp=1)
            evs[2 * k + 1].record(stream)
            ctx.update_dev(n, d_handles.data_ptr(), fr, d_boxes.data_ptr())
            evs[2 * k + 2].record(stream)
        stream.synchronize(); torch.cuda.synchronize()
        barrier()
        clk = clocks.stop() if clocks else None
        launches = ctx.launches() - l0
    total_ms = evs[0].elapsed_time(evs[-1])
    t_pred = sum(evs[2 * k].elapsed_time(evs[2 * k + 1]) for k in range(args.steps)) / args.steps
    t_upd = sum(evs[2 * k + 1].elapsed_time(evs[2 * k + 2]) for k in range(args.steps)) / args.steps
    from mot_b200.shard import max_over_ranks
    total_ms = max_over_ranks(total_ms, dev)              # device time of the step = max over ranks
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- sanity: the trackers are still on their targets (the timed work was real) ------------------------------------
    final = d_boxes.cpu().numpy().view(M.BBOX_DTYPE).reshape(n)
    drift = float(np.abs((final["l"] + final["r"]) / 2.0 - (all_boxes["l"] + all_boxes["r"]) / 2.0).max())

    # ---- strong scaling (BASELINE config 4 as worded: 64 streams in total, sharded s mod G): this rank's share of the streams -------
    strong = None
    if NS % world == 0:
        ns_s = NS // world
        n_s = ns_s * NT
        with torch.cuda.stream(stream):
            for k in range(2):
                fr = d_frames[k % RING].data_ptr()
                ctx.predict_dev(n_s, d_handles.data_ptr(), fr, d_boxes.data_ptr(), clamp=1); ctx.update_dev(n_s, d_handles.data_ptr(), fr, d_boxes.data_ptr())
            stream.synchronize(); barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(args.steps):
                fr = d_frames[k % RING].data_ptr()
                ctx.predict_dev(n_s, d_handles.data_ptr(), fr, d_boxes.data_ptr(), clamp=1); ctx.update_dev(n_s, d_handles.data_ptr(), fr, d_boxes.data_ptr())
            e1.record(stream); stream.synchronize()
        ms_s = max_over_ranks(e0.elapsed_time(e1), dev)
        strong = {"value": world * n_s * args.steps / (ms_s * 1e-3), "unit": "track-updates/s", "streams_total": NS, "streams_per_gpu": ns_s,
                  "tracks_per_gpu": n_s, "ms_per_step": ms_s / args.steps, "scaling": "strong"}

    # ---- the whole frame loop on the device (association + bookkeeping + lifecycle every frame; detections = the tracks' own boxes) ----
    full_loop = None
    if not args.no_loop:
        ctx3 = M.Context(W, H, max_tracks=n, n_frame_slots=NS * RING, kind=M.TRACKER_KCF, device=local)
        ctx3.set_stream(stream.cuda_stream)
        for i in range(NS * RING):
            ctx3.bind_device(i, frames_d[i].data_ptr(), W * 3)
        loop = M.DeviceLoop(ctx3, NS, cap=NT, max_det=NT, cost_mode=0)
        loop.kcf_windows([(WIN, WIN)])
        det_buf = np.zeros((NS, NT), M.BBOX_DTYPE)
        for s_ in range(NS):
            det_buf[s_] = boxes0[s_]
        d_det = torch.from_numpy(det_buf.view(np.uint8).reshape(NS, NT * 24)).to(dev)
        d_nd = torch.full((NS,), NT, dtype=torch.int32, device=dev)
        with torch.cuda.stream(stream):
            # the frame ring is [stream][ring]; the loop reads slot base + s, so bind ring frame r of stream s to slot r * NS + s
            for s_ in range(NS):
                for r in range(RING):
                    ctx3.bind_device(r * NS + s_, frames_d[s_ * RING + r].data_ptr(), W * 3)
            for k in range(3):
                loop.frame_base((k % RING) * NS); loop.step_dev(d_det.data_ptr(), d_nd.data_ptr())
            stream.synchronize(); barrier()
            l3 = ctx3.launches()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            kl = max(3, min(args.steps, 10))
            for k in range(kl):
                loop.frame_base(((k + 1) % RING) * NS); loop.step_dev(d_det.data_ptr(), d_nd.data_ptr())
            e1.record(stream); stream.synchronize()
        ms_l = max_over_ranks(e0.elapsed_time(e1), dev)
        ntr = sum(len(loop.tracks(s_)["tid"]) for s_ in (0, NS - 1))
        full_loop = {"value": world * n * kl / (ms_l * 1e-3), "unit": "track-updates/s", "ms_per_step": ms_l / kl, "launches_per_step": (ctx3.launches() - l3) / kl,
                     "what": "mot_tdd_step_dev: job lists, fused predict, 64 x (128x128) cost matrices + Munkres, scatter, fused update, lifecycle",
                     "tracks_alive_in_first_and_last_stream": ntr}
        loop.close(); ctx3.close()

    # ---- e2e: host arrays through the C ABI, frames from pinned host memory every step ------------------------------------
    e2e = None
    if not args.no_e2e:
        # Two frame slots per stream: while the kernels of step k run, the frames of step k+1 are already crossing PCIe on
        # the context's copy stream (mot_frame_upload); every step still uploads its 64 frames and reads its boxes back.
        ctx2 = M.Context(W, H, max_tracks=n, n_frame_slots=2 * NS, kind=M.TRACKER_KCF, device=local)
        h2 = ctx2.new(all_boxes)
        fs2 = [np.ascontiguousarray(stream_of * 2 + par) for par in range(2)]
        frames_np = [frames_pin[i].numpy() for i in range(NS * RING)]

        def upload(k):                       # frames of step k -> slot parity k & 1
            r = (k + 1) % RING
            for s_ in range(NS):
                ctx2.upload(s_ * 2 + (k & 1), frames_np[s_ * RING + r])

        for s_ in range(NS):
            ctx2.upload(s_ * 2, frames_np[s_ * RING])
        ctx2.update(h2, fs2[0], all_boxes)
        hb = all_boxes.copy()

        def e2e_step(k):
            nonlocal hb
            upload(k + 1)                    # next step's frames: overlaps with this step's kernels
            hb = ctx2.predict(h2, fs2[k & 1], hb, clamp=1)
            ctx2.update(h2, fs2[k & 1], hb)

        upload(0); upload(1)                 # both slot sets exist (the first upload of a slot allocates it)
        ctx2.sync()
        t0 = time.perf_counter(); upload(1); ctx2.sync(); h2d_s = time.perf_counter() - t0      # PCIe alone (this rank; the others may be busy)
        # every rank uploads at the same moment: what one GPU gets when all of the box's GPUs pull frames from the host at once
        barrier()
        t0 = time.perf_counter(); upload(0); upload(1); ctx2.sync(); h2d_all_s = (time.perf_counter() - t0) / 2
        h2d_all_s = max_over_ranks(h2d_all_s, dev)
        barrier()
        upload(0)
        for k in range(2):
            e2e_step(k)
        ctx2.sync(); barrier()
        ke = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for k in range(ke):
            e2e_step(2 + k)
        ctx2.sync()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt, dev)
        e2e = {"value": world * n * ke / dt, "unit": "track-updates/s", "h2d_bytes_per_step": NS * H * W * 3 + 2 * n * (24 + 8),
               "d2h_bytes_per_step": n * 24, "steps": ke, "ms_per_step": 1e3 * dt / ke,
               "h2d_alone_ms": 1e3 * h2d_s, "h2d_alone_gbs": NS * H * W * 3 / h2d_s / 1e9,
               "h2d_all_ranks_gbs": NS * H * W * 3 / h2d_all_s / 1e9, "h2d_all_ranks_ms": 1e3 * h2d_all_s,
               "h2d_bound": "a step cannot be faster than its frame upload: %.2f ms at the all-ranks rate vs %.2f ms measured" % (1e3 * h2d_all_s, 1e3 * dt / ke),
               "numa": numa_info(local)}
        ctx2.close()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dom = "update" if t_upd >= t_pred else "predict"
        bytes_dom = (B_UPDATE if dom == "update" else B_PREDICT) * n
        achieved = bytes_dom / (max(t_upd, t_pred) * 1e-3) / 1e9
        traffic = None
        try:            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("jobs_per_launch") == n:
                traffic = tj[dom]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "kcf_fused_kernel<32,32,%s>" % dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                    "ms_predict": t_pred, "ms_update": t_upd, "algorithmic_bytes_per_track": {"predict": B_PREDICT, "update": B_UPDATE},
                    "step_frac": (B_PAIR * n / ((t_pred + t_upd) * 1e-3) / 1e9) / peak}
        cb = None
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_reference(os.cpu_count() or 1, 32, 2, 24)
            cb["single_core"] = {"value": cpu_single_core(), "unit": "track-updates/s", "cores": 1,
                                 "note": "the host-core figure of the north-star target (>= 100x on one B200)"}
            cb["times"] = "td.step of the oracle loop (predict + association of 32 tracks + update + lifecycle); the association is ~1 % of it"
        line = {"metric": "KCF track-updates/sec", "value": value, "unit": "track-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C4: %d independent 1080p streams x %d KCF tracks (128x128 px windows, 32x32 cells) per GPU" % (NS, NT),
                           "streams_per_gpu": NS, "tracks_per_stream": NT, "frame": "1920x1080 BGR u8", "parallelism": "streams sharded, no collective",
                           "l2": "working set per step (%.0f MB of models + frames) exceeds the 126 MB L2" % ((31 * S * 8 + S * 4) * n / 1e6 + NS * H * W * 3 / 1e6),
                           "max_center_drift_px": drift,
                           "reference_arm": "--impl reference times td.step of the compiled reference (predict + update per track, plus the association and lifecycle of its 32-track loops: ~1 % extra work on the reference's side)"},
                "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
                "strong": strong, "full_loop": full_loop}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_other_config(args):
    """BASELINE configs 1, 2, 3, 5 (bench_assoc.py holds the workloads): one JSON line in the bench schema, the CPU oracle timed beside."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import bench_assoc as B
    B.use_mkl_fft_for_the_cpu_side()
    if args.impl == "reference":
        print(json.dumps({"impl": "reference", "config": {"workload": args.config}, "note": "the CPU oracle is timed inside the --config line itself (cpu_baseline)"}))
        return
    if args.config == "C1":
        r = B.config1(args)
        line = {"metric": "single-target KCF frames/sec (host-array API: frame upload + mot_track_batch per frame)", "value": r["gpu_frames_per_s_track_call"], "unit": "frames/s",
                "cpu_baseline": {"value": r["cpu_frames_per_s_one_core"], "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "the same 300-frame sequence, FFT " + r["cpu_fft"]}}
    elif args.config == "C2":
        r = B.config2(args)
        line = {"metric": "Kalman + Hungarian frames/sec, 64 tracks x 64 detections, one stream (device-resident loop)", "value": r["device_resident_1_stream"]["gpu_stream_frames_per_s"],
                "unit": "frames/s", "cpu_baseline": {"value": r["ref_centroid"]["cpu_frames_per_s_one_core"], "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "the same 300 frames through the oracle loop"}}
    elif args.config == "C3":
        r = B.config3(args)
        line = {"metric": "multi-target KCF frames/sec, 256 tracks in one 1080p stream, whole frame loop (device-resident)", "value": r["gpu_frames_per_s_device_loop"], "unit": "frames/s",
                "cpu_baseline": {"value": r["cpu_frames_per_s_one_core"], "unit": "frames/s", "cores": 1, "kind": "reference", "sample": "the first 8 frames of the same sequence through the oracle loop (track tables compared there)"}}
    else:
        r = B.config5(args)
        line = {"metric": "association problems/sec, %d x (%dx%d) cost matrix + Munkres (REF_CENTROID costs)" % (args.matrices, args.dim, args.dim), "value": r["ref_centroid"]["gpu_matrices_per_s"],
                "unit": "matrices/s", "cpu_baseline": {"value": 1.0 / r["ref_centroid"]["cpu_s_per_matrix_one_core"], "unit": "matrices/s", "cores": 1, "kind": "reference",
                                                       "sample": "%d of the matrices through assignmentoptimal, assignments compared bit for bit" % r["ref_centroid"]["checked"]}}
    line.update({"n_gpus": 1, "higher_is_better": True, "scaling": "none (single GPU config)", "vs_baseline": None, "data": "synthetic", "config": {"workload": r["config"]}, "details": r})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--tracks", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-loop", action="store_true")
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--matrices", type=int, default=1024)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--cpu-matrices", type=int, default=2)
    args = ap.parse_args()
    if args.config != "C4":
        return run_other_config(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
