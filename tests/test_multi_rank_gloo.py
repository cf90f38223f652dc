"""N > 1 host logic on CPU: world_size-2 gloo.  Streams shard round-robin with no data-path collective; the only
collectives are the barrier and the MAX / SUM reductions of timings and unit counts that bench.py reports."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_streams, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "multiple-object-tracking_b200"))
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mot_b200.shard import streams_of_rank, max_over_ranks, total_over_ranks
    mine = streams_of_rank(n_streams, rank, world)
    # each rank runs its own streams through the oracle's Kalman frame loop (CPU stand-in for the per-GPU work)
    import numpy as np
    import oraclelib
    from synth import Scene
    orc = oraclelib.Oracle("port")
    units = 0
    sig = []
    for s in mine:
        sc = Scene(5000 + s, 640, 480, 6, tsize=32, win=64)
        td = orc.td_new("kal", 640, 480, 32, 0)
        for f in range(5):
            sc.step(); td.step(None, sc.windows(jitter=1)); units += len(td.tracks()["tid"])
        sig.append((s, td.tracks()["boxes"]["l"].tolist()))
        td.close()
    dist.barrier()
    t = max_over_ranks(1.0 + rank)                 # pretend rank r took 1+r ms: the step time is the max
    total = total_over_ranks(units)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, sig))
    if rank == 0:
        q.put((t, total, gathered))
    dist.destroy_process_group()


def test_two_rank_stream_sharding():
    import oraclelib
    oraclelib.build_port()
    world, n_streams = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0, "step time = max over ranks"
    owned = sorted(s for mine, _ in gathered for s in mine)
    assert owned == list(range(n_streams)), "every stream owned by exactly one rank"
    assert gathered[0][0] == [0, 2, 4, 6] and gathered[1][0] == [1, 3, 5]
    # sharding does not change results: re-run every stream in this process and compare
    import numpy as np
    from synth import Scene
    orc = oraclelib.Oracle("port")
    units = 0
    for mine, sig in gathered:
        for s, boxes_l in sig:
            sc = Scene(5000 + s, 640, 480, 6, tsize=32, win=64)
            td = orc.td_new("kal", 640, 480, 32, 0)
            for f in range(5):
                sc.step(); td.step(None, sc.windows(jitter=1)); units += len(td.tracks()["tid"])
            assert td.tracks()["boxes"]["l"].tolist() == boxes_l
            td.close()
    assert total == units
