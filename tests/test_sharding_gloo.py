"""CPU, world_size 2, gloo: the slab-parallel multi-GPU path with the oracle
standing in for the kernels -- each rank filters its block-row range of the same window,
rows are gathered on rank 0 and FRAME_DIFF is all-reduced; the result equals the
single-process run bit for bit (integer sums are order independent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import conftest
    import _clips
    import _oracle
    import _params
    conftest.load_package()
    import importlib
    sh = importlib.import_module("aom_av1_psy_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, N = 160, 96, 3
    frames = _clips.moving_texture(W, H, N, 8)
    p = _params.tf_params(W, H, N)
    mb_rows = (H + 31) // 32
    b, e = sh.slab_rows(mb_rows, world, rank)
    o = _oracle.OracleFilter(p, frames)
    r = o.run(rows=(b, e))
    pad = sh.max_slab_rows(mb_rows, world)
    gathered = []
    for pl in range(3):
        bh = 32 >> (1 if pl else 0)
        slab = np.zeros((pad * bh, r["out"][pl].shape[1]), np.int32)
        slab[: (e - b) * bh] = r["out"][pl][b * bh:e * bh]
        t = torch.from_numpy(slab)
        lst = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, lst, dst=0)
        if rank == 0:
            gathered.append(sh.merge_slabs([x.numpy() for x in lst], mb_rows, world, bh))
    d = torch.from_numpy(r["diff"].copy())
    dist.all_reduce(d)
    if rank == 0:
        full = _oracle.OracleFilter(p, frames).run()
        ok = all((g == f.astype(np.int32)).all() for g, f in zip(gathered, full["out"]))
        ok = ok and (d.numpy() == full["diff"]).all()
        q.put(bool(ok))
    dist.destroy_process_group()


def test_slab_mode_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
