"""CPU, world_size 2, gloo: the slab-parallel multi-GPU driver (aom_av1_psy_b200.sharding.SlabWindow, the
code bench.py's slab mode and the NCCL test run on GPUs) with the oracle standing in for the kernels -- each
rank filters its block-row range of the same window, SlabWindow.gather() collects the slabs on rank 0 and
all-reduces FRAME_DIFF, SlabWindow.assemble() rebuilds the planes; the result equals the single-process run
bit for bit (integer sums are order independent)."""
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
    slab = sh.SlabWindow(mb_rows, world, rank)
    ps = slab.params(p)
    assert (ps["out_row_begin"], ps["out_row_end"]) == sh.slab_rows(mb_rows, world, rank)
    o = _oracle.OracleFilter(p, frames)
    r = o.run(rows=(slab.begin, slab.end))
    mine, pitch = [], []
    for pl, bh in enumerate(slab.block_h):
        width = r["out"][pl].shape[1]
        buf = np.zeros((slab.pad_rows * bh, width), np.int32)
        buf[: (slab.end - slab.begin) * bh] = r["out"][pl][slab.begin * bh:slab.end * bh]
        mine.append(torch.from_numpy(buf).reshape(-1))
        pitch.append(width)
    gathered, d = slab.gather(mine, torch.from_numpy(r["diff"].copy()), dist, torch)
    if rank == 0:
        planes = slab.assemble(gathered, pitch, torch.cat)
        full = _oracle.OracleFilter(p, frames).run()
        ok = all((g.numpy() == f.astype(np.int32)).all() for g, f in zip(planes, full["out"]))
        ok = ok and (d.numpy() == full["diff"]).all()
        q.put(bool(ok))
    else:
        assert gathered is None
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
