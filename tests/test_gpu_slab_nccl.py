"""GPU, 2 ranks, NCCL: the slab-parallel path of BASELINE.json config 4 as bench.py runs it -- one process per
GPU, every rank filters its block-row range of the same window through the library's row-range mode,
sharding.SlabWindow gathers the slabs from the library's device output planes to rank 0 over NCCL and all-reduces
FRAME_DIFF.  The gathered frame must equal the frame one GPU computes alone, bit for bit, and the oracle's rows;
so must the frame assembled without a gather, by every rank storing its rows into rank 0's planes (CUDA IPC peer mapping).
Needs two GPUs (run with `gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")]

CASES = {"1080p10_n5": (1920, 1080, 10, 5), "4k10_n3": (3840, 2160, 10, 3), "cif8_n5": (352, 288, 8, 5)}


def _worker(rank, world, port, case, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import importlib
    import torch.distributed as dist
    import conftest
    import _clips
    import _oracle
    import _params
    pkg = conftest.load_package()
    sh = importlib.import_module("aom_av1_psy_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    W, H, bd, N = CASES[case]
    frames = _clips.moving_texture(W, H, N, bd)
    p = _params.tf_params(W, H, N, bit_depth=bd, allow_hp=1)
    ctx = pkg.TemporalFilterGpu(device=rank, max_cached_frames=24)
    ids = []
    for i, (y, u, v) in enumerate(frames):
        b = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"], frame_id=1 + i).set_planes(y, u, v, extend=False)
        ctx.cache_frame(b)
        ids.append(b.frame_id)
    mb_rows = (H + 31) // 32
    sw = sh.SlabWindow(mb_rows, world, rank)
    ok = True
    for rep in range(2):  # twice: the receive buffers are reused
        _, diff = ctx.filter_resident(sw.params(p), ids)
        d = torch.from_numpy(diff.copy()).to(f"cuda:{rank}")
        g, dsum = sw.gather(sw.device_slabs(ctx, torch, f"cuda:{rank}"), d, dist, torch)
        torch.cuda.synchronize()
        if rank == 0:
            pitches = [ctx.output_device_plane(pl)[1] for pl in range(3)]
            planes = [t.cpu().numpy() for t in sw.assemble(g, pitches, torch.cat)]
            _, full_diff = ctx.filter_resident(dict(p, out_row_begin=0, out_row_end=0), ids)  # one GPU, whole frame
            out = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
            ctx.download_output(out)
            ok = ok and bool((dsum.cpu().numpy() == full_diff).all())
            for pl in range(3):
                want = out.full_blocks(pl)
                got = planes[pl][:, :want.shape[1] * want.itemsize].copy().view(want.dtype)
                ok = ok and got.shape == want.shape and bool((got == want).all())
            if rep == 0:  # the oracle on the first row of rank 1's slab: the seam between the slabs
                b1 = sh.slab_rows(mb_rows, world, 1)[0]
                o = _oracle.OracleFilter(p, frames)
                ref = o.run(record=False, rows=(b1, b1 + 1))
                o.close()
                for pl in range(3):
                    bh = 32 >> (1 if pl else 0)
                    want = ref["out"][pl][b1 * bh:(b1 + 1) * bh]
                    got = planes[pl][b1 * bh:(b1 + 1) * bh, :want.shape[1] * out.alloc[pl].itemsize].copy().view(out.dtype)
                    ok = ok and bool((got.astype(np.uint16) == want).all())
        else:
            assert g is None
    # the same slabs without a gather: every rank stores its rows into rank 0's planes over NVLink (CUDA IPC)
    sw.connect_peer_output(ctx, dist)
    if rank == 0:
        for t in sw.device_planes(ctx, torch, "cuda:0"):
            t.zero_()
        torch.cuda.synchronize()
    dist.barrier()
    _, diff = ctx.filter_resident(sw.params(p), ids)
    dsum = sw.finish_peer(torch.from_numpy(diff.copy()).to(f"cuda:{rank}"), dist)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        got = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
        ctx.download_output(got)
        sw.disconnect_peer_output(ctx)
        _, full_diff = ctx.filter_resident(dict(p, out_row_begin=0, out_row_end=0), ids)
        ref = pkg.Yv12Buffer(W, H, 1, 1, bd > 8, p["border"])
        ctx.download_output(ref)
        ok = ok and bool((dsum.cpu().numpy() == full_diff).all())
        for pl in range(3):
            ok = ok and bool((got.full_blocks(pl) == ref.full_blocks(pl)).all())
    else:
        sw.disconnect_peer_output(ctx)
    dist.barrier()
    if rank == 0:
        q.put(ok)
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", sorted(CASES))
def test_slab_gather_over_nccl_equals_one_gpu(case):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [mpc.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=600)
        assert pr.exitcode == 0
    assert q.get(timeout=10) is True
