"""GPU, world size 2, NCCL: the row-sharded match of SURVEY.md 8(e) end to end -- every rank matches its band
(+ halo) on its own GPU, the maps are gathered on rank 0 over NCCL (the path's only inter-GPU traffic) and must
equal the unsharded match bit for bit (pixels are independent; the centring constants of a band differ from the
full frame's, so "equal" is to the FP32 tables' rounding: the integer walk must agree, the values to 1e-5).
Skipped where fewer than two GPUs are visible (the CPU twin of the gather is tests/test_sharding_gloo.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from umpa_b200 import UMPAModelDF, synth
    from umpa_b200.sharding import ShardedMatcher
    d = synth.speckle_stack(8, 300, 333, seed=7, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
    sm = ShardedMatcher(UMPAModelDF, list(d["sam"]), list(d["ref"]), rank, world, window_size=2, max_shift=5)
    loc = sm.match_device()
    keys = ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls")
    full = sm.gather(loc, keys=keys, dst=0)
    if rank == 0:
        ref = UMPAModelDF(list(d["sam"]), list(d["ref"]), window_size=2, max_shift=5).match_device()
        torch.cuda.synchronize()
        res = {}
        for k in keys:
            a, b = full[k].cpu().numpy(), ref[k].cpu().numpy()
            res[k] = (a.shape == b.shape, float(np.abs(a.astype(np.float64) - b).max()), float((a != b).mean()))
        q.put(res)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_match_and_nccl_gather_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for k, (same_shape, maxdiff, frac) in res.items():
        assert same_shape, k
    assert res["err"][1] == 0 and res["debug_Ncalls"][2] < 2e-3, res          # the integer walk agrees (FP32 ties aside)
    for k in ("dx", "dy"):
        assert res[k][1] < 1e-4 or res["debug_Ncalls"][2] > 0, res
    for k in ("T", "df"):
        assert res[k][1] < 1e-4 or res["debug_Ncalls"][2] > 0, res
