"""CPU, world_size 2, gloo: the N>1 path's only exchange (gather of the row bands)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umpa_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, n_cols, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bands = sharding.row_bands(n_rows, world)
    r0, r1 = bands[rank]
    full = torch.arange(n_rows * n_cols, dtype=torch.float64).reshape(n_rows, n_cols)
    local = {"dx": full[r0:r1].clone(), "err": (full[r0:r1] % 2).to(torch.int32)}
    if rank == 1:
        local["df"] = -full[r0:r1].clone()
    else:
        local["df"] = -full[r0:r1].clone()
    got = sharding.gather_bands(local, bands, rank, keys=("dx", "df", "err", "f"), dst=0)
    if rank == 0:
        ok = (torch.equal(got["dx"], full) and torch.equal(got["df"], -full)
              and torch.equal(got["err"], (full % 2).to(torch.int32)) and "f" not in got)
        q.put(bool(ok))
    else:
        assert got is None
    dist.destroy_process_group()


def test_gather_bands_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 11, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
