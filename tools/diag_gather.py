"""Two (or more) GPUs: the final gather of the maps, one message per (peer, map) against two packed messages per peer.
Usage: torchrun --nproc-per-node 2 tools/diag_gather.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from umpa_b200.sharding import gather_bands, row_bands as split_rows
for name, N0, N1 in (("cfg2", 2034, 2034), ("cfg4", 4074, 4074)):
    bands = split_rows(N0, world)
    r0, r1 = bands[rank]
    keys = ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls")
    loc = {k: (torch.full((r1 - r0, N1), rank, dtype=torch.int32, device="cuda") if k in ("err", "debug_Ncalls")
               else torch.full((r1 - r0, N1), float(rank), dtype=torch.float64, device="cuda")) for k in keys}
    for mode in ("maps", "packed", "allgather"):
        os.environ["UMPA_GATHER"] = mode
        for _ in range(3):
            out = gather_bands(loc, bands, rank, keys=keys, dst=0)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = gather_bands(loc, bands, rank, keys=keys, dst=0)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            ok = all(bool((out[k][b0:b1] == r).all()) for k in keys for r, (b0, b1) in enumerate(bands))
            print("%s world %d %-6s %.3f ms per gather, correct %s" % (name, world, mode, e0.elapsed_time(e1) / 20, ok), flush=True)
        dist.barrier()
dist.destroy_process_group()
