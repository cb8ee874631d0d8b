"""Diagnostic: where does the end-to-end (host -> host) time of config 2 go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, synth
Na, H, W = 25, 2048, 2048
d = synth.speckle_stack(Na, H, W, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
hs = torch.empty((Na, H, W), dtype=torch.float64, pin_memory=True); hr = torch.empty_like(hs, pin_memory=True)
hs.copy_(d["sam"]); hr.copy_(d["ref"]); torch.cuda.synchronize()
del d
dev = torch.empty((2, Na, H, W), dtype=torch.float64, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    dev[0].copy_(hs, non_blocking=True); dev[1].copy_(hr, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("raw H2D 1.68 GB: %.2f ms  %.1f GB/s" % (dt * 1e3, 2 * hs.numel() * 8 / dt / 1e9))
ho = torch.empty((25, 2034, 2034), dtype=torch.float64, pin_memory=True)[:3]
do = torch.empty_like(ho, device="cuda")
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    ho.copy_(do, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("raw D2H %.0f MB: %.2f ms  %.1f GB/s" % (ho.numel() * 8 / 1e6, dt * 1e3, ho.numel() * 8 / dt / 1e9))
del dev
sam_np, ref_np = hs.numpy(), hr.numpy()
for nb in (None, "1", "2", "4", "6", "8", "16"):
    if nb is None:
        os.environ.pop("UMPA_BANDS", None)
    else:
        os.environ["UMPA_BANDS"] = nb
    tc = tm = 0.
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m = UMPAModelDF(list(sam_np), list(ref_np), window_size=2, max_shift=5)
        t1 = time.perf_counter()
        r = m.match(quiet=True, debug=False)
        t2 = time.perf_counter()
        del m
        if rep >= 2:
            tc += t1 - t0; tm += t2 - t1
    print("bands %s: ctor %.2f ms  match %.2f ms" % (nb, tc / 4 * 1e3, tm / 4 * 1e3))
