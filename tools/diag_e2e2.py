import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, synth
Na, H, W = 25, 2048, 2048
d = synth.speckle_stack(Na, H, W, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
hs = torch.empty((Na, H, W), dtype=torch.float64, pin_memory=True); hr = torch.empty_like(hs, pin_memory=True)
hs.copy_(d["sam"]); hr.copy_(d["ref"]); torch.cuda.synchronize()
del d
sam_np, ref_np = hs.numpy(), hr.numpy()
for rep in range(4):
    if rep == 3:
        os.environ["UMPA_STREAM_TRACE"] = "1"
    t0 = time.perf_counter()
    m = UMPAModelDF(list(sam_np), list(ref_np), window_size=2, max_shift=5)
    r = m.match(quiet=True, debug=False)
    print("total %.2f ms" % ((time.perf_counter() - t0) * 1e3))
    del m
