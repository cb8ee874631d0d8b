"""Config 2 end to end with float32 host frames: where the time goes (constructor / match / stream timeline)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, synth
Na, H, W = 25, 2048, 2048
d = synth.speckle_stack(Na, H, W, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
hs = torch.empty((Na, H, W), dtype=torch.float32, pin_memory=True); hr = torch.empty_like(hs, pin_memory=True)
hs.copy_(d["sam"]); hr.copy_(d["ref"]); torch.cuda.synchronize()
del d
sam_np, ref_np = hs.numpy(), hr.numpy()
for rep in range(5):
    if rep == 4:
        os.environ["UMPA_STREAM_TRACE"] = "1"
    t0 = time.perf_counter()
    m = UMPAModelDF(list(sam_np), list(ref_np), window_size=2, max_shift=5)
    t1 = time.perf_counter()
    r = m.match(quiet=True, debug=False)
    t2 = time.perf_counter()
    print("ctor %.2f ms  match %.2f ms  total %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3))
    del m
