import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, synth
d = synth.speckle_stack(25, 2048, 2048, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
mask = torch.ones((25, 2048, 2048), dtype=torch.float64, device="cuda")
mask[7, 900:920, 1100:1130] = 0.
mask[3, 100, 200] = 0.
m = UMPAModelDF(list(d["sam"]), list(d["ref"]), mask_list=list(mask), window_size=2, max_shift=5)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s0, s1 = m._convert_ROI_slice(None, None)
    c = m._coverage_device(s0, s1); torch.cuda.synchronize(); t1 = time.perf_counter()
    ch = c.cpu().numpy(); t2 = time.perf_counter()
    r = m.match(quiet=True, debug=False); t3 = time.perf_counter()
    dev = m.match_device(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print("coverage %.1f ms, to host %.1f ms, match() %.1f ms, match_device() %.1f ms, %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, m.last_match_info))
