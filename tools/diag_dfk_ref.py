"""Config 3 (DFKernel 25 x 2048^2, Nw=3, max_shift=5) on the blur-table path, assign_coordinates 'sam' and 'ref'."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDFKernel, synth
d = synth.speckle_stack(25, 2048, 2048, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
m = UMPAModelDFKernel(list(d["sam"]), list(d["ref"]), window_size=3, max_shift=5)
N0, N1 = m.sh
abc = synth.blur_abc(N0, N1, as_numpy=False).cuda()
for assign in ("sam", "ref"):
    m.assign_coordinates = assign
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = m.match_device(abc=abc); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ok = float((r["err"] == 1).float().mean())
    print("DFKernel assign=%s: %.1f ms (%s) ok %.4f" % (assign, dt * 1e3, m.last_match_info["path"], ok))
