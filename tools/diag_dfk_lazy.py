import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDFKernel, synth
d = synth.speckle_stack(25, 2048, 2048, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
m = UMPAModelDFKernel(list(d["sam"]), list(d["ref"]), window_size=3, max_shift=5)
N0, N1 = m.sh
abc = synth.blur_abc(N0, N1, as_numpy=False).cuda()
m.cuda_path = "lazy"
roi = ((1000, 1064, 1), (0, 2016, 1))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = m.match_device(ROI=roi, abc=abc[1000:1064]); torch.cuda.synchronize(); dt = time.perf_counter() - t0
n = r["f"].numel()
print("DFKernel lazy: %.1f ms for %d px -> %.3g px/s (full frame would take %.1f s)" % (dt * 1e3, n, n / dt, dt * 2016 / 64))
