import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, UMPAModelNoDF, synth
d = synth.speckle_stack(25, 2048, 2048, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
for cls in (UMPAModelDF, UMPAModelNoDF):
    m = cls(list(d["sam"]), list(d["ref"]), window_size=2, max_shift=5)
    m.cuda_path = "lazy"
    roi = ((0, 512, 1), (0, 2034, 1))
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = m.match_device(ROI=roi); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = r["f"].numel()
    print(cls.__name__, "lazy path: %.1f ms for %d px -> %.3g px/s (full frame would take %.0f ms)" % (dt * 1e3, n, n / dt, dt * 1e3 * 2034 / 512))
    mask = torch.rand((25, 2048, 2048), dtype=torch.float64, device="cuda").clamp_(.5, 1.)
    mm = cls(list(d["sam"]), list(d["ref"]), mask_list=list(mask), window_size=2, max_shift=5)
    mm.cuda_path = "lazy"
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = mm.match_device(ROI=roi); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(cls.__name__, "lazy path, masked: %.1f ms -> %.3g px/s" % (dt * 1e3, n / dt))
    del m, mm
