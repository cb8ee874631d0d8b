import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelNoDF, synth
d = synth.speckle_stack(10, 256, 256, seed=1, max_shift=4, dark_field=False)
hs = torch.from_numpy(d["sam"]).pin_memory().numpy(); hr = torch.from_numpy(d["ref"]).pin_memory().numpy()
sam, ref = list(hs), list(hr)
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = UMPAModelNoDF(sam, ref, window_size=2, max_shift=4); t1 = time.perf_counter()
    r = m.match(quiet=True, debug=False); t2 = time.perf_counter()
    del m; t3 = time.perf_counter()
    print("ctor %.3f ms  match %.3f ms  del %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
