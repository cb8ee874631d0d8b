"""Where does the first host-to-host match() of a process spend its time?  (fresh process; config 2 shapes)
Usage: python tools/diag_first_call.py [n_calls]   -- env knobs of the library apply (UMPA_HOST_THREADS=0, ...)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
t0 = time.perf_counter()
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
print("cuda context + first torch kernel %.1f ms" % (1e3 * (time.perf_counter() - t0)))
from umpa_b200 import UMPAModelDF, synth
Na, N, Nw, ms = 25, 2048, 2, 5
d = synth.speckle_stack(Na, N, N, seed=2, max_shift=ms, dark_field=True, device="cuda", as_numpy=False)
t0 = time.perf_counter()
hs = torch.empty((Na, N, N), dtype=torch.float64, pin_memory=True); hr = torch.empty_like(hs, pin_memory=True)
print("pinning the caller's 2 x 839 MB input stacks (not part of the call) %.1f ms" % (1e3 * (time.perf_counter() - t0)))
hs.copy_(d["sam"]); hr.copy_(d["ref"]); torch.cuda.synchronize()
del d
sam, ref = list(hs.numpy()), list(hr.numpy())
t0 = time.perf_counter()
x = torch.empty((5 * 2034 * 2034,), dtype=torch.float64, pin_memory=True)
print("pinning 166 MB of result maps with torch (first time) %.1f ms" % (1e3 * (time.perf_counter() - t0)))
del x
for n in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    t0 = time.perf_counter()
    m = UMPAModelDF(sam, ref, window_size=Nw, max_shift=ms)
    t1 = time.perf_counter()
    r = m.match(quiet=True, debug=False)
    t2 = time.perf_counter()
    print("call %d: constructor %.1f ms, match %.1f ms  %s" % (n, 1e3 * (t1 - t0), 1e3 * (t2 - t1), m.last_stream_info))
    del m, r
