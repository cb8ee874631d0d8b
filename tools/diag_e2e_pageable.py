"""Config 2 end to end with PAGEABLE host frames (ordinary numpy arrays), float64 and float32."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, synth
Na, H, W = 25, 2048, 2048
d = synth.speckle_stack(Na, H, W, seed=2, max_shift=5, dark_field=True, device="cuda", as_numpy=False)
s64, r64 = d["sam"].cpu().numpy(), d["ref"].cpu().numpy()
del d
s32, r32 = s64.astype(np.float32), r64.astype(np.float32)
for name, (s, r) in (("float64 pageable", (s64, r64)), ("float32 pageable", (s32, r32))):
    for rep in range(5):
        if rep == 4 and len(sys.argv) > 1:
            os.environ["UMPA_STREAM_TRACE"] = "1"
        t0 = time.perf_counter()
        m = UMPAModelDF(list(s), list(r), window_size=2, max_shift=5)
        res = m.match(quiet=True, debug=False)
        dt = time.perf_counter() - t0
        info = m.last_stream_info
        del m
    os.environ.pop("UMPA_STREAM_TRACE", None)
    print("%s: %.2f ms  %s" % (name, dt * 1e3, info))
