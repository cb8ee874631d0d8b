import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import umpa_b200
from umpa_b200 import synth
d = synth.speckle_stack(6, 200, 220, seed=31, max_shift=4, dark_field=True)
for which in ("nan", "inf", "both"):
    sam, ref = np.array(d["sam"]), np.array(d["ref"])
    clean = umpa_b200.UMPAModelDF(list(sam), list(ref), max_shift=4); clean.cuda_path = "table"
    want = clean.match(quiet=True, debug=False)
    bad = []
    if which in ("nan", "both"): sam[2, 66, 100] = np.nan; bad.append((66, 100))
    if which in ("inf", "both"): ref[1, 120, 50] = np.inf; bad.append((120, 50))
    m = umpa_b200.UMPAModelDF(list(sam), list(ref), max_shift=4); m.cuda_path = "table"
    got = m.match(quiet=True, debug=False)
    pad = m.padding
    diff = (np.abs(got["dx"] - want["dx"]) > 1e-4) | (got["err"] != want["err"]) | ~np.isfinite(got["dx"])
    ys, xs = np.nonzero(diff)
    print(which, "pixels that differ:", len(ys), "raw rows %d..%d cols %d..%d" % (ys.min() + pad, ys.max() + pad, xs.min() + pad, xs.max() + pad) if len(ys) else "")
    for (y, x) in bad:
        dist = np.maximum(np.abs(ys + pad - y), np.abs(xs + pad - x))
        print("   bad pixel", (y, x), "chebyshev distance of differing pixels: max", dist.max() if len(ys) else None, np.bincount(dist)[:14] if len(ys) else None)
