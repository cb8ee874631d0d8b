"""NOT part of the test suite: non-finite input pixels.  With a NaN cost next to finite ones the reference's walk
(Optim.cpp:233-479) can step back and forth between two evaluated shifts for ever -- ncalls only counts
evaluations -- and so could ours (observed: this script hung the GPU box before walk.cuh got its idle-visit
guard).  Run it under `timeout` to check the guard and the locality of the damage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import umpa_b200
from umpa_b200 import synth

KEYS = ("f", "T", "dx", "dy", "err", "debug_Ncalls")
for path in ("table", "lazy"):
    d = synth.speckle_stack(6, 200, 220, seed=31, max_shift=4, dark_field=True)
    sam, ref = np.array(d["sam"]), np.array(d["ref"])
    clean = umpa_b200.UMPAModelDF(list(sam), list(ref), max_shift=4)
    clean.cuda_path = path
    want = clean.match(quiet=True)
    sam[2, 66, 100] = np.nan                      # rows 0, 6, 12, ... are the sampled ones (H // 32 = 6)
    ref[1, 120, 50] = np.inf
    m = umpa_b200.UMPAModelDF(list(sam), list(ref), max_shift=4)
    m.cuda_path = path
    got = m.match(quiet=True)
    pad = m.padding
    yy, xx = np.mgrid[pad:200 - pad, pad:220 - pad]
    far = np.ones(yy.shape, bool)
    for (y, x) in ((66, 100), (120, 50)):
        far &= (np.abs(yy - y) > pad) | (np.abs(xx - x) > pad)
    same = (got["err"] == want["err"]) & (got["debug_Ncalls"] == want["debug_Ncalls"])
    ok = far & same & (want["err"] == 1)
    print(path, "far pixels with the clean err/Ncalls: %.5f" % same[far].mean(),
          {k: float(np.abs(got[k][ok] - want[k][ok]).max()) for k in ("dx", "dy", "T", "df")},
          "near pixels ok:", float((got["err"][~far] == 1).mean()))
