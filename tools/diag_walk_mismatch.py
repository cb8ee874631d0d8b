"""Config 2: where does the integer walk on the FP32 tables differ from the FP64 lazy evaluation (the reference's own
arithmetic), and why?  For every such pixel both 5x5 cost caches are printed with the smallest gap between any two
entries of the FP64 cache -- a walk that turned the other way at a near-tie shows up as a gap at FP32 noise level.
Usage: python tools/diag_walk_mismatch.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, synth
Na, N, Nw, ms = 25, 2048, 2, 5
d = synth.speckle_stack(Na, N, N, seed=2, max_shift=ms, dark_field=True, device="cuda", as_numpy=False)
m = UMPAModelDF(list(d["sam"]), list(d["ref"]), window_size=Nw, max_shift=ms)
tab = m.match_device()
m.cuda_path = "lazy"
laz = m.match_device()
torch.cuda.synchronize()
diff = (tab["debug_Ncalls"] != laz["debug_Ncalls"]) | (tab["err"] != laz["err"])
far = ((tab["dx"] - laz["dx"]).abs() > 1e-3) | ((tab["dy"] - laz["dy"]).abs() > 1e-3)
idx = torch.nonzero(diff | far).cpu().numpy()
scale = float(laz["f"].median())
print("pixels whose walk differs: %d of %d; final position differs on %d; cost scale %.4g" % (int(diff.sum()), diff.numel(), int(far.sum()), scale))
np.set_printoptions(precision=9, linewidth=200)
for i, j in idx[:12]:
    roi = ((int(i), int(i) + 1, 1), (int(j), int(j) + 1, 1))
    m.cuda_path = "table"
    a = m.match(ROI=roi, quiet=True, debug=True)
    m.cuda_path = "lazy"
    b = m.match(ROI=roi, quiet=True, debug=True)
    da, db = a["debug_d"][0, 0], b["debug_d"][0, 0]
    known = db[db > -.5]
    gaps = np.sort(np.abs(known[:, None] - known[None, :])[np.triu_indices(len(known), 1)])
    print("pixel (%d, %d): table Ncalls %d dx %.6f dy %.6f f %.9g | lazy Ncalls %d dx %.6f dy %.6f f %.9g" % (
        i, j, a["debug_Ncalls"][0, 0], a["dx"][0, 0], a["dy"][0, 0], a["f"][0, 0],
        b["debug_Ncalls"][0, 0], b["dx"][0, 0], b["dy"][0, 0], b["f"][0, 0]))
    print("   smallest gaps between two entries of the FP64 cache / cost scale:", gaps[:3] / scale)
    print("   FP64 cache:\n", db.reshape(5, 5), "\n   table cache:\n", da.reshape(5, 5))
