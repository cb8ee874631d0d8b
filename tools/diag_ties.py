"""Diagnostic: where does the table path's integer walk differ from the lazy path's?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from umpa_b200 import UMPAModelDF, synth
d = synth.speckle_stack(12, 300, 333, seed=4, max_shift=5, dark_field=True)
m = UMPAModelDF(d["sam"], d["ref"], window_size=2, max_shift=5)
m.cuda_path = "lazy"
exp = m.match(quiet=True)
m.cuda_path = "table"
got = m.match(quiet=True)
ok = exp["err"] == 1
print("err equal", np.array_equal(exp["err"], got["err"]), "ok", ok.mean())
bad = np.argwhere(ok & (exp["debug_Ncalls"] != got["debug_Ncalls"]))
print("ncalls mismatches", len(bad))
for i, j in bad[:10]:
    print((i, j), exp["debug_Ncalls"][i, j], got["debug_Ncalls"][i, j], "dx", exp["dx"][i, j], got["dx"][i, j], "dy", exp["dy"][i, j], got["dy"][i, j])
    print(" exp d", np.array2string(exp["debug_d"][i, j].reshape(5, 5), precision=9))
    print(" got d", np.array2string(got["debug_d"][i, j].reshape(5, 5), precision=9))
for k in ("dx", "dy", "T", "df", "f"):
    print(k, np.abs(exp[k] - got[k])[ok].max(), np.percentile(np.abs(exp[k] - got[k])[ok], 99.9))
