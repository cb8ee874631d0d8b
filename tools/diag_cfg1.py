import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from umpa_b200 import UMPAModelNoDF, synth
from oracle import port
np.set_printoptions(precision=6, linewidth=200)
d = synth.speckle_stack(10, 256, 256, seed=1, max_shift=4, dark_field=False)
exp = port.OracleModel("NoDF", d["sam"], d["ref"], window_size=2, max_shift=4).match()
m = UMPAModelNoDF(d["sam"], d["ref"], window_size=2, max_shift=4)
got = m.match(quiet=True)
ok = exp["err"] == 1
rel = np.abs(got["T"] - exp["T"]) / np.abs(exp["T"])
idx = np.argwhere(ok & (rel > 1e-4))
print("n deviating", len(idx))
for i, j in idx[:4]:
    print("pixel", i, j, "T", got["T"][i, j], exp["T"][i, j], "dx", got["dx"][i, j], exp["dx"][i, j], "dy", got["dy"][i, j], exp["dy"][i, j], "Ncalls", got["debug_Ncalls"][i, j], exp["debug_Ncalls"][i, j])
    print(" ref d:\n", exp["debug_d"][i, j].reshape(5, 5))
    print(" got d - ref d:\n", got["debug_d"][i, j].reshape(5, 5) - exp["debug_d"][i, j].reshape(5, 5))
