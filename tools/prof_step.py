"""Two device-resident matches of one configuration (default: config 2) -- the command line profiled under ncu.
Usage: python tools/prof_step.py [cfg2|cfg1|cfg4|cfg5] [n_matches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from umpa_b200 import UMPAModelDF, UMPAModelNoDF, synth
CFG = {"cfg1": (UMPAModelNoDF, 10, 256, 2, 4), "cfg2": (UMPAModelDF, 25, 2048, 2, 5),
       "cfg4": (UMPAModelDF, 40, 4096, 3, 8), "cfg5": (UMPAModelNoDF, 4, 2048, 6, 4)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cls, Na, N, Nw, ms = CFG[name]
d = synth.speckle_stack(Na, N, N, seed=2, max_shift=ms, dark_field=cls is UMPAModelDF, device="cuda", as_numpy=False)
m = cls(list(d["sam"]), list(d["ref"]), window_size=Nw, max_shift=ms)
for _ in range(reps):
    out = m.match_device()
torch.cuda.synchronize()
print(name, "ok fraction", float((out["err"] == 1).float().mean()), m.last_match_info)
