"""Device-resident match of config 2 (or cfg5 / cfg4 shapes) with a random dead-pixel map shared by all frames:
the corrected table walk against the FP64 lazy evaluation (UMPA_MASK_TABLES=0).
Usage: python tools/prof_masked.py [cfg2|cfg4|cfg5] [dead fraction, default 0.03] [--no-lazy]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from umpa_b200 import UMPAModelDF, UMPAModelNoDF, synth
CFG = {"cfg2": (UMPAModelDF, 25, 2048, 2, 5), "cfg4": (UMPAModelDF, 40, 4096, 3, 8), "cfg5": (UMPAModelNoDF, 4, 2048, 6, 4)}
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "cfg2"
frac = float(args[1]) if len(args) > 1 else .03
cls, Na, N, Nw, ms = CFG[name]
d = synth.speckle_stack(Na, N, N, seed=2, max_shift=ms, dark_field=cls is UMPAModelDF, device="cuda", as_numpy=False)
g = torch.Generator(device="cpu").manual_seed(1)
M = (torch.rand((N, N), generator=g, dtype=torch.float64) >= frac).to(torch.float64).cuda()
masks = [M] * Na


def run(tag):
    m = cls(list(d["sam"]), list(d["ref"]), mask_list=masks, window_size=Nw, max_shift=ms)
    m.match_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        out = m.match_device()
    e1.record()
    torch.cuda.synchronize()
    ms_ = e0.elapsed_time(e1) / n
    px = out["err"].numel()
    print("%s %s dead %.3f: %.2f ms per match (%.3e px/s), path %s, ok fraction %.4f" % (
        name, tag, frac, ms_, px / ms_ * 1e3, m.last_match_info, float((out["err"] == 1).float().mean())), flush=True)
    return out


a = run("corrected table walk")
if "--no-lazy" not in sys.argv:
    os.environ["UMPA_MASK_TABLES"] = "0"
    b = run("lazy evaluation     ")
    ok = (a["err"] == 1) & (b["err"] == 1)
    print("err maps equal:", bool((a["err"] == b["err"]).all()), " Ncalls differ on", int((a["debug_Ncalls"] != b["debug_Ncalls"])[ok].sum()),
          "of", int(ok.sum()), " max |dx| deviation", float((a["dx"] - b["dx"]).abs()[ok & (a["debug_Ncalls"] == b["debug_Ncalls"])].max()))
