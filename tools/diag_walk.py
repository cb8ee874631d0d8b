"""Stage times of a config-2 step (and a noisy variant) + a checksum of the result maps, for comparing tuning
builds of the library (UMPA_LIB=...): the checksums of two builds must agree."""
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from umpa_b200 import UMPAModelDF, UMPAModelNoDF, _capi, synth
for name, cls, kw in (("cfg2", UMPAModelDF, dict(noise=0.)), ("cfg2 noisy", UMPAModelDF, dict(noise=.2, amplitude=2.5))):
    Na, Nw, ms = (25, 2, 5) if cls is UMPAModelDF else (4, 6, 4)
    d = synth.speckle_stack(Na, 2048, 2048, seed=2, max_shift=ms, dark_field=cls is UMPAModelDF, device="cuda", as_numpy=False, **kw)
    m = cls(list(d["sam"]), list(d["ref"]), window_size=Nw, max_shift=ms)
    for _ in range(3):
        out = m.match_device()
    torch.cuda.synchronize()
    _capi.check(_capi.lib().umpa_set_profiling(m._h, 1))
    acc = np.zeros(4)
    for _ in range(10):
        m.match_device()
        buf = (C.c_float * 4)()
        _capi.lib().umpa_last_stage_ms(m._h, buf, 4)
        acc += np.array(buf[:4])
    h = hashlib.sha1()
    for k in sorted(out):
        if torch.is_tensor(out[k]):
            h.update(out[k].cpu().numpy().tobytes())
    print("%-10s moments %.3f cross %.3f mean %.3f walk %.3f  total %.3f ms  ok %.4f calls %.2f  sha %s" % (
        (name,) + tuple(acc / 10) + (acc.sum() / 10, float((out["err"] == 1).float().mean()),
                                     float(out["debug_Ncalls"].float().mean()), h.hexdigest()[:12])))
    del m, d, out
