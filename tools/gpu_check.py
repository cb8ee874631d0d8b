#!/usr/bin/env python
"""Diagnostic run on a GPU box: every golden case through both CUDA paths, statistics printed
(not asserted).  Usage: python tools/gpu_check.py > gpurun_out/gpu_check.log"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import golden_names, load_case      # noqa: E402
from gpu_common import TABLE_CASES, run_case     # noqa: E402


def stats(got, exp):
    out = []
    eg, ee = got["err"], exp["err"]
    out.append("err_mismatch=%d/%d" % ((eg != ee).sum(), ee.size))
    ok = (ee == 1) & (eg == 1)
    out.append("ncalls_mismatch=%d" % (got["debug_Ncalls"][ok] != exp["debug_Ncalls"][ok]).sum())
    for k in ("dx", "dy", "T", "df", "f"):
        if k in exp and k in got and ok.any():
            g, e = got[k][ok], exp[k][ok]
            scale = np.maximum(1., np.abs(e)) if k in ("dx", "dy") else np.maximum(np.abs(e), 1e-300)
            rel = np.abs(g - e) / scale
            out.append("%s: max=%.2e p99.9=%.2e n>1e-4=%d" % (k, np.nanmax(rel), np.nanpercentile(rel, 99.9), (rel > 1e-4).sum()))
    return "  ".join(out)


def main():
    for name in golden_names():
        case = load_case(name)
        for path in ("lazy", "table"):
            if path == "table" and name not in TABLE_CASES:
                continue
            try:
                t = time.time()
                m, got = run_case(case, path)
                dt = time.time() - t
                print("%-16s %-5s [%s] %.3fs %s" % (name, path, m.last_match_info, dt, stats(got, case["expected"])), flush=True)
            except Exception:
                print("%-16s %-5s FAILED" % (name, path))
                traceback.print_exc(file=sys.stdout)
                sys.stdout.flush()


if __name__ == "__main__":
    main()
