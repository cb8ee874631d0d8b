"""Shared helpers for the parity tests: golden-fixture loading and comparison rules.

Comparison rules (SURVEY.md 8d "Parity criteria"):
  * ``err`` maps must be equal (a documented handful of FP32 near-ties may be allowed by the caller);
  * on pixels where the reference succeeded (err == 1): integer walk (Ncalls) equal,
    dx/dy within tol * max(1, |ref|), T/df/f within tol relative;
  * on err == 0 pixels only ``err`` is compared -- the reference's other outputs
    are partly uninitialised there (Model.cpp:566, 927).
"""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not p.endswith("hooks.npz"))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    c = {k: z[k] for k in z.files}
    Na = int(c["Na"])
    c["sam"] = [c.pop(f"sam{k}") for k in range(Na)]
    c["ref"] = [c.pop(f"ref{k}") for k in range(Na)]
    c["mask"] = [c.pop(f"mask{k}") for k in range(Na)] if "mask0" in c else None
    c["pos"] = [tuple(int(v) for v in p) for p in c["pos"]] if "pos" in c else None
    c["kind"] = str(c["kind"])
    for k in ("Nw", "max_shift", "padding"):
        c[k] = int(c[k])
    c["assign"] = str(c["assign"]) if "assign" in c else None
    c["subpx"] = int(c["subpx"]) if "subpx" in c else None
    c["step"] = int(c["step"]) if "step" in c else None
    c["ROI"] = tuple(tuple(int(v) for v in r) for r in c["ROI"]) if "ROI" in c else None
    c["dxdy"] = tuple(float(v) for v in c["dxdy"]) if "dxdy" in c else None
    c["abc"] = c.get("abc")
    c["expected"] = {k[4:]: c[k] for k in list(c) if k.startswith("out_")}
    return c


def roi_of(case):
    """The ((start,stop,step),(start,stop,step)) the reference ended up using."""
    r = case["ROI_after"]
    return tuple(tuple(int(v) for v in x) for x in r)


def compare(got, exp, tol=1e-9, walk_exact=True, max_err_mismatch=0, max_outliers=0, outlier_tol=None,
            label=""):
    """Returns a dict of statistics; raises AssertionError with a readable message."""
    stats = {}
    err_g, err_e = np.asarray(got["err"]), np.asarray(exp["err"])
    assert err_g.shape == err_e.shape, f"{label}: shape {err_g.shape} vs {err_e.shape}"
    mism = int((err_g != err_e).sum())
    stats["err_mismatch"] = mism
    assert mism <= max_err_mismatch, f"{label}: err map differs in {mism} pixels"
    ok = (err_e == 1) & (err_g == 1)
    if walk_exact and "debug_Ncalls" in got and "debug_Ncalls" in exp:
        nm = int((np.asarray(got["debug_Ncalls"])[ok] != np.asarray(exp["debug_Ncalls"])[ok]).sum())
        stats["ncalls_mismatch"] = nm
        assert nm <= max_outliers, f"{label}: Ncalls differs in {nm} ok-pixels"
    for k in ("dx", "dy", "T", "df", "f"):
        if k not in exp or k not in got:
            continue
        g, e = np.asarray(got[k], dtype=np.float64)[ok], np.asarray(exp[k], dtype=np.float64)[ok]
        if g.size == 0:
            continue
        scale = np.maximum(1., np.abs(e)) if k in ("dx", "dy") else np.maximum(np.abs(e), 1e-300)
        rel = np.abs(g - e) / scale
        rel = np.where(np.isfinite(rel), rel, np.inf)
        bad = rel > tol
        stats[k + "_max"] = float(rel.max())
        stats[k + "_outliers"] = int(bad.sum())
        assert bad.sum() <= max_outliers, (
            f"{label}: {k} exceeds {tol:g} in {int(bad.sum())} of {g.size} ok-pixels (max {rel.max():.3g})")
        if outlier_tol is not None and bad.any():
            assert rel.max() <= outlier_tol, f"{label}: {k} outlier {rel.max():.3g} > {outlier_tol:g}"
    return stats
