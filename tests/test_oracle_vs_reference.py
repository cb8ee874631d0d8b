"""CPU: the C restatement (oracle/) against the UNMODIFIED reference compiled into oracle/_ref, live, on
seeded inputs that are not among the golden fixtures.  Skipped where the compiled reference is not
available (it is built in the build container from /root/reference and travels with the repo)."""
import numpy as np
import pytest

from helpers import compare
from oracle import port, ref as oref
from umpa_b200 import synth

R = oref.load(build_if_missing=True)
pytestmark = pytest.mark.skipif(R is None, reason="compiled reference (oracle/_ref) not available")


def _check(kind, sam, ref, label, **kw):
    cls = {"NoDF": R.UMPAModelNoDF, "DF": R.UMPAModelDF, "DFKernel": R.UMPAModelDFKernel}[kind]
    ckw = {k: kw[k] for k in ("mask_list", "pos_list", "window_size", "max_shift") if k in kw}
    rm = cls([np.ascontiguousarray(s) for s in sam], [np.ascontiguousarray(r) for r in ref], **ckw)
    om = port.OracleModel(kind, sam, ref, **ckw)
    mkw = {}
    if kind == "DFKernel":
        mkw["abc"] = synth.blur_abc(*rm.sh)
    exp = rm.match(num_threads=2, quiet=True, **mkw)
    got = om.match(**mkw)
    assert tuple(om.extent) == tuple(rm.extent) and om.padding == rm.padding
    compare(got, exp, tol=1e-9, max_outliers=max(2, exp["err"].size // 40), outlier_tol=5e-3, label=label)
    ok = exp["err"] == 1
    np.testing.assert_allclose(got["debug_d"][ok], exp["debug_d"][ok], rtol=1e-10, atol=1e-12)
    # single-pixel hooks
    p = rm.padding
    for (i, j, si, sj) in ((p + 3, p + 5, 0, 1), (p + 8, p + 2, -1, 0)):
        a = rm.cost(i, j, si, sj, .5, .1, .4) if kind == "DFKernel" else rm.cost(i, j, si, sj)
        b, st = om.cost(i, j, si, sj, abc=(.5, .1, .4))
        np.testing.assert_allclose(b[:len(a)], a, rtol=1e-10)


@pytest.mark.parametrize("kind,Nw,ms", [("NoDF", 1, 3), ("DF", 2, 4), ("DF", 3, 5), ("DFKernel", 1, 3)])
def test_live_reference(kind, Nw, ms):
    d = synth.speckle_stack(4, 46 if kind != "DFKernel" else 56, 50 if kind != "DFKernel" else 58, seed=30 + Nw + ms,
                            max_shift=ms, dark_field=kind != "NoDF")
    _check(kind, d["sam"], d["ref"], "%s Nw=%d ms=%d" % (kind, Nw, ms), window_size=Nw, max_shift=ms)


def test_live_reference_masked_and_stepped():
    d = synth.speckle_stack(4, 48, 52, seed=41, max_shift=4, dark_field=True)
    rng = np.random.default_rng(2)
    mask = np.where(rng.random(d["sam"].shape) < .03, 0., 1.)
    _check("DF", d["sam"], d["ref"], "masked", mask_list=[m for m in mask], window_size=2, max_shift=4)
    pos = [(0, 0), (4, 0), (0, 6), (3, 3)]
    shp = [(44, 46), (42, 50), (46, 44), (40, 48)]
    sam = [d["sam"][k][py:py + h, px:px + w].copy() for k, ((py, px), (h, w)) in enumerate(zip(pos, shp))]
    ref = [d["ref"][k][py:py + h, px:px + w].copy() for k, ((py, px), (h, w)) in enumerate(zip(pos, shp))]
    _check("NoDF", sam, ref, "stepped", pos_list=[np.array(q) for q in pos], window_size=2, max_shift=4)


def test_hook_test_convolve_against_the_reference():
    """The product's numpy `test_convolve` hook (model.pyx:94-102 -> Utils.cpp:85-97) against the reference's own."""
    from umpa_b200 import model as pm
    rng = np.random.default_rng(5)
    img = np.ascontiguousarray(rng.random((40, 37)))
    for Nk in (1, 3, 8):
        ker = np.ascontiguousarray(rng.random((2 * Nk + 1, 2 * Nk + 1)))
        for (i, j) in ((Nk, Nk), (20, 18), (40 - Nk - 1, 37 - Nk - 1)):
            np.testing.assert_allclose(pm.test_convolve(img, i, j, ker), R.test_convolve(img, i, j, ker), rtol=1e-13)
