"""CPU: the C restatement (oracle/) against golden vectors made by the compiled reference."""
import numpy as np
import pytest

from helpers import compare, golden_names, load_case, roi_of, GOLDEN
from oracle import port


def _model(c):
    m = port.OracleModel(c["kind"], c["sam"], c["ref"], mask_list=c["mask"], pos_list=c["pos"],
                         window_size=c["Nw"], max_shift=c["max_shift"])
    m.set_options(sub_pixel_mode=-1 if c["subpx"] is None else c["subpx"],
                  reference_shift=1 if c["assign"] == "ref" else 0)
    return m


@pytest.mark.parametrize("name", golden_names())
def test_match_equals_reference(name):
    c = load_case(name)
    m = _model(c)
    assert m.padding == c["padding"]
    assert tuple(m.extent) == tuple(int(v) for v in c["extent"])
    np.testing.assert_allclose(m.window, c["window"], rtol=0, atol=1e-15)
    got = m.match(ROI=roi_of(c), dxdy=c["dxdy"], abc=c["abc"])
    exp = c["expected"]
    # Newton in spmin stops on an absolute step (Optim.cpp:123): in a few ill-conditioned noisy
    # pixels the -ffast-math reference and this strict build stop one iteration apart.
    st = compare(got, exp, tol=1e-9, max_outliers=max(2, exp["err"].size // 40), outlier_tol=5e-3, label=name)
    ok = exp["err"] == 1
    np.testing.assert_allclose(got["debug_d"][ok], exp["debug_d"][ok], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(got["debug_a"][ok], exp["debug_a"][ok], rtol=1e-10, atol=1e-12)
    # failed pixels: T/df come from the last successful cost call, dx/dy from the walk
    bad = (exp["err"] == 0) & (exp["debug_Ncalls"] > 0)
    if bad.any():
        np.testing.assert_array_equal(got["debug_Ncalls"][bad], exp["debug_Ncalls"][bad])
        np.testing.assert_allclose(got["T"][bad], exp["T"][bad], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got["dx"][bad], exp["dx"][bad], rtol=0, atol=1e-12)
        np.testing.assert_allclose(got["dy"][bad], exp["dy"][bad], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", golden_names())
def test_cost_probes(name):
    c = load_case(name)
    m = _model(c)
    for i, j, si, sj, f, t, v in c["cost_probes"]:
        vals, st = m.cost(int(i), int(j), si, sj, abc=(.5, .1, .4))
        assert st == 1
        n = 3 if c["kind"] == "DF" else 2
        np.testing.assert_allclose(vals[:n], [f, t, v][:n], rtol=1e-10)


def test_hooks():
    z = np.load(GOLDEN + "/hooks.npz")
    for row in z["blocks"]:
        a = row[:16]
        pq, vq = port.spmin_quad(a)
        ps, vs = port.spmin(a)
        np.testing.assert_allclose(np.r_[pq, vq], row[16:19], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(np.r_[ps, vs], row[19:22], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(port.blur_kernel(*z["kernel_abc"]), z["kernel"], rtol=1e-12, atol=1e-300)
