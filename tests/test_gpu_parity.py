"""GPU: the CUDA paths (through the ctypes C-ABI) against the reference's golden vectors."""
import numpy as np
import pytest

from helpers import compare, golden_names, load_case
from gpu_common import TABLE_CASES, product_model, run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_lazy_path_equals_reference(name):
    """FP64 lazy path: same arithmetic as the reference -> tight tolerance (Newton-stop outliers
    as documented in test_oracle_golden.py)."""
    case = load_case(name)
    m, got = run_case(case, "lazy")
    assert m.last_match_info["path"] == "lazy"
    exp = case["expected"]
    compare(got, exp, tol=1e-9, max_outliers=max(2, exp["err"].size // 40), outlier_tol=5e-3, label=name)
    ok = exp["err"] == 1
    np.testing.assert_allclose(got["debug_d"][ok], exp["debug_d"][ok], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["debug_a"][ok], exp["debug_a"][ok], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", TABLE_CASES)
def test_table_path_equals_reference(name):
    """FP32-table path: err map and integer walk equal except a documented handful of FP32
    near-ties on the noisy sets; dx, dy, T, df, f within 1e-4 (north_star tolerance)."""
    case = load_case(name)
    m, got = run_case(case, "table")
    assert m.last_match_info["path"] == "table"
    exp = case["expected"]
    noisy = "noisy" in name
    n = exp["err"].size
    compare(got, exp, tol=1e-4, max_err_mismatch=n // 200 if noisy else 0,
            max_outliers=max(2, n // 50) if noisy else max(1, n // 500), outlier_tol=5e-2, label=name)


@pytest.mark.parametrize("name", ["nodf_clean", "df_clean", "dfk_clean", "df_masked", "df_positions"])
def test_cost_probes(name):
    case = load_case(name)
    m = product_model(case, "lazy")
    for i, j, si, sj, f, t, v in case["cost_probes"]:
        if case["kind"] == "DFKernel":
            c = m.cost(int(i), int(j), si, sj, .5, .1, .4)
        else:
            c = m.cost(int(i), int(j), si, sj)
        np.testing.assert_allclose(c, [f, t, v][:len(c)], rtol=1e-10)


def test_min_hook_matches_match():
    case = load_case("df_clean")
    m, got = run_case(case, "lazy")
    p = m.padding
    for (xi, xj) in ((0, 0), (5, 7), (20, 30)):
        v = m.min(p + xi, p + xj)
        np.testing.assert_allclose(v, [got[k][xi, xj] for k in ("f", "T", "dx", "dy", "df")], rtol=1e-12)


def test_coverage_and_geometry():
    case = load_case("df_positions")
    m = product_model(case, "lazy")
    assert tuple(m.extent) == tuple(int(v) for v in case["extent"])
    cov = m.coverage()
    assert cov.shape == tuple(case["expected"]["err"].shape)
    from oracle import port
    o = port.OracleModel("DF", case["sam"], case["ref"], pos_list=case["pos"], window_size=case["Nw"],
                         max_shift=case["max_shift"])
    np.testing.assert_array_equal(cov, o.coverage(((0, cov.shape[0], 1), (0, cov.shape[1], 1))))
