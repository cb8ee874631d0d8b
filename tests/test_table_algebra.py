"""CPU: the algebraic regrouping behind the table path is exact.

tests/table_algebra.py evaluates every shift's cost the way the CUDA kernels do (centred frames,
frame-summed unwindowed correlation + ONE separable window filter per shift, per-frame filtered
images for the mean term, FP64 solve).  In float64 it must reproduce the oracle's direct
evaluation to rounding; in float32 it bounds the error the kernels can have."""
import numpy as np
import pytest

from helpers import load_case
from oracle import port
from table_algebra import cost_tables


@pytest.mark.parametrize("name", ["nodf_clean", "df_clean", "df_lowcontrast", "df_nw3_ms6"])
def test_regrouped_cost_equals_direct(name):
    c = load_case(name)
    sam, ref = np.array(c["sam"]), np.array(c["ref"])
    df = c["kind"] == "DF"
    o = port.OracleModel(c["kind"], c["sam"], c["ref"], window_size=c["Nw"], max_shift=c["max_shift"])
    h, p = c["max_shift"] - 1, c["padding"]
    N0, N1 = o.extent
    cost64, T64, D64 = cost_tables(sam, ref, c["Nw"], c["max_shift"], df, np.float64)
    cost32, T32, D32 = cost_tables(sam, ref, c["Nw"], c["max_shift"], df, np.float32)
    rng = np.random.default_rng(0)
    scale = np.median(cost64)
    for _ in range(60):
        i, j = int(rng.integers(0, N0)), int(rng.integers(0, N1))
        si, sj = int(rng.integers(-h, h + 1)), int(rng.integers(-h, h + 1))
        (f, t, v), st = o.cost(p + i, p + j, si, sj)
        assert st == 1
        assert abs(cost64[si + h, sj + h, i, j] - f) <= 1e-12 * max(scale, abs(f))
        assert abs(T64[si + h, sj + h, i, j] - t) <= 1e-11 * abs(t)
        if df:
            assert abs(D64[si + h, sj + h, i, j] - v) <= 1e-10 * max(1., abs(v))
        # FP32 accumulation: error stays at the 1e-5 level of the cost scale
        assert abs(cost32[si + h, sj + h, i, j] - f) <= 3e-5 * scale
