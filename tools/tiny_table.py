"""Smallest table-path run (golden df_clean) -- for compute-sanitizer / ncu on the GPU box."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_case          # noqa: E402
from gpu_common import run_case        # noqa: E402
import numpy as np                     # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "df_clean"
case = load_case(name)
m, got = run_case(case, "table")
exp = case["expected"]
print("err equal:", np.array_equal(got["err"], exp["err"]), "max |dx diff|:", np.abs(got["dx"] - exp["dx"])[exp["err"] == 1].max())
