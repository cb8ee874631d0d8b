"""correct_bad_pixels (UMPA/align.py:661-732): the numpy restatement against vectors made by the
reference's own function (CPU), and the CUDA kernel against both (GPU, bit-exact)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import port

CASES = [("img", "img_th3_it1", dict(th=3)), ("img", "img_th3_it2", dict(th=3, iterations=2)),
         ("img", "img_th3_it3", dict(th=3, iterations=3)), ("img", "img_auto", dict()),
         ("img", "img_auto_p5", dict(p=5.)), ("stack", "stack_th4", dict(th=4))]


def _golden():
    return np.load(os.path.join(GOLDEN, "post_badpix.npz"))


@pytest.mark.parametrize("src,key,kw", CASES)
def test_oracle_equals_reference(src, key, kw):
    z = _golden()
    np.testing.assert_array_equal(port.correct_bad_pixels(z[src], **kw), z[key])


def test_oracle_nothing_to_correct():
    z = _golden()
    np.testing.assert_array_equal(port.correct_bad_pixels(np.clip(z["img"], -2, 2), 3), z["clean"])


@pytest.mark.gpu
@pytest.mark.parametrize("src,key,kw", CASES)
def test_cuda_equals_reference(src, key, kw):
    from umpa_b200 import align
    z = _golden()
    np.testing.assert_array_equal(align.correct_bad_pixels(z[src], **kw), z[key])


@pytest.mark.gpu
def test_cuda_edge_cases():
    import torch
    from umpa_b200 import align
    z = _golden()
    np.testing.assert_array_equal(align.correct_bad_pixels(np.clip(z["img"], -2, 2), 3), z["clean"])
    t = torch.as_tensor(z["img"]).cuda()
    out = align.correct_bad_pixels(t, 3)
    assert isinstance(out, torch.Tensor) and out.is_cuda
    np.testing.assert_array_equal(out.cpu().numpy(), z["img_th3_it1"])
    assert align.correct_bad_pixels(np.zeros((0, 5)), 3).shape == (0, 5)
    with pytest.raises(NotImplementedError):
        align.correct_bad_pixels(z["stack"], 4, dims=(0, 1))
    rng = np.random.default_rng(3)
    big = rng.normal(0, 2., (2034, 2034))
    np.testing.assert_array_equal(align.correct_bad_pixels(big, 5), port.correct_bad_pixels(big, 5))


@pytest.mark.gpu
def test_umpa_normal_and_nobias_wrappers():
    """align.py:12-117 on the GPU: match + (bias subtraction +) correction equal the same steps done with
    the reference-checked pieces one by one."""
    from umpa_b200 import UMPAModelDF, align, synth
    d = synth.speckle_stack(6, 80, 84, seed=8, max_shift=3, dark_field=True, noise=.3, amplitude=1.2)
    sam, ref = list(d["sam"]), list(d["ref"])
    plain = UMPAModelDF(sam, ref, window_size=1, max_shift=3).match(quiet=True)
    bias = UMPAModelDF(ref, ref, window_size=1, max_shift=3).match(quiet=True)
    n = align.UMPA_normal(sam, ref, window=1, shift=3)
    u = align.UMPA_nobias(sam, ref, window=1, shift=3)
    for k in ("dx", "dy"):
        np.testing.assert_array_equal(n[k], port.correct_bad_pixels(plain[k], 3))
        np.testing.assert_array_equal(u[k], port.correct_bad_pixels(plain[k] - bias[k], 3))
    for k in ("T", "df", "f", "err"):
        np.testing.assert_array_equal(n[k], plain[k])
    r = align.UMPA_normal(sam, ref, window=1, shift=3, ROI=(slice(2, 40, 2), slice(1, 60, 3)))
    assert r["dx"].shape == (19, 20)


@pytest.mark.gpu
def test_speckle_matching_wrappers():
    """UMPA/speckle_matching.py:12-75: match / match_unbiased are the model calls they wrap."""
    from umpa_b200 import UMPAModelDF, UMPAModelNoDF, match, match_unbiased, synth
    d = synth.speckle_stack(5, 64, 72, seed=6, max_shift=4, dark_field=True)
    sam, ref = list(d["sam"]), list(d["ref"])
    a = match(sam, ref, 2)
    b = UMPAModelDF(sam, ref, window_size=2).match(step=1, quiet=True)
    for k in ("dx", "dy", "T", "df", "f", "err"):
        np.testing.assert_array_equal(a[k], b[k])
    assert "df" not in match(sam, ref, 2, df=False) and UMPAModelNoDF is not None
    bias = UMPAModelDF(ref, ref, window_size=2).match(step=2, quiet=True)
    u = match_unbiased(sam, ref, 2, step=2)
    p = UMPAModelDF(sam, ref, window_size=2).match(step=2, quiet=True)
    np.testing.assert_array_equal(u["dx"], p["dx"] - bias["dx"])
    np.testing.assert_array_equal(u["dy"], p["dy"] - bias["dy"])
    np.testing.assert_array_equal(match_unbiased(sam, ref, 2, step=2, bias=False)["dx"], p["dx"])
    np.testing.assert_array_equal(match_unbiased(sam, ref, 2, step=2, bias=(1., 2.))["dy"], p["dy"] - 2.)
    nc = [np.asfortranarray(s) for s in sam]
    np.testing.assert_array_equal(match(nc, ref, 2)["dx"], a["dx"])
