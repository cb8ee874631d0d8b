"""GPU: BASELINE.json's configurations at FULL size, through size-independent properties (the
oracle would need minutes to hours here): pipelining invariance, ROI consistency, the FP64 lazy path
on a strided subsample, recovery of the synthetic truth, scaling laws of the model."""
import os

import numpy as np
import pytest

from helpers import compare_fp32

pytestmark = pytest.mark.gpu


def _stacks(Na, H, W, ms, dark_field, seed=2, pinned=False):
    import torch
    from umpa_b200 import synth
    d = synth.speckle_stack(Na, H, W, seed=seed, max_shift=ms, dark_field=dark_field, device="cuda", as_numpy=False)
    if not pinned:
        return d
    out = {}
    for k in ("sam", "ref"):
        h = torch.empty(d[k].shape, dtype=torch.float64, pin_memory=True)
        h.copy_(d[k])
        out[k] = h
    torch.cuda.synchronize()
    out.update({k: d[k].cpu().numpy() for k in ("dx", "dy", "T")})
    return out


def _env(**kw):
    class _E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return _E()


def test_cfg2_pipelining_invariance_and_truth():
    """Config 2 (DF 25 x 2048^2, Nw=2, max_shift=5) from pinned host frames: the banded, host-converted
    pipeline returns bit-identical maps to one unbanded upload; the synthetic displacement is recovered."""
    from umpa_b200 import UMPAModelDF
    d = _stacks(25, 2048, 2048, 5, True, pinned=True)
    sam, ref = list(d["sam"].numpy()), list(d["ref"].numpy())
    m = UMPAModelDF(sam, ref, window_size=2, max_shift=5)
    a = m.match(quiet=True, debug=False)
    info = m.last_stream_info
    assert m.last_match_info["path"] == "table" and info["bands"] > 1, info
    with _env(UMPA_BANDS=1, UMPA_HOST_THREADS=0):
        m1 = UMPAModelDF(sam, ref, window_size=2, max_shift=5)
        b = m1.match(quiet=True, debug=False)
        assert m1.last_stream_info["bands"] == 1 and m1.last_stream_info["host_threads"] == 0
    for k in ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls"):
        assert np.array_equal(a[k], b[k]), "%s differs between the pipelined and the single-band call" % k
    assert a["f"].shape == (2034, 2034) and (a["err"] == 1).mean() > .999
    ok = a["err"] == 1
    assert np.abs(a["dx"] - d["dx"][7:-7, 7:-7])[ok].mean() < .1
    assert np.abs(a["dy"] - d["dy"][7:-7, 7:-7])[ok].mean() < .1
    # a second match on the resident model (no upload) and a cropped one return the same pixels
    c = m.match(quiet=True, debug=False)
    r = m.match(ROI=((1000, 1040, 1), (3, 2000, 1)), quiet=True, debug=False)
    for k in ("f", "T", "dx", "dy", "df", "err"):
        assert np.array_equal(a[k], c[k])
        assert np.array_equal(r[k], a[k][1000:1040, 3:2000]), k
    # FP64 lazy path on a strided subsample of the same model (uploads the FP64 rows it needs)
    m.cuda_path = "lazy"
    exp = m.match(ROI=((5, 2034, 64), (9, 2034, 64)), quiet=True)
    got = {k: v[5::64, 9::64] for k, v in a.items()}
    st = compare_fp32(got, exp, label="cfg2 table-vs-lazy")
    assert st["n_ok"] == exp["err"].size


def test_cfg2_scaling_laws():
    """DF model: scaling the sample by s scales T by s and leaves dx, dy, df alone; scaling BOTH stacks
    leaves T alone and scales the cost by s^2 (Model.cpp:631-862 is homogeneous)."""
    from umpa_b200 import UMPAModelDF
    d = _stacks(25, 2048, 2048, 5, True)
    roi = ((0, 2034, 8), (0, 2034, 8))
    base = UMPAModelDF(list(d["sam"]), list(d["ref"]), window_size=2, max_shift=5).match(ROI=roi, quiet=True, debug=False)
    half = UMPAModelDF(list(d["sam"] * .5), list(d["ref"]), window_size=2, max_shift=5).match(ROI=roi, quiet=True, debug=False)
    both = UMPAModelDF(list(d["sam"] * 3.), list(d["ref"] * 3.), window_size=2, max_shift=5).match(ROI=roi, quiet=True, debug=False)
    same = (base["debug_Ncalls"] == half["debug_Ncalls"]) & (base["debug_Ncalls"] == both["debug_Ncalls"])
    assert same.mean() > .998                               # FP32 near-ties may walk differently
    np.testing.assert_allclose(half["T"][same], .5 * base["T"][same], rtol=2e-5)
    np.testing.assert_allclose(both["T"][same], base["T"][same], rtol=2e-5)
    np.testing.assert_allclose(both["f"][same], 9. * base["f"][same], rtol=2e-3, atol=1e-6 * np.median(base["f"]) * 9)
    for k in ("dx", "dy"):
        assert np.percentile(np.abs(half[k] - base[k])[same], 99.5) < 1e-4
        assert np.percentile(np.abs(both[k] - base[k])[same], 99.5) < 1e-4
    assert np.percentile(np.abs(half["df"] - base["df"])[same], 99.5) < 1e-4


def test_cfg3_dfkernel_blur_tables_vs_lazy():
    """Config 3 (DFKernel 25 x 2048^2, Nw=3, max_shift=5): the fused blur-table kernel against the FP64
    lazy path (the reference's arithmetic) on a strided subsample, full-size model."""
    from umpa_b200 import UMPAModelDFKernel, synth
    d = _stacks(25, 2048, 2048, 5, True)
    m = UMPAModelDFKernel(list(d["sam"]), list(d["ref"]), window_size=3, max_shift=5)
    N0, N1 = m.sh
    assert (N0, N1) == (2016, 2016)
    abc = synth.blur_abc(N0, N1)
    full = m.match(abc=abc, quiet=True, debug=False)
    assert m.last_match_info["path"] == "table" and (full["err"] == 1).mean() > .999
    m.cuda_path = "lazy"
    exp = m.match(ROI=((3, 2016, 48), (7, 2016, 48)), abc=np.ascontiguousarray(abc[3::48, 7::48]), quiet=True)
    got = {k: v[3::48, 7::48] for k, v in full.items()}
    st = compare_fp32(got, exp, label="cfg3 table-vs-lazy")
    assert st["n_ok"] > .99 * exp["err"].size


def test_dfkernel_assign_ref_blur_tables_vs_lazy():
    """DFKernel with assign_coordinates='ref' (the sample window moves): role-swapped blur tables against
    the FP64 lazy path on a strided subsample of a 12 x 640 x 768 model, two window sizes."""
    from umpa_b200 import UMPAModelDFKernel, synth
    d = _stacks(12, 640, 768, 4, True)
    for Nw, ms in ((2, 4), (3, 5)):
        m = UMPAModelDFKernel(list(d["sam"]), list(d["ref"]), window_size=Nw, max_shift=ms)
        m.assign_coordinates = "ref"
        N0, N1 = m.sh
        abc = synth.blur_abc(N0, N1)
        full = m.match(abc=abc, quiet=True)
        assert m.last_match_info["path"] == "table" and (full["err"] == 1).mean() > .99
        m.cuda_path = "lazy"
        exp = m.match(ROI=((1, N0, 16), (5, N1, 16)), abc=np.ascontiguousarray(abc[1::16, 5::16]), quiet=True)
        got = {k: v[1::16, 5::16] for k, v in full.items()}
        st = compare_fp32(got, exp, label="dfk assign=ref table-vs-lazy Nw=%d" % Nw)
        assert st["n_ok"] > .98 * exp["err"].size


def test_cfg4_large_field_roi_consistency():
    """Config 4 (DF 40 x 4096^2, Nw=3, max_shift=8; 5.4 GB of FP32 stacks, 2 x 15 GB of tables): full
    match on the device, a cropped match and the FP64 lazy path on a strided subsample agree."""
    from umpa_b200 import UMPAModelDF
    d = _stacks(40, 4096, 4096, 8, True)
    m = UMPAModelDF(list(d["sam"]), list(d["ref"]), window_size=3, max_shift=8)
    full = m.match_device()
    assert m.last_match_info["path"] == "table" and tuple(full["f"].shape) == (4074, 4074)
    assert float((full["err"] == 1).double().mean()) > .999
    r = m.match(ROI=((2000, 2016, 1), (100, 4000, 1)), quiet=True, debug=False)
    for k in ("f", "T", "dx", "dy", "df"):
        assert np.array_equal(r[k], full[k][2000:2016, 100:4000].cpu().numpy()), k
    got = {k: v[11::128, 5::128].cpu().numpy() for k, v in full.items() if hasattr(v, "cpu")}
    del full
    m.cuda_path = "lazy"
    exp = m.match(ROI=((11, 4074, 128), (5, 4074, 128)), quiet=True)
    compare_fp32(got, exp, label="cfg4 table-vs-lazy")


def test_cfg5_large_window_vs_lazy():
    """Config 5 (NoDF 4 x 2048^2, Nw=6, default max_shift=4): window-dominated regime."""
    from umpa_b200 import UMPAModelNoDF
    d = _stacks(4, 2048, 2048, 4, False)
    m = UMPAModelNoDF(list(d["sam"]), list(d["ref"]), window_size=6)
    full = m.match(quiet=True, debug=False)
    assert full["f"].shape == (2028, 2028) and m.last_match_info["path"] == "table"
    m.cuda_path = "lazy"
    exp = m.match(ROI=((1, 2028, 32), (2, 2028, 32)), quiet=True)
    got = {k: v[1::32, 2::32] for k, v in full.items()}
    compare_fp32(got, exp, label="cfg5 table-vs-lazy")


def test_cfg2_masked_mixed_path():
    """Config 2 with a realistic mask stack (ones, a few dead pixels, a dead block): pixels with no mask value
    != 1 within reach are bit-identical to the unmasked model (table kernels), the others equal the FP64 lazy
    path (the reference's masked arithmetic)."""
    import time
    import torch
    from umpa_b200 import UMPAModelDF
    d = _stacks(25, 2048, 2048, 5, True)
    sam, ref = list(d["sam"]), list(d["ref"])
    mask = torch.ones((25, 2048, 2048), dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cpu").manual_seed(3)
    for _ in range(60):
        k, y, x = (int(torch.randint(0, n, (1,), generator=g)) for n in (25, 2048, 2048))
        mask[k, y, x] = 0.
    mask[7, 900:920, 1100:1130] = 0.
    plain = UMPAModelDF(sam, ref, window_size=2, max_shift=5).match(quiet=True, debug=False)
    m = UMPAModelDF(sam, ref, mask_list=list(mask), window_size=2, max_shift=5)
    got = m.match(quiet=True, debug=False)
    assert m.last_match_info["path"] == "mixed"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    got = m.match(quiet=True, debug=False)
    t_mixed = time.perf_counter() - t0
    bad = (mask != 1.).any(dim=0)[None, None].double()
    reach = torch.nn.functional.max_pool2d(bad, 15, stride=1, padding=7)[0, 0][7:-7, 7:-7].cpu().numpy() > 0
    assert 0.001 < reach.mean() < .05
    for k in ("f", "T", "dx", "dy", "df", "err"):
        assert np.array_equal(got[k][~reach], plain[k][~reach]), k
    m.cuda_path = "lazy"
    roi = ((880, 940, 1), (1080, 1150, 1))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    exp = m.match(ROI=roi, quiet=True, debug=False)
    t_lazy_roi = time.perf_counter() - t0
    sub = reach[880:940, 1080:1150]
    assert sub.mean() > .2
    for k in ("f", "T", "dx", "dy", "df", "err"):
        assert np.array_equal(got[k][880:940, 1080:1150][sub], exp[k][sub]), k
    print("masked config 2: mixed path %.1f ms (host maps included); lazy path on a %d-pixel ROI %.1f ms"
          % (1e3 * t_mixed, exp["err"].size, 1e3 * t_lazy_roi))


def test_pipelined_variants_bit_identical():
    """The banded, host-converted call with a strided ROI / a start guess / the DFKernel model returns the
    bits of the corresponding device-resident call (host constants in both: same centred FP32 stacks)."""
    import torch
    from umpa_b200 import UMPAModelDF, UMPAModelDFKernel, synth
    d = _stacks(8, 1536, 1280, 5, True, pinned=True)
    sam, ref = list(d["sam"].numpy()), list(d["ref"].numpy())
    with _env(UMPA_BANDS=1, UMPA_HOST_THREADS=0):
        base = UMPAModelDF(sam, ref, window_size=2, max_shift=5)
        full = base.match(quiet=True, debug=False)
    m = UMPAModelDF(sam, ref, window_size=2, max_shift=5)
    got = m.match(step=3, quiet=True, debug=False)
    assert m.last_stream_info["bands"] > 1 and m.last_stream_info["host_threads"] > 0, m.last_stream_info
    for k in ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls"):
        assert np.array_equal(got[k], full[k][::3, ::3]), k
    m = UMPAModelDF(sam, ref, window_size=2, max_shift=5)
    g2 = m.match(dxdy=(1., -1.), ROI=((10, 1500, 1), (4, 1200, 1)), quiet=True, debug=False)
    with _env(UMPA_BANDS=1, UMPA_HOST_THREADS=0):
        b2 = UMPAModelDF(sam, ref, window_size=2, max_shift=5).match(dxdy=(1., -1.), ROI=((10, 1500, 1), (4, 1200, 1)),
                                                                      quiet=True, debug=False)
    for k in ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls"):
        assert np.array_equal(g2[k], b2[k]), k
    # DFKernel, 3 bands + host conversion vs one band
    sam3, ref3 = [s[:, :640] .copy() for s in sam[:4]], [r[:, :640].copy() for r in ref[:4]]
    hs = [torch.from_numpy(x).pin_memory().numpy() for x in sam3]
    hr = [torch.from_numpy(x).pin_memory().numpy() for x in ref3]
    mk = UMPAModelDFKernel(hs, hr, window_size=2, max_shift=4)
    abc = synth.blur_abc(*mk.sh)
    gk = mk.match(abc=abc, quiet=True, debug=False)
    assert mk.last_match_info["path"] == "table" and mk.last_stream_info["bands"] > 1
    with _env(UMPA_BANDS=1, UMPA_HOST_THREADS=0):
        bk = UMPAModelDFKernel(hs, hr, window_size=2, max_shift=4).match(abc=abc, quiet=True, debug=False)
    for k in ("f", "T", "dx", "dy", "err", "debug_Ncalls"):
        assert np.array_equal(gk[k], bk[k]), k


def test_small_and_odd_shapes():
    """Edge shapes through the table path: one frame, frames barely larger than the padding, widths that are
    not multiples of 4, the smallest shift range -- against the lazy path."""
    from umpa_b200 import UMPAModelDF, UMPAModelNoDF, synth
    for (Na, H, W, Nw, ms, cls) in ((3, 40, 43, 1, 2, UMPAModelNoDF), (1, 30, 31, 2, 2, UMPAModelNoDF), (2, 17, 19, 2, 2, UMPAModelDF),
                                    (3, 64, 70, 1, 3, UMPAModelNoDF), (5, 33, 129, 3, 4, UMPAModelDF)):
        d = synth.speckle_stack(Na, max(H, 64), max(W, 64), seed=Na + H, max_shift=max(ms, 3),
                                dark_field=cls is UMPAModelDF, amplitude=.4)
        sam = [np.ascontiguousarray(x[:H, :W]) for x in d["sam"]]
        ref = [np.ascontiguousarray(x[:H, :W]) for x in d["ref"]]
        m = cls(sam, ref, window_size=Nw, max_shift=ms)
        m.cuda_path = "lazy"
        exp = m.match(quiet=True)
        m.cuda_path = "table"
        got = m.match(quiet=True)
        assert m.last_match_info["path"] == "table" and got["f"].shape == exp["f"].shape
        compare_fp32(got, exp, label="shape %s" % ((Na, H, W, Nw, ms),), max_exception_frac=.1)
    # a 1x1 window on 3 frames is left to the FP64 path (too few samples per cost for FP32 sums)
    d = synth.speckle_stack(3, 64, 70, seed=1, max_shift=3, dark_field=False, amplitude=.4)
    m = UMPAModelNoDF(list(d["sam"]), list(d["ref"]), window_size=0, max_shift=3)
    m.match(quiet=True)
    assert m.last_match_info["path"] == "lazy"


def test_pageable_frames_go_through_host_conversion():
    """Ordinary (pageable) numpy frames: every row is converted by the host threads into pinned staging
    (the driver's own staging of pageable FP64 would be the slow path); same bits as pinned frames."""
    from umpa_b200 import UMPAModelDF
    d = _stacks(6, 1536, 1280, 5, True, pinned=True)
    pinned_s, pinned_r = list(d["sam"].numpy()), list(d["ref"].numpy())
    a = UMPAModelDF(pinned_s, pinned_r, window_size=2, max_shift=5).match(quiet=True, debug=False)
    pag_s, pag_r = [np.array(x, copy=True) for x in pinned_s], [np.array(x, copy=True) for x in pinned_r]
    m = UMPAModelDF(pag_s, pag_r, window_size=2, max_shift=5)
    b = m.match(quiet=True, debug=False)
    info = m.last_stream_info
    assert info["bands"] > 1 and info["host_threads"] > 0 and info["host_rows_per_frame"] == 1536, info
    for k in ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls"):
        assert np.array_equal(a[k], b[k]), k
    # the FP64 stacks are fetched when a hook needs them
    v = m.min(m.padding + 10, m.padding + 20)
    assert abs(v[2] - b["dx"][10, 20]) < 1e-3
