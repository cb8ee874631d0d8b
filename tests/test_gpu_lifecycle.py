"""GPU: handle life cycle and threading, as INTEGRATION.md states them -- different handles are independent
(they share the per-device streams, the block cache and the pinned staging), a model per projection does not
grow device memory, and a handle can be matched again after its frames were replaced."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KEYS = ("f", "T", "dx", "dy", "err", "debug_Ncalls")


def _stack(seed, H=160, W=176, Na=6, ms=4):
    from umpa_b200 import synth
    return synth.speckle_stack(Na, H, W, seed=seed, max_shift=ms, dark_field=True)


def _same(a, b):
    for k in KEYS:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_handles_in_concurrent_threads_do_not_interfere():
    import umpa_b200
    jobs = []
    for t, (cls, kw) in enumerate(((umpa_b200.UMPAModelDF, {}), (umpa_b200.UMPAModelNoDF, {}),
                                   (umpa_b200.UMPAModelDF, {"window_size": 3}), (umpa_b200.UMPAModelDFKernel, {}))):
        d = _stack(100 + t, H=150 + 8 * t)
        jobs.append((cls, kw, d))

    def once(cls, kw, d):
        m = cls(list(d["sam"]), list(d["ref"]), max_shift=4, **kw)
        extra = {"abc": umpa_b200.synth.blur_abc(*m.sh)} if cls is umpa_b200.UMPAModelDFKernel else {}
        return m.match(quiet=True, **extra)

    serial = [once(*j) for j in jobs]
    results, errors = [None] * len(jobs), []

    def worker(n):
        try:
            for _ in range(6):
                results[n] = once(*jobs[n])
                _same(results[n], serial[n])
        except Exception as e:                  # noqa: BLE001 -- reported below, in the main thread
            errors.append((n, repr(e)))

    threads = [threading.Thread(target=worker, args=(n,)) for n in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for n in range(len(jobs)):
        _same(results[n], serial[n])


def test_model_per_projection_does_not_grow_device_memory():
    import torch
    import umpa_b200
    d = _stack(7, H=256, W=256, Na=8)
    used = []
    first = None
    for it in range(24):
        m = umpa_b200.UMPAModelDF(list(d["sam"]), list(d["ref"]), max_shift=4)
        res = m.match(quiet=True)
        if first is None:
            first = res
        else:
            _same(res, first)
        del m
        torch.cuda.synchronize()
        free, total = torch.cuda.mem_get_info()
        used.append(total - free)
    # the block cache fills during the first models and is reused afterwards
    assert max(used[8:]) <= used[7] + (8 << 20), used


def test_rematch_same_handle_and_other_options():
    import umpa_b200
    d = _stack(11)
    m = umpa_b200.UMPAModelDF(list(d["sam"]), list(d["ref"]), max_shift=4)
    a = m.match(quiet=True)
    b = m.match(quiet=True)
    _same(a, b)
    half = m.match(step=2, quiet=True)
    for k in KEYS:
        assert np.array_equal(half[k], a[k][::2, ::2], equal_nan=True), k
    m.assign_coordinates = "ref"
    r = m.match(step=1, quiet=True)
    assert m.last_match_info["path"] == "table" and (r["err"] == 1).mean() > .9
    m.assign_coordinates = "sam"
    _same(m.match(quiet=True), a)


def _f32_pair(seed, H, W, Na=8, pinned=False):
    """float32 stacks and their exact float64 widening"""
    import torch
    d = _stack(seed, H=H, W=W, Na=Na)
    s32, r32 = np.asarray(d["sam"], dtype=np.float32), np.asarray(d["ref"], dtype=np.float32)
    if pinned:                                  # one pinned allocation per stack: the 2-D copy route
        hs, hr = (torch.empty(s32.shape, dtype=torch.float32, pin_memory=True) for _ in range(2))
        hs.copy_(torch.from_numpy(s32)); hr.copy_(torch.from_numpy(r32))
        s32, r32 = hs.numpy(), hr.numpy()
    return s32, r32, s32.astype(np.float64), r32.astype(np.float64)


@pytest.mark.parametrize("H,W,pinned", [(700, 640, True), (700, 640, False), (531, 515, True)])
def test_float32_host_frames_equal_their_widening(H, W, pinned):
    """umpa_set_frames_f32: float32 frames cross PCIe as they are and are widened / centred on the GPU;
    every result equals the float64 route on the widened arrays bit for bit -- pipelined or not, table
    and lazy path, the cost() hook after a pipelined match (no FP64 copy on the device until then)."""
    import umpa_b200
    s32, r32, s64, r64 = _f32_pair(21, H, W, pinned=pinned)
    for cls in (umpa_b200.UMPAModelDF, umpa_b200.UMPAModelNoDF):
        m64 = cls(list(s64), list(r64), max_shift=4)
        want = m64.match(quiet=True)
        m32 = cls(list(s32), list(r32), max_shift=4)
        assert m32._f32
        got = m32.match(quiet=True)
        assert m32.last_match_info["path"] == "table" and m32.last_stream_info["bands"] > 1
        _same(got, want)
        assert np.array_equal(got["debug_d"], want["debug_d"], equal_nan=True)
        i, j = m32.padding + 11, m32.padding + 17
        assert m32.cost(i, j, 1, -2) == m64.cost(i, j, 1, -2)          # widens the FP64 stacks now
        m32.cuda_path = m64.cuda_path = "lazy"
        roi = ((5, 60, 3), (7, 90, 4))
        _same(m32.match(ROI=roi, quiet=True), m64.match(ROI=roi, quiet=True))
    # without the pipeline (everything uploaded before the first kernel)
    import os
    os.environ["UMPA_NO_STREAMING"] = "1"
    try:
        m32 = umpa_b200.UMPAModelDF(list(s32), list(r32), max_shift=4)
        _same(m32.match(quiet=True), umpa_b200.UMPAModelDF(list(s64), list(r64), max_shift=4).match(quiet=True))
    finally:
        del os.environ["UMPA_NO_STREAMING"]


def test_float32_host_frames_masked_and_dfkernel():
    import umpa_b200
    s32, r32, s64, r64 = _f32_pair(23, 300, 333, Na=6)
    mask = np.ones_like(s32)
    mask[:, 100:110, 50:70] = 0.
    mask[2, 200, 200] = .5
    m32 = umpa_b200.UMPAModelDF(list(s32), list(r32), mask_list=list(mask), max_shift=4)
    m64 = umpa_b200.UMPAModelDF(list(s64), list(r64), mask_list=list(mask.astype(np.float64)), max_shift=4)
    got, want = m32.match(quiet=True), m64.match(quiet=True)
    assert m32.last_match_info["path"] == "mixed"
    _same(got, want)
    k32 = umpa_b200.UMPAModelDFKernel(list(s32), list(r32), max_shift=4)
    k64 = umpa_b200.UMPAModelDFKernel(list(s64), list(r64), max_shift=4)
    abc = umpa_b200.synth.blur_abc(*k32.sh)
    _same(k32.match(abc=abc, quiet=True), k64.match(abc=abc, quiet=True))
