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
