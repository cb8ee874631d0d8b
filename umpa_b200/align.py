"""The callers of the matching path in the reference's ``UMPA/align.py`` that real pipelines use:
``correct_bad_pixels`` (align.py:661-732) and the two helpers ``UMPA_normal`` / ``UMPA_nobias``
(align.py:12-117), with the reference's signatures.  The correction runs on the GPU
(``umpa_correct_bad_pixels``, post.cu); in the two helpers the displacement maps go from the match
kernels to the correction (and the bias subtraction) without leaving device memory."""
import ctypes as C

import numpy as np
import torch

from . import _capi, model


def _device_correct(img_t, lo, hi, iterations, bias_t=None):
    """img_t: CUDA float64 tensor (..., N0, N1); returns a new tensor."""
    img_t = img_t.contiguous()
    N0, N1 = int(img_t.shape[-2]), int(img_t.shape[-1])
    nimg = int(img_t.numel() // max(1, N0 * N1))
    out = torch.empty_like(img_t)
    need_scratch = bias_t is not None or iterations > 1
    scratch = torch.empty_like(img_t) if need_scratch else None
    if bias_t is not None:
        bias_t = bias_t.contiguous()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _capi.check(_capi.lib().umpa_correct_bad_pixels(
        C.c_void_p(img_t.data_ptr()), C.c_void_p(bias_t.data_ptr()) if bias_t is not None else None,
        C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()) if scratch is not None else None,
        nimg, N0, N1, float(lo), float(hi), int(iterations), st))
    return out


def correct_bad_pixels(img_in, th=None, iterations=1, dims=(-2, -1), p=0.5):
    """align.py:661-732.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.  ``dims`` must be
    the last two axes (the reference's default; what UMPA_normal / UMPA_nobias use)."""
    nd = img_in.ndim
    if tuple(sorted(d % nd for d in dims)) != (nd - 2, nd - 1):
        raise NotImplementedError("correct_bad_pixels on the GPU works along the last two axes (dims=(-2, -1))")
    is_t = isinstance(img_in, torch.Tensor)
    t = img_in if is_t else torch.as_tensor(np.ascontiguousarray(img_in, dtype=np.float64))
    t = t.to(device="cuda", dtype=torch.float64)
    if th is None:                                  # align.py:702-703
        h = t.detach().cpu().numpy()
        lims = [np.percentile(h, p), np.percentile(h, 100 - p)]
    elif np.ndim(th) == 0:
        lims = [-th, th]                            # align.py:705
    else:
        raise TypeError("th must be a scalar or None (the reference negates it, align.py:705)")
    lo, hi = min(lims), max(lims)
    if t.numel() == 0 or not bool(((t < lo) | (t > hi)).any()):      # align.py:709-711
        out = t.clone()
    else:
        out = _device_correct(t, lo, hi, iterations)
    return out if is_t else out.cpu().numpy().astype(img_in.dtype, copy=False)


def _build(cls, sams, refs, window, shift, pos_list, mask_list):
    return cls(sams, refs, window_size=window, max_shift=shift, pos_list=pos_list, mask_list=mask_list)


def _roi(m, ROI):
    if ROI == (slice(None, None, None), slice(None, None, None)):
        return None
    return ROI


def _finish(m, dev, dx, dy):
    dev.pop("_keepalive", None)
    res = {k: v.cpu().numpy() for k, v in dev.items()}
    res["dx"], res["dy"] = dx.cpu().numpy(), dy.cpu().numpy()
    return res


def UMPA_normal(sams, refs, window=1, shift=3, pos_list=None, mask_list=None, assign_coordinates='sam',
                num_threads=None, ROI=(slice(None, None, None), slice(None, None, None))):
    """align.py:12-60: UMPAModelDF.match + bad-pixel correction (threshold = shift) of dx and dy."""
    m = _build(model.UMPAModelDF, sams, refs, window, shift, pos_list, mask_list)
    m.assign_coordinates = assign_coordinates
    dev = m.match_device(ROI=_roi(m, ROI), debug=model.DEBUG)
    dx = _device_correct(dev["dx"], -shift, shift, 1) if dev["dx"].numel() else dev["dx"]
    dy = _device_correct(dev["dy"], -shift, shift, 1) if dev["dy"].numel() else dev["dy"]
    return _finish(m, dev, dx, dy)


def UMPA_nobias(sams, refs, window=1, shift=3, pos_list=None, mask_list=None, assign_coordinates='sam',
                num_threads=None, ROI=(slice(None, None, None), slice(None, None, None))):
    """align.py:62-117: as UMPA_normal, after subtracting the bias found by matching refs against refs
    (the bias model keeps the default assign_coordinates, like the reference)."""
    m = _build(model.UMPAModelDF, sams, refs, window, shift, pos_list, mask_list)
    b = _build(model.UMPAModelDF, refs, refs, window, shift, pos_list, mask_list)
    m.assign_coordinates = assign_coordinates
    dev = m.match_device(ROI=_roi(m, ROI), debug=model.DEBUG)
    devb = b.match_device(ROI=_roi(b, ROI), debug=False)
    if dev["dx"].numel():
        dx = _device_correct(dev["dx"], -shift, shift, 1, bias_t=devb["dx"])
        dy = _device_correct(dev["dy"], -shift, shift, 1, bias_t=devb["dy"])
    else:
        dx, dy = dev["dx"], dev["dy"]
    return _finish(m, dev, dx, dy)
