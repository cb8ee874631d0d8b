"""ctypes front-end of oracle/libumpa_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY -- see the header of umpa_oracle.c.  The class mirrors
the constructor / match / cost / min / coverage surface of the reference's
``UMPA/model.pyx`` closely enough that parity tests can run the same call on the
oracle, on the compiled reference (oracle/ref.py) and on the CUDA product.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_KINDS = {"NoDF": 0, "DF": 1, "DFKernel": 2}
_NPARAM = {"NoDF": 4, "DF": 5, "DFKernel": 7}
_SAFE_CROP = {"NoDF": 0, "DF": 0, "DFKernel": 8}    # model.pyx:762,828,904

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build_port()
        L = C.CDLL(path)
        dp, ip, pp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_void_p)
        L.uo_create.restype = C.c_void_p
        L.uo_create.argtypes = [C.c_int, C.c_int, ip, ip, pp, pp, pp, C.c_int, dp, C.c_int, C.c_int]
        L.uo_destroy.argtypes = [C.c_void_p]
        L.uo_set_window.argtypes = [C.c_void_p, C.c_int, dp]
        L.uo_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.uo_coverage.restype = C.c_double
        L.uo_coverage.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.uo_make_kernel.argtypes = [C.c_double, C.c_double, C.c_double, dp]
        L.uo_cost.restype = C.c_int
        L.uo_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]
        L.uo_spmin.restype = C.c_double
        L.uo_spmin.argtypes = [dp, dp]
        L.uo_spmin_quad.restype = C.c_double
        L.uo_spmin_quad.argtypes = [dp, dp]
        L.uo_min.restype = C.c_int
        L.uo_min.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, dp, dp, ip]
        L.uo_match.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               dp, C.c_double, C.c_int, dp, dp, ip, dp, dp, ip, C.c_int]
        L.uo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def make_window(Nw):
    """model.pyx:691-696"""
    h = np.hamming(2 * Nw + 1)
    w = np.multiply.outer(h, h)
    w /= w.sum()
    return np.ascontiguousarray(w, dtype=np.float64)


def spmin(a, pos=(0., 0.)):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(16)
    p = np.array(pos, dtype=np.float64)
    v = lib().uo_spmin(_dp(a), _dp(p))
    return p, v


def spmin_quad(a, pos=(0., 0.)):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(16)
    p = np.array(pos, dtype=np.float64)
    v = lib().uo_spmin_quad(_dp(a), _dp(p))
    return p, v


def blur_kernel(a, b, c):
    k = np.empty((17, 17), dtype=np.float64)
    lib().uo_make_kernel(a, b, c, _dp(k))
    return k


class OracleModel:
    def __init__(self, kind, sam_list, ref_list, mask_list=None, pos_list=None,
                 window_size=2, max_shift=4):
        self.kind = kind
        self.Nparam = _NPARAM[kind]
        self.sam = [np.ascontiguousarray(s, dtype=np.float64) for s in sam_list]
        self.ref = [np.ascontiguousarray(r, dtype=np.float64) for r in ref_list]
        self.mask = None if mask_list is None else [np.ascontiguousarray(m, dtype=np.float64) for m in mask_list]
        self.Na = len(self.sam)
        self.dim = np.array([s.shape for s in self.sam], dtype=np.int32).reshape(self.Na, 2)
        if pos_list is None:
            self.pos = np.zeros((self.Na, 2), dtype=np.int32)
        else:
            self.pos = np.ascontiguousarray(np.array(pos_list), dtype=np.int32).reshape(self.Na, 2)
        self.Nw = int(window_size)
        self.max_shift = int(max_shift)
        self.padding = self.max_shift + self.Nw + _SAFE_CROP[kind]          # model.pyx:286
        self.window = make_window(self.Nw)

        def parr(frames):
            arr = (C.c_void_p * self.Na)(*[f.ctypes.data for f in frames])
            return C.cast(arr, C.POINTER(C.c_void_p))
        self._h = lib().uo_create(_KINDS[kind], self.Na, _ip(self.dim), _ip(self.pos),
                                  parr(self.sam), parr(self.ref),
                                  parr(self.mask) if self.mask is not None else None,
                                  self.Nw, _dp(self.window), self.max_shift, self.padding)
        self.sub_pixel_mode = -1
        self.reference_shift = 0

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.uo_destroy(self._h)
            self._h = None

    def set_options(self, sub_pixel_mode=-1, reference_shift=0):
        self.sub_pixel_mode, self.reference_shift = int(sub_pixel_mode), int(reference_shift)
        lib().uo_set_options(self._h, self.sub_pixel_mode, self.reference_shift)

    def set_Nw(self, Nw):
        self.Nw = int(Nw)
        self.window = make_window(self.Nw)
        lib().uo_set_window(self._h, self.Nw, _dp(self.window))

    @property
    def extent(self):
        """model.pyx:531-549"""
        pmax = np.max(self.pos + self.dim, axis=0)
        return int(pmax[0] - 2 * self.padding), int(pmax[1] - 2 * self.padding)

    def cost(self, i, j, sx, sy, abc=(0., 0., 0.)):
        vals = np.zeros(3)
        abc = np.array(abc, dtype=np.float64)
        st = lib().uo_cost(self._h, int(i), int(j), int(round(sx)), int(round(sy)), _dp(abc), _dp(vals))
        return vals, st

    def min(self, i, j, abc=None, uv=(0., 0.)):
        vals = np.zeros(self.Nparam)
        if abc is not None:
            vals[4:7] = abc
        uv = np.array(uv, dtype=np.float64)
        d, a, n = np.zeros(25), np.zeros(16), C.c_int(0)
        ok = lib().uo_min(self._h, int(i), int(j), _dp(vals), _dp(uv), _dp(d), _dp(a), C.byref(n))
        return vals, ok, d, a, n.value

    def coverage(self, ROI):
        (s0, e0, t0), (s1, e1, t1) = ROI
        N0, N1 = 1 + (e0 - s0 - 1) // t0, 1 + (e1 - s1 - 1) // t1
        out = np.zeros((N0, N1))
        for xi in range(N0):
            for xj in range(N1):
                out[xi, xj] = lib().uo_coverage(self._h, self.padding + s0 + t0 * xi, self.padding + s1 + t1 * xj)
        return out

    def match(self, ROI=None, dxdy=None, abc=None, num_threads=0, debug=True, gate=True):
        """Pixel loop of model.pyx:334-497 over ROI=((start,stop,step),(start,stop,step))."""
        if ROI is None:
            N0, N1 = self.extent
            ROI = ((0, N0, 1), (0, N1, 1))
        (s0, e0, t0), (s1, e1, t1) = ROI
        N0, N1 = 1 + (e0 - s0 - 1) // t0, 1 + (e1 - s1 - 1) // t1
        values = np.zeros((N0, N1, self.Nparam))
        if self.kind == "DFKernel":
            values[:, :, 4:7] = abc
        uv = np.zeros((N0, N1, 2))
        if dxdy is not None:
            uv[:, :, 0] = dxdy[0]
            uv[:, :, 1] = dxdy[1]
        err = np.zeros((N0, N1), dtype=np.int32)
        dd = np.zeros((N0, N1, 25)) if debug else None
        da = np.zeros((N0, N1, 16)) if debug else None
        nc = np.zeros((N0, N1), dtype=np.int32)
        cover, thr = None, 0.
        if gate and (self.mask is not None or np.any(self.pos != 0)):
            cover = self.coverage(ROI)
            thr = .1 * cover.max() / self.Na                               # model.pyx:431
        lib().uo_match(self._h, self.padding + s0, t0, N0, self.padding + s1, t1, N1,
                       _dp(cover) if cover is not None else None, thr, self.Nparam,
                       _dp(values), _dp(uv), _ip(err),
                       _dp(dd) if debug else None, _dp(da) if debug else None, _ip(nc), int(num_threads))
        out = {"f": values[:, :, 0].copy(), "T": values[:, :, 1].copy(),
               "dx": values[:, :, 2].copy(), "dy": values[:, :, 3].copy(),
               "err": err, "debug_Ncalls": nc}
        if self.kind == "DF":
            out["df"] = values[:, :, 4].copy()
        if debug:
            out["debug_d"], out["debug_a"] = dd, da
        return out


def max_threads():
    return lib().uo_max_threads()


def correct_bad_pixels(img_in, th=None, iterations=1, p=0.5):
    """Restatement of UMPA/align.py:661-732 for dims=(-2, -1) (numpy, test infrastructure only).
    Values outside [-th, th] (or the p / 100-p percentiles, align.py:702-705) are replaced by the
    median of their four neighbours; |i-1| reflects at the low edge, N-2 replaces N at the high
    edge (align.py:720-727); every iteration gathers all neighbours before it assigns (713-731) and
    revisits the SAME pixels (the mask is taken once, align.py:707-708)."""
    img = np.array(img_in, copy=True)
    lims = [np.percentile(img, p), np.percentile(img, 100 - p)] if th is None else [-th, th]
    bad = (img < min(lims)) | (img > max(lims))
    if not bad.any():
        return img
    N0, N1 = img.shape[-2:]
    idx = np.nonzero(bad)
    i, j = idx[-2], idx[-1]
    lead = idx[:-2]
    for _ in range(iterations):
        nb = np.stack([img[lead + (np.abs(i - 1), j)], img[lead + (np.where(i + 1 == N0, N0 - 2, i + 1), j)],
                       img[lead + (i, np.abs(j - 1))], img[lead + (i, np.where(j + 1 == N1, N1 - 2, j + 1))]])
        nb.sort(axis=0)
        img[idx] = (nb[1] + nb[2]) / 2.
    return img
