"""Drop-in host classes for the reference's ``UMPA/model.pyx``.

``UMPAModelNoDF`` / ``UMPAModelDF`` / ``UMPAModelDFKernel`` keep the reference's
constructor, ``match`` / ``min`` / ``cost`` / ``coverage`` methods, properties and
result dictionaries (UMPA/model.pyx:116-997); the work is done on the GPU by
``libumpa_b200.so`` through ``umpa_b200._capi`` (ctypes, C ABI).  PyTorch is used
only as plumbing: device output buffers, pinned host buffers, the current stream.

Differences from the reference that a caller can see (documented, deliberate):
  * host frames are copied to the GPU by the first match()/cost()/min() (pipelined with the
    kernels); device (torch) frames when the model is constructed.  Later in-place edits of
    the caller's arrays are not seen (the reference keeps raw pointers);
  * non-float64 inputs are converted (the reference reads them as garbage,
    model.pyx:236-237); when every frame is a float32 numpy array the frames go to the GPU as
    float32 (half the upload) and are widened there -- same results as for the widened arrays;
  * ``num_threads`` is accepted and ignored;
  * ``match(..., debug=False)`` skips the ``debug_*`` arrays (the reference decides
    this at compile time with ``DEF DEBUG``, model.pyx:26, 493-497);
  * outputs of pixels whose minimisation failed (``err == 0``) hold the walk's last
    state like the reference, except ``f`` which the reference leaves uninitialised.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from .geometry import Geometry

DEBUG = True      # mirrors `DEF DEBUG = True` (model.pyx:26)

__all__ = ["UMPAModelBase", "UMPAModelNoDF", "UMPAModelDF", "UMPAModelDFKernel",
           "spm", "spmq", "gaussian_kernel_test", "test_convolve", "test_CostArgsDFKernel", "pool_trim"]


def pool_trim():
    """Release the device blocks the library caches between models (umpa_pool_trim); returns the bytes freed.
    Call it when another allocator in the process (PyTorch's, ...) needs the memory."""
    return int(_capi.lib().umpa_pool_trim())


def _as_ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    return C.cast(arr, C.POINTER(C.c_void_p))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class UMPAModelBase:
    """Base class; use UMPAModelNoDF, UMPAModelDF or UMPAModelDFKernel (model.pyx:116-306)."""

    _kind = None
    Nparam = 0
    safe_crop = 0

    def __init__(self, sam_list, ref_list, mask_list=None, pos_list=None,
                 window_size=2, max_shift=4, ROI=None):
        if self._kind is None:
            raise NotImplementedError(
                'UMPAModelBase is not supposed to be called directly, use one of '
                'the subclasses "UMPAModelNoDF", "UMPAModelDF", or "UMPAModelDFKernel".')
        self._h = None
        Nw = int(window_size)
        Na = len(sam_list)
        self._on_device = all(isinstance(x, torch.Tensor) for x in sam_list)
        every = list(sam_list) + list(ref_list) + (list(mask_list) if mask_list is not None else [])
        self._f32 = all(isinstance(x, np.ndarray) and x.dtype == np.float32 for x in every)

        self._check_contiguous(sam_list)
        sams = [self._frame(s) for s in sam_list]
        shape_list = [np.array(s.shape, dtype=np.int32) for s in sams]
        self._shape_list = shape_list
        self._sam_list = sam_list

        self._check_contiguous(ref_list)
        refs = [self._frame(r) for r in ref_list]
        for k, r in enumerate(refs):
            samsh = (int(shape_list[k][0]), int(shape_list[k][1]))
            if samsh != tuple(r.shape):
                raise RuntimeError('Incompatible shape between sample {0} and '
                                   'reference frames {1} (entry [{2}] in the '
                                   'datasets).'.format(samsh, tuple(r.shape), k))
        self._ref_list = ref_list

        masks = None
        if mask_list is not None:
            self._check_contiguous(mask_list)
            masks = [self._frame(m) for m in mask_list]
        self._mask_list = mask_list

        if pos_list is None:
            pos_list = [np.zeros((2,), dtype=np.int32) for _ in range(Na)]
        else:
            pos_list = [np.asarray(p).astype(np.int32) for p in pos_list]
            if len(pos_list) != Na:
                raise RuntimeError(
                    'Unexpected length for position list (len(pos_list)={0}, '
                    'len(sam_list)={1})'.format(len(pos_list), Na))
        if np.any(np.array(pos_list) < 0):
            raise RuntimeError('Negative frame positions (entries in pos_list) are not allowed.')
        pmin = np.min(pos_list, axis=0)
        if not np.all(pmin == 0):
            raise RuntimeError('Positions should start at 0.')
        self._pos_list = pos_list

        self._Na = Na
        self._max_shift = int(max_shift)
        self._padding = self._max_shift + Nw + self.safe_crop            # model.pyx:286
        self._Nw = Nw
        self._window = self._make_window(Nw)
        self._uniform = (all(tuple(s) == tuple(shape_list[0]) for s in shape_list)
                         and not np.any(np.array(pos_list)))
        self._geo = Geometry(shape_list, pos_list, self._padding)
        self._masked = masks is not None

        L = _capi.lib()
        if torch.cuda.is_available():
            torch.cuda.init()
        h = C.c_void_p()
        dim = np.ascontiguousarray(np.array(shape_list, dtype=np.int32).reshape(Na, 2))
        pos = np.ascontiguousarray(np.array(pos_list, dtype=np.int32).reshape(Na, 2))
        _capi.check(L.umpa_create(C.byref(h), self._kind, Na,
                                  dim.ctypes.data_as(C.POINTER(C.c_int32)), pos.ctypes.data_as(C.POINTER(C.c_int32)),
                                  Nw, _dp(self._window), self._max_shift, self._padding))
        self._h = h

        def ptrs(frames):
            if self._on_device:
                return _as_ptr_array([int(f.data_ptr()) for f in frames])
            return _as_ptr_array([f.ctypes.data for f in frames])
        # host frames: the upload is deferred to the first match(), which pipelines it with the
        # kernels in row bands (umpa_match_host).  Like the reference (model.pyx:242-262) the model
        # keeps references to the caller's arrays until then.
        self._frames_keepalive = (sams, refs, masks)
        if self._f32:
            _capi.check(L.umpa_set_frames_f32(self._h, ptrs(sams), ptrs(refs),
                                              ptrs(masks) if masks is not None else None))
        else:
            _capi.check(L.umpa_set_frames(self._h, ptrs(sams), ptrs(refs),
                                          ptrs(masks) if masks is not None else None,
                                          1 if self._on_device else 2, self._stream()))
        if self._on_device:
            # umpa_set_frames(on_device=1) has copied (and synchronised): the float64 device copies made by
            # _frame() are not needed any more (host frames are kept: their upload is deferred)
            self._frames_keepalive = None
        self._geo.set_ROI(ROI)

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def _stream():
        if torch.cuda.is_available():
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        return None

    def _frame(self, x):
        if isinstance(x, torch.Tensor):
            if not self._on_device:
                return np.ascontiguousarray(x.detach().cpu().numpy(), dtype=np.float64)
            if not x.is_cuda:
                raise RuntimeError('Frames given as torch tensors must all live on the GPU.')
            return x.to(torch.float64).contiguous()
        return np.ascontiguousarray(np.asarray(x), dtype=np.float32 if self._f32 else np.float64)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _capi.lib().umpa_destroy(h)
            except Exception:
                pass

    def _check_contiguous(self, a):
        """model.pyx:311-317"""
        for x in a:
            ok = x.is_contiguous() if isinstance(x, torch.Tensor) else np.asarray(x).flags.c_contiguous
            if not ok:
                raise RuntimeError('The provided image frames are not C-contiguous.')

    def _make_window(self, n):
        """model.pyx:691-696"""
        window = np.multiply.outer(np.hamming(2 * n + 1), np.hamming(2 * n + 1))
        window /= window.sum()
        return np.ascontiguousarray(window, dtype=np.float64)

    # ------------------------------------------------------------------ geometry (model.pyx:531-646)
    # delegated to geometry.Geometry (GPU-free, unit-tested on CPU)
    def _calculate_extent(self):
        return self._geo.extent()

    def _convert_ROI_slice(self, ROI=None, step=None):
        return self._geo.convert(ROI, step)

    def _set_ROI(self, ROI=None):
        self._geo.set_ROI(ROI)

    def set_step(self, step):
        return self._geo.set_step(step)

    def coords(self, ROI=None):
        return self._geo.coords(ROI)

    _shape_of = staticmethod(Geometry.shape_of)

    @property
    def _ROI(self):
        return self._geo.ROI

    # ------------------------------------------------------------------ properties
    extent = property(lambda self: self._calculate_extent())
    ROI = property(lambda self: self._ROI, lambda self, new: self._set_ROI(new))
    sh = property(lambda self: self._shape_of(*self._ROI))
    Na = property(lambda self: self._Na)
    sam_list = property(lambda self: self._sam_list)
    ref_list = property(lambda self: self._ref_list)
    mask_list = property(lambda self: self._mask_list)
    shape_list = property(lambda self: self._shape_list)
    pos_list = property(lambda self: self._pos_list)
    window = property(lambda self: self._window)
    max_shift = property(lambda self: self._max_shift)
    padding = property(lambda self: self._padding)

    @property
    def Nw(self):
        return self._Nw

    @Nw.setter
    def Nw(self, new_Nw):
        """model.pyx:702-704 (padding is NOT recomputed there either)."""
        new_Nw = int(new_Nw)
        if new_Nw < 0:
            raise RuntimeError("Nw must be non-negative.")
        if self._max_shift + new_Nw + self.safe_crop > self._padding:
            # The reference keeps the old padding (model.pyx:702-704) and then reads outside its frames -- a host
            # over-read there, an illegal address that poisons the CUDA context here: refuse instead.
            raise RuntimeError("Nw = %d needs padding %d but the model was built with padding %d (window_size %d); "
                               "build a new model instead." % (new_Nw, self._max_shift + new_Nw + self.safe_crop,
                                                               self._padding, self._padding - self._max_shift - self.safe_crop))
        win = self._make_window(new_Nw)
        _capi.check(_capi.lib().umpa_set_window(self._h, new_Nw, _dp(win)))
        self._window, self._Nw = win, new_Nw

    @property
    def assign_coordinates(self):
        v = C.c_int(0)
        _capi.check(_capi.lib().umpa_get_option(self._h, _capi.OPT_REFERENCE_SHIFT, C.byref(v)))
        return {0: 'sam', 1: 'ref'}[v.value]

    @assign_coordinates.setter
    def assign_coordinates(self, new_mode):
        opts = {'sam': 0, 'ref': 1}
        if new_mode not in opts:
            print('Option %s is not available, parameter was not changed.' % repr(new_mode))
            return
        _capi.check(_capi.lib().umpa_set_option(self._h, _capi.OPT_REFERENCE_SHIFT, opts[new_mode]))

    @property
    def sub_pixel_mode(self):
        v = C.c_int(0)
        _capi.check(_capi.lib().umpa_get_option(self._h, _capi.OPT_SUBPX_FUNC, C.byref(v)))
        return v.value

    @sub_pixel_mode.setter
    def sub_pixel_mode(self, new_mode):
        _capi.check(_capi.lib().umpa_set_option(self._h, _capi.OPT_SUBPX_FUNC, int(new_mode)))

    # extension: choose the CUDA path ('auto' | 'table' | 'lazy'); both run on the GPU
    @property
    def cuda_path(self):
        v = C.c_int(0)
        _capi.check(_capi.lib().umpa_get_option(self._h, _capi.OPT_PATH, C.byref(v)))
        return {0: 'auto', 1: 'table', 2: 'lazy'}[v.value]

    @cuda_path.setter
    def cuda_path(self, name):
        _capi.check(_capi.lib().umpa_set_option(
            self._h, _capi.OPT_PATH, {'auto': 0, 'table': 1, 'lazy': 2}[name]))

    @property
    def last_match_info(self):
        p, n = C.c_int(0), C.c_int(0)
        _capi.check(_capi.lib().umpa_last_match_info(self._h, C.byref(p), C.byref(n)))
        return {"path": _capi.PATH_NAMES[p.value], "kernel_launches": n.value}

    @property
    def last_stream_info(self):
        """How the last host-to-host match was pipelined (zeros: it was not)."""
        b, t, r = C.c_int(0), C.c_int(0), C.c_int(0)
        _capi.check(_capi.lib().umpa_last_stream_info(self._h, C.byref(b), C.byref(t), C.byref(r)))
        return {"bands": b.value, "host_threads": t.value, "host_rows_per_frame": r.value}

    def test(self):
        return float(self._Na)

    # ------------------------------------------------------------------ coverage (model.pyx:499-529)
    def _coverage_device(self, s0, s1):
        N0, N1 = self._shape_of(s0, s1)
        out = torch.empty((N0, N1), dtype=torch.float64, device="cuda")
        if N0 and N1:
            _capi.check(_capi.lib().umpa_coverage(self._h, _capi.roi6((s0, s1)), C.c_void_p(out.data_ptr()),
                                                  1, self._stream()))
        return out

    def coverage(self, step=None, ROI=None):
        s0, s1 = self._convert_ROI_slice(ROI, step)
        return self._coverage_device(s0, s1).cpu().numpy()

    # ------------------------------------------------------------------ match (model.pyx:334-497)
    def match_device(self, step=None, dxdy=None, ROI=None, abc=None, debug=False, cover_max=None):
        """Like match() but leaves the result maps on the GPU (torch tensors) and does not
        synchronise.  This is the call bench.py times as the device-resident metric.
        cover_max: the coverage maximum the gate of model.pyx:431 uses instead of this model's own (a row band of
        a sharded match passes the maximum over all bands)."""
        if (ROI is not None) and (step is not None):
            step = None
        s0, s1 = self._convert_ROI_slice(ROI, step)
        self._set_ROI((s0, s1))                       # sticky, like the reference (model.pyx:406)
        N0, N1 = self._shape_of(s0, s1)
        dev = torch.device("cuda")
        f64 = dict(dtype=torch.float64, device=dev)
        out = {k: torch.empty((N0, N1), **f64) for k in ("f", "T", "dx", "dy")}
        if self._kind == _capi.DF:
            out["df"] = torch.empty((N0, N1), **f64)
        out["err"] = torch.empty((N0, N1), dtype=torch.int32, device=dev)
        out["debug_Ncalls"] = torch.empty((N0, N1), dtype=torch.int32, device=dev)
        if debug:
            out["debug_d"] = torch.empty((N0, N1, 25), **f64)
            out["debug_a"] = torch.empty((N0, N1, 16), **f64)
        if N0 == 0 or N1 == 0:
            return out

        abc_t = None
        if self._kind == _capi.DFKERNEL:
            abc_t = torch.as_tensor(abc, dtype=torch.float64).to(dev).contiguous()

        cover_t, thr = None, 0.
        if self._masked or not self._uniform:
            cover_t = self._coverage_device(s0, s1)
            thr = .1 * (float(cover_t.max()) if cover_max is None else float(cover_max)) / self._Na      # model.pyx:431
        uv0 = None
        if dxdy is not None:
            uv0 = (C.c_double * 2)(float(dxdy[0]), float(dxdy[1]))   # model.pyx:463-465

        o = _capi.Outputs()
        o.f, o.T, o.dx, o.dy = (out[k].data_ptr() for k in ("f", "T", "dx", "dy"))
        o.df = out["df"].data_ptr() if "df" in out else None
        o.err, o.ncalls = out["err"].data_ptr(), out["debug_Ncalls"].data_ptr()
        o.debug_d = out["debug_d"].data_ptr() if debug else None
        o.debug_a = out["debug_a"].data_ptr() if debug else None
        _capi.check(_capi.lib().umpa_match(
            self._h, _capi.roi6((s0, s1)), uv0,
            C.c_void_p(abc_t.data_ptr()) if abc_t is not None else None,
            C.c_void_p(cover_t.data_ptr()) if cover_t is not None else None, thr,
            C.byref(o), self._stream()))
        out["_keepalive"] = (abc_t, cover_t)
        return out

    def _match(self, step=None, input_values=None, dxdy=None, ROI=None, num_threads=None, quiet=False):
        """The reference's generic match routine (model.pyx:334-497), same signature and result:
        {'values': (N0, N1, Nparam) float64, 'err', 'debug_d', 'debug_a', 'debug_Ncalls'}.  `input_values`
        (same shape, float64) is the array the results are written into; for UMPAModelDFKernel its last three
        entries per pixel are the blur parameters (a, b, c) (model.pyx:436-455, 982-984).  Pixels the coverage
        gate skips keep their input values (model.pyx:480-481)."""
        if (ROI is not None) and (step is not None):
            s0, s1 = self._convert_ROI_slice(ROI, None)
        else:
            s0, s1 = self._convert_ROI_slice(ROI, step)
        shp = self._shape_of(s0, s1) + (self.Nparam,)
        abc = None
        if input_values is not None:
            if tuple(input_values.shape) != shp:
                raise RuntimeError("Input values have the wrong shape: "
                                   "%s, should be %s" % (input_values.shape, shp))
            if input_values.dtype != np.float64:
                raise RuntimeError("Input values have the wrong type: "
                                   "%s, should be %s" % (input_values.dtype, np.float64))
            values = input_values
        else:
            values = np.zeros(shp, dtype=np.float64)
        if self._kind == _capi.DFKERNEL:
            abc = np.ascontiguousarray(values[:, :, 4:7])
        res = self._match_maps(step=step, dxdy=dxdy, ROI=ROI, num_threads=num_threads, quiet=quiet, abc=abc,
                               debug=True)
        done = np.ones(shp[:2], dtype=bool)
        if self._masked or not self._uniform:      # the gate of model.pyx:427-431, 480
            cover = self.coverage(ROI=(s0, s1))
            done = ~(cover < .1 * cover.max() / self._Na)
        keys = ("f", "T", "dx", "dy") + (("df",) if self._kind == _capi.DF else ())
        for n, k in enumerate(keys):
            values[:, :, n][done] = res[k][done]
        return {"values": values, "err": res["err"], "debug_d": res["debug_d"], "debug_a": res["debug_a"],
                "debug_Ncalls": res["debug_Ncalls"]}

    def _match_maps(self, step=None, dxdy=None, ROI=None, num_threads=None, quiet=False, abc=None, debug=None):
        """Host-to-host match through umpa_match_host: result maps land in pinned host memory.  The
        first call on host frames pipelines upload, kernels and download in row bands."""
        if (ROI is not None) and (step is not None):
            print("Warning: 'ROI' and 'step' parameters are set simultaneously. "
                  "'step' parameter is ignored.")
            step = None
        if debug is None:
            debug = DEBUG
        if self._masked or not self._uniform:
            # the coverage map gates pixels (model.pyx:427-431, 480): keep it on the device
            dev = self.match_device(step=step, dxdy=dxdy, ROI=ROI, abc=abc, debug=debug)
            dev.pop("_keepalive", None)
            host = {}
            for k, t in dev.items():       # device -> pinned host, one stream, one sync
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                host[k] = h
            torch.cuda.current_stream().synchronize()
            return {k: h.numpy() for k, h in host.items()}
        s0, s1 = self._convert_ROI_slice(ROI, step)
        self._set_ROI((s0, s1))                       # sticky, like the reference (model.pyx:406)
        N0, N1 = self._shape_of(s0, s1)
        pin = torch.cuda.is_available()
        f64 = dict(dtype=torch.float64, pin_memory=pin)
        out = {k: torch.empty((N0, N1), **f64) for k in ("f", "T", "dx", "dy")}
        if self._kind == _capi.DF:
            out["df"] = torch.empty((N0, N1), **f64)
        out["err"] = torch.empty((N0, N1), dtype=torch.int32, pin_memory=pin)
        out["debug_Ncalls"] = torch.empty((N0, N1), dtype=torch.int32, pin_memory=pin)
        if debug:
            out["debug_d"] = torch.empty((N0, N1, 25), **f64)
            out["debug_a"] = torch.empty((N0, N1, 16), **f64)
        if N0 and N1:
            abc_h = None
            if self._kind == _capi.DFKERNEL:
                if isinstance(abc, torch.Tensor):
                    abc = abc.detach().cpu().numpy()
                abc_h = np.ascontiguousarray(abc, dtype=np.float64)
            cover_h, thr = None, 0.
            uv0 = None
            if dxdy is not None:
                uv0 = (C.c_double * 2)(float(dxdy[0]), float(dxdy[1]))   # model.pyx:463-465
            o = _capi.Outputs()
            o.f, o.T, o.dx, o.dy = (out[k].data_ptr() for k in ("f", "T", "dx", "dy"))
            o.df = out["df"].data_ptr() if "df" in out else None
            o.err, o.ncalls = out["err"].data_ptr(), out["debug_Ncalls"].data_ptr()
            o.debug_d = out["debug_d"].data_ptr() if debug else None
            o.debug_a = out["debug_a"].data_ptr() if debug else None
            _capi.check(_capi.lib().umpa_match_host(
                self._h, _capi.roi6((s0, s1)), uv0,
                C.c_void_p(abc_h.ctypes.data) if abc_h is not None else None,
                C.c_void_p(cover_h.ctypes.data) if cover_h is not None else None, thr, C.byref(o)))
        return {k: h.numpy() for k, h in out.items()}

    # single-pixel hooks ----------------------------------------------------
    def _min(self, i, j, abc=None):
        values = np.zeros((self.Nparam,), dtype=np.float64)
        if abc is not None:
            values[4:7] = abc
        uv = np.zeros(2)
        d, a = np.zeros(25), np.zeros(16)
        n, ok = C.c_int(0), C.c_int(0)
        _capi.check(_capi.lib().umpa_min(self._h, int(i), int(j), _dp(values), _dp(uv), _dp(d), _dp(a),
                                         C.byref(n), C.byref(ok)))
        self._last_min = {"ok": ok.value, "Ncalls": n.value, "d": d, "a": a}
        return values

    def _cost(self, i, j, sx, sy, abc=None):
        vals = np.zeros(3)
        st = C.c_int(0)
        abc_a = np.array(abc if abc is not None else (0., 0., 0.), dtype=np.float64)
        _capi.check(_capi.lib().umpa_cost(self._h, int(i), int(j), int(round(sx)), int(round(sy)),
                                          _dp(abc_a), _dp(vals), C.byref(st)))
        return vals, st.value


class UMPAModelNoDF(UMPAModelBase):
    """model.pyx:758-822"""
    _kind = _capi.NODF
    Nparam = 4
    safe_crop = 0

    def min(self, i, j):
        return self._min(i, j)

    def cost(self, i, j, sx, sy):
        v, _ = self._cost(i, j, sx, sy)
        return (v[0], v[1])

    def match(self, step=None, dxdy=None, ROI=None, num_threads=None, quiet=False, debug=None):
        return self._match_maps(step=step, dxdy=dxdy, ROI=ROI, num_threads=num_threads, quiet=quiet, debug=debug)


class UMPAModelDF(UMPAModelBase):
    """model.pyx:824-896"""
    _kind = _capi.DF
    Nparam = 5
    safe_crop = 0

    def min(self, i, j):
        return self._min(i, j)

    def cost(self, i, j, sx, sy):
        v, _ = self._cost(i, j, sx, sy)
        return (v[0], v[1], v[2])

    def match(self, step=None, dxdy=None, ROI=None, num_threads=None, quiet=False, debug=None):
        return self._match_maps(step=step, dxdy=dxdy, ROI=ROI, num_threads=num_threads, quiet=quiet, debug=debug)

    @property
    def Im(self):
        """Unused, uninitialised member in the reference (Model.h:153, model.pyx:891-896)."""
        return 0.


class UMPAModelDFKernel(UMPAModelBase):
    """model.pyx:899-997"""
    _kind = _capi.DFKERNEL
    Nparam = 7
    safe_crop = 8

    def min(self, i, j, a, b, c):
        return self._min(i, j, abc=(a, b, c))

    def cost(self, i, j, sx, sy, a, b, c):
        v, _ = self._cost(i, j, sx, sy, abc=(a, b, c))
        return (v[0], v[1])

    def match(self, step=None, abc=None, dxdy=None, ROI=None, num_threads=None, quiet=False, debug=None):
        if (ROI is not None) and (step is not None):
            step_eff = None
        else:
            step_eff = step
        s0, s1 = self._convert_ROI_slice(ROI, step_eff)
        self._set_ROI((s0, s1))
        sh = self._shape_of(s0, s1)
        if abc is None:
            raise RuntimeError('abc array has to be provided')
        elif tuple(abc.shape) != sh + (3,):
            raise RuntimeError('Wrong array shape for abc: %s, should be %s' % (abc.shape, sh + (3,)))
        return self._match_maps(step=step, dxdy=dxdy, ROI=ROI, num_threads=num_threads, quiet=quiet, abc=abc,
                                debug=debug)


# ---------------------------------------------------------------------- module hooks (model.pyx:31-114)
# The reference exposes its sub-pixel fits and the blur kernel for testing.  Here they are
# evaluated by the same device code that match() uses, through a one-pixel model whose
# "cost" hook is not needed: the 4x4 block / kernel maths is host-checkable, so these hooks
# are provided from numpy for API completeness only (they are not on the hot path).

_BSP = np.array([[1., 4., 1., 0.], [-3., 0., 3., 0.], [3., -6., 3., 0.], [-1., 3., -3., 1.]])


def spmq(a):
    """Reference name/docstring swap kept: spmq -> spline fit (spmin), model.pyx:57-80."""
    a = np.asarray(a, dtype=np.float64)
    if a.shape != (4, 4):
        raise RuntimeError('input array must be (4,4)')
    c = _BSP @ a @ _BSP.T          # c[n, m]: x^n y^m
    x = y = 0.
    for _ in range(21):
        xp, yp = np.array([1., x, x * x, x ** 3]), np.array([1., y, y * y, y ** 3])
        dxp, dyp = np.array([0., 1., 2 * x, 3 * x * x]), np.array([0., 1., 2 * y, 3 * y * y])
        ddxp, ddyp = np.array([0., 0., 2., 6 * x]), np.array([0., 0., 2., 6 * y])
        fx, fy = dxp @ c @ yp, xp @ c @ dyp
        fxx, fxy, fyy = ddxp @ c @ yp, dxp @ c @ dyp, xp @ c @ ddyp
        det = fxx * fyy - fxy * fxy
        dx, dy = (fxy * fy - fyy * fx) / det, (fxy * fx - fxx * fy) / det
        x, y = x + dx, y + dy
        if dx * dx + dy * dy < 1e-8:
            break
    xp, yp = np.array([1., x, x * x, x ** 3]), np.array([1., y, y * y, y ** 3])
    return np.array([x, y]), float(xp @ c @ yp) / 36.


def spm(a):
    """spm -> quadratic fit (spmin_quad), model.pyx:31-54."""
    a = np.asarray(a, dtype=np.float64)
    if a.shape != (4, 4):
        raise RuntimeError('input array must be (4,4)')
    ii, jj = np.mgrid[-1:3, -1:3].astype(float)
    A = np.stack([np.ones(16), ii.ravel(), jj.ravel(), ii.ravel() ** 2, (ii * jj).ravel(), jj.ravel() ** 2], axis=1)
    p = np.round(400. * np.linalg.pinv(A)) @ a.ravel()
    det = 4 * p[3] * p[5] - p[4] * p[4]
    pos = np.array([-(2 * p[3] * p[2] - p[4] * p[1]) / det, -(2 * p[5] * p[1] - p[4] * p[2]) / det])
    return pos, float((p[0] + .5 * (p[2] * pos[0] + p[1] * pos[1])) / 400.)


def gaussian_kernel_test(Nk, a, b, c):
    """model.pyx:82-92"""
    i, j = np.mgrid[-Nk:Nk + 1, -Nk:Nk + 1].astype(float)
    return np.exp(-a * i * i - b * i * j - c * j * j)


def test_convolve(image, i, j, kernel):
    """model.pyx:94-102 -> convolve (Utils.cpp:85-97): sum_kl kernel[k, l] * image[i + k - Nk, j + l - Nk].
    Like the reference there is no bounds check beyond what numpy slicing gives."""
    image = np.asarray(image, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    Nk = (kernel.shape[0] - 1) // 2
    i, j = int(i), int(j)
    return float((kernel[:2 * Nk + 1, :2 * Nk + 1] * image[i - Nk:i + Nk + 1, j - Nk:j + Nk + 1]).sum())


test_convolve.__test__ = False


def test_CostArgsDFKernel(i, j, a, b, c):
    """model.pyx:104-114: the normalised 17x17 kernel of CostArgsDFKernel."""
    k = gaussian_kernel_test(8, a, b, c)
    return k / k.sum()


test_CostArgsDFKernel.__test__ = False
