"""ctypes binding of libumpa_b200.so (include/umpa_b200.h).  No fallback: if the CUDA
library is missing or a call fails, this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UMPA_LIB") or os.path.join(HERE, "csrc", "libumpa_b200.so")   # UMPA_LIB: tuning builds

NODF, DF, DFKERNEL = 0, 1, 2
OPT_SUBPX_FUNC, OPT_REFERENCE_SHIFT, OPT_PATH = 1, 2, 3
PATH_AUTO, PATH_TABLE, PATH_LAZY = 0, 1, 2
PATH_NAMES = {0: "none", 1: "table", 2: "lazy", 3: "mixed", 4: "masked_table"}

EXPORTS = ("umpa_create", "umpa_destroy", "umpa_set_frames", "umpa_set_frames_f32", "umpa_set_window", "umpa_set_option",
           "umpa_get_option", "umpa_match", "umpa_match_host", "umpa_cost", "umpa_min", "umpa_coverage", "umpa_correct_bad_pixels",
           "umpa_last_match_info", "umpa_last_stream_info", "umpa_set_profiling", "umpa_last_stage_ms", "umpa_device_bytes", "umpa_pool_trim", "umpa_table_plan",
           "umpa_fma_peak", "umpa_host_sampled_mean", "umpa_host_sampled_mean_f32", "umpa_host_center_rows", "umpa_host_center_rows_f32", "umpa_last_error", "umpa_version")


class Outputs(C.Structure):
    """struct umpa_outputs"""
    _fields_ = [("f", C.c_void_p), ("T", C.c_void_p), ("dx", C.c_void_p), ("dy", C.c_void_p),
                ("df", C.c_void_p), ("err", C.c_void_p), ("ncalls", C.c_void_p),
                ("debug_d", C.c_void_p), ("debug_a", C.c_void_p)]


class UmpaError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UmpaError(
            "libumpa_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python umpa_b200/build.py`; there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.umpa_create.restype = C.c_int
    L.umpa_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, ip, ip, C.c_int, dp, C.c_int, C.c_int]
    L.umpa_destroy.restype = None
    L.umpa_destroy.argtypes = [vp]
    L.umpa_set_frames.restype = C.c_int
    L.umpa_set_frames.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_int, vp]
    L.umpa_set_frames_f32.restype = C.c_int
    L.umpa_set_frames_f32.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.umpa_set_window.restype = C.c_int
    L.umpa_set_window.argtypes = [vp, C.c_int, dp]
    L.umpa_set_option.restype = C.c_int
    L.umpa_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.umpa_get_option.restype = C.c_int
    L.umpa_get_option.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.umpa_match.restype = C.c_int
    L.umpa_match.argtypes = [vp, ip, dp, vp, vp, C.c_double, C.POINTER(Outputs), vp]
    L.umpa_match_host.restype = C.c_int
    L.umpa_match_host.argtypes = [vp, ip, dp, vp, vp, C.c_double, C.POINTER(Outputs)]
    L.umpa_cost.restype = C.c_int
    L.umpa_cost.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_int)]
    L.umpa_min.restype = C.c_int
    L.umpa_min.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.umpa_coverage.restype = C.c_int
    L.umpa_coverage.argtypes = [vp, ip, vp, C.c_int, vp]
    L.umpa_correct_bad_pixels.restype = C.c_int
    L.umpa_correct_bad_pixels.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, vp]
    L.umpa_last_match_info.restype = C.c_int
    L.umpa_last_match_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.umpa_last_stream_info.restype = C.c_int
    L.umpa_last_stream_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.umpa_set_profiling.restype = C.c_int
    L.umpa_set_profiling.argtypes = [vp, C.c_int]
    L.umpa_last_stage_ms.restype = C.c_int
    L.umpa_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    L.umpa_device_bytes.restype = C.c_int64
    L.umpa_device_bytes.argtypes = [vp]
    L.umpa_pool_trim.restype = C.c_int64
    L.umpa_pool_trim.argtypes = []
    L.umpa_table_plan.restype = C.c_int
    L.umpa_table_plan.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)]
    L.umpa_fma_peak.restype = C.c_int
    L.umpa_fma_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.umpa_host_sampled_mean.restype = C.c_double
    L.umpa_host_sampled_mean.argtypes = [dp, C.c_int, C.c_int, C.c_int]
    L.umpa_host_sampled_mean_f32.restype = C.c_double
    L.umpa_host_sampled_mean_f32.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int]
    L.umpa_host_center_rows.restype = None
    L.umpa_host_center_rows.argtypes = [C.POINTER(C.c_float), dp, C.c_int, C.c_int, C.c_int, C.c_double]
    L.umpa_host_center_rows_f32.restype = None
    L.umpa_host_center_rows_f32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_double]
    L.umpa_last_error.restype = C.c_char_p
    L.umpa_version.restype = C.c_char_p
    _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise UmpaError(lib().umpa_last_error().decode("utf-8", "replace") or ("umpa error %d" % rc))


def roi6(ROI):
    (s0, e0, t0), (s1, e1, t1) = ROI
    return (C.c_int32 * 6)(int(s0), int(e0), int(t0), int(s1), int(e1), int(t1))
