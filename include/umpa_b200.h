/*
 * umpa_b200.h -- C ABI of libumpa_b200.so, the B200 (sm_100a) implementation of
 * UMPA++'s per-pixel window-matching path.
 *
 * This is the drop-in boundary.  In the reference the seam is the C++ class
 * models::ModelBase<double> and its subclasses, driven from Cython
 * (UMPA/Model.pxd:31-69, UMPA/model.pyx:116-997).  Every entry point below names
 * the reference interface it replaces (paths relative to the reference root).
 * Plain pointers and sizes only; no C++/torch types.  All functions return
 * UMPA_OK (0) or a negative error code; umpa_last_error() gives the message of
 * the calling thread's last failure.  Nothing throws across the boundary.
 *
 * Conventions
 *  - frames are row-major (rows, cols); "i"/"0" is the row axis, "j"/"1" the column axis;
 *  - pixel coordinates (i, j) are RAW frame coordinates (padding included), exactly
 *    what ModelBase::min / cost_interface take (UMPA/model.pyx:482-486, 780-789);
 *  - a shift (si, sj) is valid when |si|,|sj| <= max_shift-1 (UMPA/lib/Model.cpp:372-399);
 *  - the model works on the CUDA device that is current when umpa_create() is called;
 *  - "stream" arguments are a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef UMPA_B200_H
#define UMPA_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define UMPA_API __attribute__((visibility("default")))
#else
#define UMPA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct umpa_model umpa_model;

/* model kinds: ModelNoDF / ModelDF / ModelDFKernel (UMPA/lib/Model.h:126-191) */
enum { UMPA_NODF = 0, UMPA_DF = 1, UMPA_DFKERNEL = 2 };

/* return codes */
enum {
    UMPA_OK = 0,
    UMPA_ERR_ARG = -1,        /* bad argument (message says which) */
    UMPA_ERR_CUDA = -2,       /* a CUDA call or kernel failed */
    UMPA_ERR_STATE = -3,      /* call order (e.g. match before frames were set) */
    UMPA_ERR_UNSUPPORTED = -4 /* request outside what the selected path implements */
};

/* options for umpa_set_option() */
enum {
    UMPA_OPT_SUBPX_FUNC = 1,      /* -1 spline (default) | 0 none | 1 quadratic; ModelBase::subpx_func, Model.cpp:214-221 */
    UMPA_OPT_REFERENCE_SHIFT = 2, /* 0 (default) | 1; ModelBase::reference_shift, Model.cpp:215, 408-421 */
    UMPA_OPT_PATH = 3             /* UMPA_PATH_*: which CUDA path match() uses */
};

/* UMPA_OPT_PATH values.  All are CUDA paths; there is no CPU path.
 *  TABLE: FP32 exhaustive shift tables + FP64 per-pixel solve and walk.  NoDF / DF: cross-correlation summed
 *         over frames, then Hamming-filtered (max_shift <= 10, Nw <= 6, either assign_coordinates); DFKernel: the
 *         per-pixel blur fused into the window pass (Nw <= 3, max_shift <= 6).  Needs a separable window.
 *  LAZY:  one thread per pixel evaluates the reference's cost() on demand in FP64, in the reference's
 *         summation order (all models, masks, positions, any window / max_shift).
 *  AUTO:  TABLE when the model has equal frames at position 0 and no masks; with masks or ragged frames /
 *         positions the MIXED path: TABLE on every pixel for which it is exact (no mask value != 1 within reach;
 *         every frame either contains the pixel's reach or misses it), LAZY on the rest; LAZY when TABLE does
 *         not apply at all.  Requesting TABLE on a model that is not eligible is an error.
 *  MASKED (reported only): the MIXED path of a model whose masks are 0 / 1 and the same in every frame (a dead-pixel
 *         map): the pixels with a dead pixel within reach are matched from the same tables, their sums corrected
 *         by the few window positions the mask removes (Model.cpp:461-499, 775-847), instead of going to LAZY (NoDF /
 *         DF, either assign_coordinates).
 *         UMPA_MASK_TABLES=0 in the environment turns it off. */
enum { UMPA_PATH_AUTO = 0, UMPA_PATH_TABLE = 1, UMPA_PATH_LAZY = 2,
       UMPA_PATH_MIXED = 3 /* reported by umpa_last_match_info only (see AUTO above) */,
       UMPA_PATH_MASKED = 4 /* reported only */ };

/* Output maps of umpa_match*, all row-major (N0, N1); any pointer may be NULL to
 * skip that map.  Replaces the `values`, `err`, `debug_*` arrays that
 * UMPAModelBase._match allocates (UMPA/model.pyx:442-474) and the per-key copies
 * of UMPAModel*.match (model.pyx:815-822, 881-889, 990-997). */
typedef struct umpa_outputs {
    double *f;         /* values[...,0]  cost at the (sub-pixel) minimum      */
    double *T;         /* values[...,1]  transmission                          */
    double *dx;        /* values[...,2]  = uv[1], column shift                 */
    double *dy;        /* values[...,3]  = uv[0], row shift                    */
    double *df;        /* values[...,4]  dark field (DF only)                  */
    int32_t *err;      /* error_status.ok per pixel (1 = ok)                   */
    int32_t *ncalls;   /* debug_Ncalls                                         */
    double *debug_d;   /* (N0, N1, 25) minimizer_debug.d  (UMPA/lib/Optim.h:15-21) */
    double *debug_a;   /* (N0, N1, 16) minimizer_debug.a                       */
} umpa_outputs;

/* ---- lifetime -----------------------------------------------------------
 * umpa_create replaces `new Model{NoDF,DF,DFKernel}<double>(Na, dim, sams, refs,
 * masks, pos, Nw, win, max_shift, padding)` (UMPA/lib/Model.cpp:193-216, 597-604,
 * 963-968; called from UMPA/model.pyx:292, 769, 835, 911).
 *   dim, pos : Na x 2 int32 (rows, cols) per frame   (model.pyx:226-239, 265-283)
 *   win      : (2Nw+1)^2 doubles, row-major          (model.pyx:691-696)
 * Frames are supplied afterwards with umpa_set_frames(). */
UMPA_API int umpa_create(umpa_model **out, int kind, int Na, const int32_t *dim, const int32_t *pos,
                int Nw, const double *win, int max_shift, int padding);

/* replaces `del self.c_model` (UMPA/model.pyx:308-309) */
UMPA_API void umpa_destroy(umpa_model *m);

/* ---- inputs --------------------------------------------------------------
 * The reference keeps raw double* into the caller's numpy arrays
 * (vector<T*> sam/ref/mask, UMPA/model.pyx:235-262).  Here the frames are staged
 * into device memory owned by the handle: FP64 copies (LAZY path, cost/min hooks)
 * and, when a TABLE path is eligible, centred FP32 stacks.
 *   sam, ref : Na pointers to row-major float64 frames of shape dim[k]
 *   mask     : Na pointers or NULL (no masks)
 *   on_device: 0 = host pointers (pageable or pinned), copied before the call returns;
 *              1 = device pointers, copied before the call returns;
 *              2 = host pointers, copy DEFERRED: like the reference, the handle keeps the
 *                  pointers and the buffers must stay valid and unchanged for its lifetime.
 *                  The first umpa_match_host() then pipelines upload, kernels and download
 *                  in row bands (host threads convert part of the rows to centred FP32 on the
 *                  way, see UMPA_HOST_THREADS in INTEGRATION.md); any other entry point
 *                  uploads whatever it needs first. */
UMPA_API int umpa_set_frames(umpa_model *m, const double *const *sam, const double *const *ref,
                    const double *const *mask, int on_device, void *stream);

/* The same for float32 HOST frames (detector data): the reference only reads float64 buffers
 * (UMPA/model.pyx:236-237 casts whatever it is given to double*), so its callers widen such data first;
 * here they cross PCIe as they are -- half the bytes -- and are widened on the GPU (exact), which gives the
 * results of umpa_set_frames() on the widened frames bit for bit.  Always deferred (like on_device = 2):
 * the pointers must stay valid until the first umpa_match_host / umpa_match / umpa_cost / umpa_min. */
UMPA_API int umpa_set_frames_f32(umpa_model *m, const float *const *sam, const float *const *ref,
                        const float *const *mask);

/* replaces ModelBase::set_window (UMPA/lib/Model.cpp:239-246; Nw setter model.pyx:702-704) */
UMPA_API int umpa_set_window(umpa_model *m, int Nw, const double *win);

/* replaces direct writes to c_model.subpx_func / reference_shift (model.pyx:742, 755) */
UMPA_API int umpa_set_option(umpa_model *m, int option, int value);
UMPA_API int umpa_get_option(const umpa_model *m, int option, int *value);

/* ---- the hot path --------------------------------------------------------
 * umpa_match replaces the OpenMP pixel loop of UMPAModelBase._match
 * (UMPA/model.pyx:476-492): for xi < N0, xj < N1 it runs Model*::min at raw pixel
 * (padding + start0 + step0*xi, padding + start1 + step1*xj).
 *   roi    : {start0, stop0, step0, start1, stop1, step1}   (model.pyx:409-415)
 *   uv0    : start guess (row shift, col shift) applied to every pixel, or NULL
 *            for (0,0)                                       (model.pyx:461-465)
 *   abc    : (N0, N1, 3) float64 blur parameters, DFKernel only (model.pyx:973-984)
 *   cover  : (N0, N1) float64 coverage map + threshold gate, or NULL = no gate
 *            (model.pyx:427-431, 480); skipped pixels keep zeros
 *   out    : DEVICE pointers (umpa_match) / HOST pointers (umpa_match_host)
 *   abc/cover are DEVICE pointers for umpa_match, HOST pointers for umpa_match_host.
 * umpa_match is asynchronous on `stream`; umpa_match_host returns after the
 * outputs are in host memory. */
UMPA_API int umpa_match(umpa_model *m, const int32_t roi[6], const double uv0[2], const double *abc,
               const double *cover, double cover_threshold, const umpa_outputs *out, void *stream);
UMPA_API int umpa_match_host(umpa_model *m, const int32_t roi[6], const double uv0[2], const double *abc,
                    const double *cover, double cover_threshold, const umpa_outputs *out);

/* ---- single-pixel entry points (debug / tests) ---------------------------
 * umpa_cost replaces Model*::cost_interface (UMPA/lib/Model.cpp:533-542, 887-897,
 * 1181-1192; model.pyx:780-789, 846-856, 925-940): values = {cost, t, v};
 * *status gets the error_status bits (1 ok, 2 bound_error, 4 dimension, 8 positive).
 * umpa_min replaces Model*::min for one pixel (Model.cpp:562-578, 923-940, 1222-1238;
 * model.pyx:325-332, 772-778): values has Nparam entries {f, T, dx, dy, [df | a, b, c]}
 * (a, b, c read on input for DFKernel), uv is in/out.  Both always use the LAZY path
 * and synchronise. */
UMPA_API int umpa_cost(umpa_model *m, int i, int j, int si, int sj, const double abc[3],
              double values[3], int *status);
UMPA_API int umpa_min(umpa_model *m, int i, int j, double *values, double uv[2],
             double dbg_d[25], double dbg_a[16], int *ncalls, int *ok);

/* replaces the double loop over ModelBase::coverage (UMPA/model.pyx:499-529,
 * UMPA/lib/Model.cpp:273-314).  out: (N0, N1) float64, host (on_device=0) or device. */
UMPA_API int umpa_coverage(umpa_model *m, const int32_t roi[6], double *out, int on_device, void *stream);

/* ---- post-processing of the displacement maps -------------------------------
 * umpa_correct_bad_pixels replaces correct_bad_pixels(img, th, iterations, dims=(-2,-1))
 * (UMPA/align.py:661-732) as UMPA_normal / UMPA_nobias apply it to dx and dy right after
 * match() (align.py:58-60, 111-114), with the bias subtraction of UMPA_nobias fused in:
 *   out = correct(img - bias),  values outside [lo, hi] replaced by the median of their four
 *   neighbours (reflected at the edges exactly as the reference indexes them).
 * All pointers are DEVICE pointers to (nimg, N0, N1) float64; bias may be NULL; scratch (same
 * size) is needed when bias is given, when img == out, or for iterations > 1.  Asynchronous on
 * `stream`.  (NaNs: fmin/fmax skip them, numpy's median propagates them.) */
UMPA_API int umpa_correct_bad_pixels(const double *img, const double *bias, double *out, double *scratch,
                            int64_t nimg, int N0, int N1, double lo, double hi, int iterations, void *stream);

/* ---- introspection --------------------------------------------------------- */
/* which path the last umpa_match used (UMPA_PATH_TABLE / UMPA_PATH_LAZY) and how
 * many kernels it launched */
UMPA_API int umpa_last_match_info(const umpa_model *m, int *path, int *kernel_launches);
/* how the last pipelined umpa_match_host ran: row bands, host conversion threads, rows per
 * frame that crossed PCIe as host-converted FP32 (all 0 when the match was not pipelined) */
UMPA_API int umpa_last_stream_info(const umpa_model *m, int *bands, int *host_threads, int *host_rows);
/* device time of the stages of the last TABLE-path match, measured with CUDA events on the
 * launching stream when profiling is enabled via umpa_set_profiling(m, 1):
 * ms[0] moments, ms[1] cross table, ms[2] mean table, ms[3] walk; returns count written */
UMPA_API int umpa_set_profiling(umpa_model *m, int enable);
UMPA_API int umpa_last_stage_ms(umpa_model *m, float *ms, int n);
/* bytes of device memory currently owned by the handle */
UMPA_API int64_t umpa_device_bytes(const umpa_model *m);
/* Freed frame / scratch blocks are cached for the next model (callers build one model per projection; the
 * reference's constructor is free, UMPA/model.pyx:138-296).  The cache is capped at UMPA_POOL_GB (default: a
 * quarter of the device memory) and emptied when the library's own cudaMalloc fails; umpa_pool_trim() empties
 * it on demand (e.g. before another library needs the memory) and returns the bytes released. */
UMPA_API int64_t umpa_pool_trim(void);

/* Host helpers of the pipelined upload (hoststage.cu), exposed for CPU-side tests: the centring constant of
 * a frame (mean over rows 0, step, 2 step, ...) and the FP64 -> centred FP32 conversion of `rows` rows of W
 * doubles into rows of `pitch` floats (zero padded) -- the arithmetic of center_frames: (float)(x - c). */
UMPA_API double umpa_host_sampled_mean(const double *frame, int H, int W, int step);
UMPA_API double umpa_host_sampled_mean_f32(const float *frame, int H, int W, int step);
UMPA_API void umpa_host_center_rows(float *dst, const double *src, int rows, int W, int pitch, double c);
UMPA_API void umpa_host_center_rows_f32(float *dst, const float *src, int rows, int W, int pitch, double c);

/* Geometry the table path chooses for the cross table of a match over rows x cols pixels (host arithmetic, no
 * device needed): out = chunk rows EH, order (0 halo tiles, 1 streaming chunk-major, 2 streaming pass-major),
 * warp groups G, passes over the frames, frames per TMA box, ring stages, output columns per strip TW, row
 * segments, column strips, threads per CTA.  For bench.py's executed-work model and the tests. */
UMPA_API int umpa_table_plan(int Na, int Nw, int max_shift, int rows, int cols, int sm_count, int out[10]);

/* Measured FP32-FMA peak of the current device (dependent-free FFMA chains on all SMs, CUDA
 * events): the denominator of the FP32-FMA roofline that bench.py reports. */
UMPA_API int umpa_fma_peak(double *tflops, int *sm_count);

UMPA_API const char *umpa_last_error(void);
UMPA_API const char *umpa_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UMPA_B200_H */
