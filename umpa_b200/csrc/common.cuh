// common.cuh -- shared declarations of libumpa_b200 (host + device).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/umpa_b200.h"

#define UMPA_KWS 8                         // KERNEL_WINDOW_SIZE, UMPA/lib/Model.h:7
#define UMPA_KSIDE (2 * UMPA_KWS + 1)
#define UMPA_MAX_CALLS 500                 // UMPA/lib/Optim.cpp:14
#define UMPA_MAX_K 31                      // largest supported window side (Nw <= 15)

// error_status bits (UMPA/lib/Optim.h:7-12)
#define UMPA_ST_OK 1
#define UMPA_ST_BOUND 2
#define UMPA_ST_DIM 4
#define UMPA_ST_POS 8

void umpa_set_error(const char *fmt, ...);

#define UMPA_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            umpa_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return UMPA_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

// ---------------------------------------------------------------- device views

// What the LAZY path needs to evaluate the reference's cost() for any pixel.
struct LazyView {
    int kind, Na, Nw, max_shift, padding, subpx, refshift, masked;
    const double *const *sam;    // device arrays of device pointers
    const double *const *ref;
    const double *const *mask;
    const int *dim;              // Na x 2
    const int *pos;              // Na x 2
    const double *win;           // K*K
    const double *quad;          // 6x16 least-squares matrix of the quadratic sub-pixel fit
};

// Geometry of one match() call.
struct RoiView {
    int off0, step0, N0;         // raw row of output row xi = off0 + step0*xi   (off = padding + start)
    int off1, step1, N1;
    double uv0[2];               // start guess (row, col)
    const double *abc;           // (N0,N1,3) or nullptr
    const double *cover;         // (N0,N1) or nullptr
    double cover_threshold;
    // masked models on the mixed path: dirty[n] != 0 where a mask value != 1 is within reach of the pixel;
    // a kernel handles the pixel iff dirty == nullptr or (dirty[n] != 0) == dirty_want
    const unsigned char *dirty;
    int dirty_want;
};

// ---------------------------------------------------------------- the handle

struct Scratch {
    void *p = nullptr;
    size_t bytes = 0;
};

struct umpa_model {
    int kind = 0, Na = 0, Nw = 0, K = 0, max_shift = 0, padding = 0;
    int subpx = -1, refshift = 0, path_opt = UMPA_PATH_AUTO;
    int device = 0, sm_count = 148;
    std::vector<int> dim, pos;                   // host copies
    std::vector<double> win;                     // K*K
    bool uniform = false;                        // equal shapes and zero positions
    int H = 0, W = 0;                            // common shape when uniform, else the canvas circumscribing all frames
    bool separable = false;                      // win == g (x) g
    std::vector<double> g;                       // 1-D factor, K
    double win_sum = 0.;
    bool masked = false, frames_set = false;

    // device: FP64 frames (LAZY path) -- one allocation per stack
    double *d_sam64 = nullptr, *d_ref64 = nullptr, *d_mask64 = nullptr;
    std::vector<size_t> frame_off;               // element offset of frame k inside the stack
    size_t stack_elems = 0;
    const double **d_sam_ptrs = nullptr, **d_ref_ptrs = nullptr, **d_mask_ptrs = nullptr;
    int *d_dim = nullptr, *d_pos = nullptr;
    double *d_win = nullptr, *d_quad = nullptr;

    // device: centred FP32 stacks (TABLE path); pitch in floats, multiple of 4
    float *d_sam32 = nullptr, *d_ref32 = nullptr;
    int pitch = 0;
    float *d_mean_s = nullptr, *d_mean_r = nullptr, *d_g = nullptr;   // centring constants d_k, c_k (FP32 copies)
    double *d_means64 = nullptr;                 // [2*Na]: sample then reference centring constants (FP64)
    double *d_consts = nullptr;                  // [3]: sum_k c_k d_k, sum_k c_k^2, sum_k d_k^2
    double *d_partials = nullptr;

    // host frames whose upload is deferred to the first match (umpa_set_frames with on_device = 2):
    // umpa_match_host then pipelines upload, kernels and download in row bands
    std::vector<const double *> h_sam, h_ref, h_mask;
    std::vector<const float *> h_sam_f, h_ref_f, h_mask_f;    // umpa_set_frames_f32: float32 host frames instead
    bool host_f32 = false;
    bool host_pending = false;                   // nothing of the host frames is on the device yet
    bool fp64_missing = false;                   // FP32 stacks complete, FP64 stacks not (rows the host converted)
    cudaStream_t s_copy = nullptr, s_comp = nullptr, s_out = nullptr;
    Scratch outbuf;
    int stream_bands = 0, stream_threads = 0, stream_host_rows = 0;   // how the last pipelined match ran
    void *h_small = nullptr;                     // pinned: constants on their way to the device

    // TABLE-path scratch (grow-only)
    Scratch filtA, filtB, auxS, auxR, tabX;
    Scratch maskbad, dirty;                      // masked models: row-dilated "mask != 1" image [H][W], per-ROI dirty map
    bool maskbad_valid = false;
    // masks that are 0 / 1 and the same in every frame take the corrected table walk (table_path.cu: masked_walk_kernel)
    Scratch maskbits, maskflags;                 // dead-pixel bit image [H][W/32 + 2]; classification flags (device)
    Scratch maskwin;                             // Nw <= 3: the dead pixels of every pixel's window, one 64-bit word per pixel
    int maskwin_Nw = -1;
    Scratch fmImgS, fmImgR;                      // per-pixel frame sums of the centred stacks (float4 images)
    Scratch fmS, fmR, fmA;                       // frame-minor copies [H][pitch][4 ceil(Na/4)] of the FP32 stacks / filtA
    int mask_mode = 0;                           // 0 not classified yet | 1 general (lazy evaluation) | 2 binary and shared
    bool moments_valid = false;

    // One device block for all the small per-model arrays (shapes, window, pointer tables, constants):
    // carved once in umpa_create, recycled through the block cache -- callers create a model per
    // projection, and a dozen cudaMalloc/cudaFree pairs per model cost more than a 256^2 match.
    char *arena = nullptr;
    size_t arena_bytes = 0;
    bool win_own = false;                        // d_win / d_g outside the arena (window side > UMPA_MAX_K)
    bool h_small_own = false;                    // h_small outside the pinned chunk pool

    // bookkeeping
    int last_path = 0, last_launches = 0;
    bool profiling = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    int64_t dev_bytes = 0;
    // recorded on the stream of the model's last launch; the next call waits for it (capi.cu: order_after_last)
    cudaEvent_t last_ev = nullptr;
    bool last_ev_set = false;
};

int scratch_reserve(umpa_model *m, Scratch &s, size_t bytes);
cudaError_t pool_malloc(void **p, size_t bytes);   // cached cudaMalloc / cudaFree
void *pinned_small_take(size_t bytes, bool *own);  // pinned host memory for constants on their way to the device
void pinned_small_give(void *p, bool own);
void pool_free(void *p, size_t bytes);
size_t pool_trim();                                // frees every cached block; returns the bytes released

// implemented in lazy_path.cu
int lazy_match(umpa_model *m, const RoiView &roi, const umpa_outputs &out, cudaStream_t st);
int lazy_cost(umpa_model *m, int i, int j, int si, int sj, const double abc[3], double values[3], int *status);
int lazy_min(umpa_model *m, int i, int j, double *values, double uv[2], double *dd, double *da, int *ncalls, int *ok);
int coverage_map(umpa_model *m, const RoiView &roi, double *out_dev, cudaStream_t st);

// implemented in table_path.cu
int table_prepare_frames(umpa_model *m, cudaStream_t st);      // FP64 stacks -> centred FP32 stacks (all three steps)
int table_alloc32(umpa_model *m);                              // 1. FP32 stacks + constants (no-op when not applicable)
int table_means(umpa_model *m, cudaStream_t st);               // 2. centring constants from the sampled rows (see table_row_step)
int table_center_rows_inplace(umpa_model *m, int y0, int y1, cudaStream_t st);   // 3'. raw FP32 rows already in the FP32 stacks
int table_center_rows(umpa_model *m, int y0, int y1, cudaStream_t st);   // 3. rows [y0,y1) of every frame -> centred FP32
int table_set_means(umpa_model *m, const double *mu, cudaStream_t st);   // 2'. constants computed by the host (mu: 2*Na)
int table_row_step(int H);                                     // rows y = 0, step, 2 step, ... define the centring constants
bool table_eligible(const umpa_model *m, const RoiView &roi, std::string *why, bool mixed = false);   // mixed: masks / ragged frames allowed
int mixed_match(umpa_model *m, const RoiView &roi, const umpa_outputs &out, cudaStream_t st);   // masked or ragged NoDF/DF
int table_match(umpa_model *m, const RoiView &roi, const umpa_outputs &out, cudaStream_t st);

// implemented in hoststage.cu (host code)
double host_sampled_mean(const double *frame, int H, int W, int step);
double host_sampled_mean_f32(const float *frame, int H, int W, int step);     // == host_sampled_mean of the widened frame
void host_center_rows(float *dst, const double *src, int rows, int W, int pitch, double c);
void host_center_rows_f32(float *dst, const float *src, int rows, int W, int pitch, double c);

// implemented in kernel_path.cu: per-pixel FP32 tables of UMPAModelDFKernel (blur fused into the window pass)
bool ktable_supported(int Nw, int max_shift, int step0, bool refshift);
int ktable_row_floats(int max_shift);                          // floats per pixel row: t5c[S^2], t3c[S^2], sigma-1
int ktable_build(umpa_model *m, const RoiView &roi, float *tab, cudaStream_t st);
