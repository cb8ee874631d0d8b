// post.cu -- the step that follows match() in every pipeline of the reference: bad-pixel
// correction of the displacement maps (UMPA/align.py:661-732, called from UMPA_normal /
// UMPA_nobias, align.py:58-60, 111-114), with the optional bias subtraction fused in, on the
// device: the maps never leave HBM between match() and the correction.
#include <algorithm>

#include "common.cuh"

namespace {

// out = in - bias (bias may be NULL); one pass, so that the correction below reads a finished map
__global__ void subtract_kernel(const double *in, const double *bias, double *out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bias ? in[i] - bias[i] : in[i];
}

// One iteration of correct_bad_pixels on a stack of `nimg` images of N0 x N1 (dims = (-2,-1)):
// a value outside [lo, hi] is replaced by the median of its four neighbours, reflected at the
// edges the way the reference indexes them (|i-1| above / left, N-2 instead of N below / right,
// align.py:720-727).  The neighbours are read from the input of this iteration (the reference
// gathers all of them before it assigns, align.py:713-731).
// The set of bad pixels is fixed by the FIRST input (align.py:707-708: the mask is taken once and the same
// pixels are revisited by every iteration); `bad` is that mask, or NULL to test `in` itself (one iteration).
__global__ void mask_kernel(const double *in, unsigned char *bad, size_t n, double lo, double hi)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bad[i] = in[i] < lo || in[i] > hi;
}

__global__ void bad_pixel_kernel(const double *in, const unsigned char *bad, double *out, int N0, int N1, size_t nimg,
                                 double lo, double hi)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N1) return;
    for (size_t z = blockIdx.z; z < nimg; z += gridDim.z) {
        const double *img = in + z * (size_t)N0 * N1;
        const double v = img[(size_t)i * N1 + j];
        double r = v;
        if (bad ? bad[z * (size_t)N0 * N1 + (size_t)i * N1 + j] != 0 : (v < lo || v > hi)) {
            const int iu = abs(i - 1), id = i + 1 == N0 ? N0 - 2 : i + 1;
            const int jl = abs(j - 1), jr = j + 1 == N1 ? N1 - 2 : j + 1;
            const double a = img[(size_t)iu * N1 + j], b = img[(size_t)id * N1 + j];
            const double c = img[(size_t)i * N1 + jl], d = img[(size_t)i * N1 + jr];
            // median of four = mean of the two middle values (numpy: mean of the sorted middle pair);
            // sort the pairs, then the middle two are max(lows) and min(highs) -- in either order
            const double lo1 = fmin(a, b), hi1 = fmax(a, b), lo2 = fmin(c, d), hi2 = fmax(c, d);
            r = (fmax(lo1, lo2) + fmin(hi1, hi2)) * .5;
        }
        out[z * (size_t)N0 * N1 + (size_t)i * N1 + j] = r;
    }
}

}  // namespace

extern "C" UMPA_API int umpa_correct_bad_pixels(const double *img, const double *bias, double *out, double *scratch,
                                                int64_t nimg, int N0, int N1, double lo, double hi, int iterations,
                                                void *stream)
{
    if (!img || !out) { umpa_set_error("umpa_correct_bad_pixels: NULL argument"); return UMPA_ERR_ARG; }
    if (nimg < 0 || N0 < 0 || N1 < 0 || iterations < 0) { umpa_set_error("umpa_correct_bad_pixels: negative size"); return UMPA_ERR_ARG; }
    const size_t n = (size_t)nimg * N0 * N1;
    if (n == 0) return UMPA_OK;
    if ((N0 < 2 || N1 < 2) && iterations > 0) { umpa_set_error("umpa_correct_bad_pixels: images must be at least 2 x 2"); return UMPA_ERR_ARG; }
    if (iterations > 1 && !scratch) { umpa_set_error("umpa_correct_bad_pixels: scratch needed for more than one iteration"); return UMPA_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    if (iterations == 0) {
        if (bias || img != out) subtract_kernel<<<nblk, 256, 0, st>>>(img, bias, out, n);
        UMPA_CUDA(cudaGetLastError());
        return UMPA_OK;
    }
    // iteration `it` writes to out or scratch, alternating, so that the last one lands in `out`; the
    // correction reads neighbours, so its input must be another buffer than its destination
    auto dest = [&](int it) { return ((iterations - 1 - it) % 2 == 0) ? out : scratch; };
    const bool stage = bias != nullptr || img == dest(0);
    if ((stage || iterations > 1) && !scratch) {
        umpa_set_error("umpa_correct_bad_pixels: scratch buffer needed (bias, in-place input or more than one iteration)");
        return UMPA_ERR_ARG;
    }
    const double *cur = img;
    if (stage) {
        double *tmp = dest(0) == out ? scratch : out;
        subtract_kernel<<<nblk, 256, 0, st>>>(img, bias, tmp, n);
        cur = tmp;
    }
    unsigned char *bad = nullptr;                // more than one iteration: the mask of the first input
    if (iterations > 1) {
        UMPA_CUDA(cudaMallocAsync((void **)&bad, n, st));
        mask_kernel<<<nblk, 256, 0, st>>>(cur, bad, n, lo, hi);
    }
    dim3 grid((N1 + 127) / 128, N0, (unsigned)std::min<int64_t>(nimg, 64));
    for (int it = 0; it < iterations; it++) {
        bad_pixel_kernel<<<grid, 128, 0, st>>>(cur, bad, dest(it), N0, N1, (size_t)nimg, lo, hi);
        cur = dest(it);
    }
    if (bad) UMPA_CUDA(cudaFreeAsync(bad, st));
    UMPA_CUDA(cudaGetLastError());
    return UMPA_OK;
}
