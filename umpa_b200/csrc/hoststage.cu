// hoststage.cu -- host-side helpers of the pipelined upload (capi.cu, streamed_match): the centring
// constants of a frame from its sampled rows, and the FP64 -> centred FP32 conversion that lets part
// of a stack cross PCIe at half the bytes.  Plain host code (no device code in this unit).
#if defined(__x86_64__)
#include <immintrin.h>
#define UMPA_HOST_AVX2 1         // x86-64 hosts: AVX2 path chosen at run time; other hosts (aarch64: Grace + Blackwell) scalar
#else
#define UMPA_HOST_AVX2 0
#endif
#include <stddef.h>
#include <stdint.h>

#include "common.cuh"

// mean of rows 0, step, 2 step, ... of one H x W frame (fixed summation order: row by row, 4 lanes;
// float32 frames are widened element by element, so they give the mean of the widened frame bit for bit)
// A NaN / Inf pixel (a dead pixel divided by its flat field) must stay a local defect -- the windows that touch
// it -- as it is in the reference; it counts as 0 here so that it does not poison the frame's constant.
static inline double finite_or_zero(double v) { return v - v == 0. ? v : 0.; }

template <typename T>
static double sampled_mean(const T *frame, int H, int W, int step)
{
    double s = 0.;
    size_t count = 0;
    for (int y = 0; y < H; y += step) {
        const T *row = frame + (size_t)y * W;
        double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
        int x = 0;
        for (; x + 4 <= W; x += 4) {
            a0 += finite_or_zero((double)row[x]); a1 += finite_or_zero((double)row[x + 1]);
            a2 += finite_or_zero((double)row[x + 2]); a3 += finite_or_zero((double)row[x + 3]);
        }
        for (; x < W; x++) a0 += finite_or_zero((double)row[x]);
        s += (a0 + a1) + (a2 + a3);
        count += (size_t)W;
    }
    return s / (double)count;
}

double host_sampled_mean(const double *frame, int H, int W, int step) { return sampled_mean(frame, H, W, step); }
double host_sampled_mean_f32(const float *frame, int H, int W, int step) { return sampled_mean(frame, H, W, step); }

namespace {

void convert_scalar(float *dst, const double *src, size_t n, double c)
{
    for (size_t i = 0; i < n; i++) dst[i] = (float)(src[i] - c);
}

#if UMPA_HOST_AVX2
__attribute__((target("avx2"))) void convert_avx2(float *dst, const double *src, size_t n, double c)
{
    const __m256d vc = _mm256_set1_pd(c);
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 31); i++) dst[i] = (float)(src[i] - c);
    for (; i + 8 <= n; i += 8) {
        const __m128 lo = _mm256_cvtpd_ps(_mm256_sub_pd(_mm256_loadu_pd(src + i), vc));
        const __m128 hi = _mm256_cvtpd_ps(_mm256_sub_pd(_mm256_loadu_pd(src + i + 4), vc));
        _mm256_stream_ps(dst + i, _mm256_set_m128(hi, lo));      // the staging buffer is write-only for the CPU
    }
    for (; i < n; i++) dst[i] = (float)(src[i] - c);
    _mm_sfence();
}
#endif

void convert_scalar(float *dst, const float *src, size_t n, double c)
{
    for (size_t i = 0; i < n; i++) dst[i] = (float)((double)src[i] - c);
}

#if UMPA_HOST_AVX2
__attribute__((target("avx2"))) void convert_avx2(float *dst, const float *src, size_t n, double c)
{
    const __m256d vc = _mm256_set1_pd(c);
    size_t i = 0;
    for (; i < n && ((uintptr_t)(dst + i) & 31); i++) dst[i] = (float)((double)src[i] - c);
    for (; i + 8 <= n; i += 8) {
        const __m256 v = _mm256_loadu_ps(src + i);
        const __m128 lo = _mm256_cvtpd_ps(_mm256_sub_pd(_mm256_cvtps_pd(_mm256_castps256_ps128(v)), vc));
        const __m128 hi = _mm256_cvtpd_ps(_mm256_sub_pd(_mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)), vc));
        _mm256_stream_ps(dst + i, _mm256_set_m128(hi, lo));
    }
    for (; i < n; i++) dst[i] = (float)((double)src[i] - c);
    _mm_sfence();
}
#else
inline void convert_avx2(float *dst, const double *src, size_t n, double c) { convert_scalar(dst, src, n, c); }
inline void convert_avx2(float *dst, const float *src, size_t n, double c) { convert_scalar(dst, src, n, c); }
#endif

template <typename T>
void center_rows(float *dst, const T *src, int rows, int W, int pitch, double c)
{
#if UMPA_HOST_AVX2
    static const bool avx2 = __builtin_cpu_supports("avx2");
#else
    const bool avx2 = false;
#endif
    if (pitch == W) {
        if (avx2) convert_avx2(dst, src, (size_t)rows * W, c); else convert_scalar(dst, src, (size_t)rows * W, c);
        return;
    }
    for (int y = 0; y < rows; y++) {
        float *d = dst + (size_t)y * pitch;
        if (avx2) convert_avx2(d, src + (size_t)y * W, W, c); else convert_scalar(d, src + (size_t)y * W, W, c);
        for (int x = W; x < pitch; x++) d[x] = 0.f;
    }
}

}  // namespace

// dst[y][x] = (float)(src[y][x] - c) for `rows` rows of W doubles -> rows of `pitch` floats (zero padded).
// Same arithmetic as center_frames (table_path.cu): FP64 subtraction, one rounding to FP32.
void host_center_rows(float *dst, const double *src, int rows, int W, int pitch, double c) { center_rows(dst, src, rows, W, pitch, c); }
// float32 source rows: widened, centred, rounded once -- what the device does to rows that went up raw (center_inplace)
void host_center_rows_f32(float *dst, const float *src, int rows, int W, int pitch, double c) { center_rows(dst, src, rows, W, pitch, c); }

// ---- test hooks (no GPU needed): the two host helpers above through the C ABI ----------------
extern "C" UMPA_API double umpa_host_sampled_mean(const double *frame, int H, int W, int step)
{
    return host_sampled_mean(frame, H, W, step);
}

extern "C" UMPA_API double umpa_host_sampled_mean_f32(const float *frame, int H, int W, int step)
{
    return host_sampled_mean_f32(frame, H, W, step);
}

extern "C" UMPA_API void umpa_host_center_rows_f32(float *dst, const float *src, int rows, int W, int pitch, double c)
{
    host_center_rows_f32(dst, src, rows, W, pitch, c);
}

extern "C" UMPA_API void umpa_host_center_rows(float *dst, const double *src, int rows, int W, int pitch, double c)
{
    host_center_rows(dst, src, rows, W, pitch, c);
}
