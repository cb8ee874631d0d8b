// host convert benchmark: dst[i] = (float)(src[i] - c), multi-threaded
#include <thread>
#include <vector>
#include <cstddef>
#include <cstdint>
#include <immintrin.h>
extern "C" void conv_rows(float *dst, const double *src, size_t n, double c, int threads)
{
    auto work = [&](size_t lo, size_t hi) {
        const __m256d vc = _mm256_set1_pd(c);
        size_t i = lo;
        for (; i < hi && ((uintptr_t)(dst + i) & 31); i++) dst[i] = (float)(src[i] - c);
        for (; i + 8 <= hi; i += 8) {
            __m256d a = _mm256_sub_pd(_mm256_loadu_pd(src + i), vc);
            __m256d b = _mm256_sub_pd(_mm256_loadu_pd(src + i + 4), vc);
            __m128 fa = _mm256_cvtpd_ps(a), fb = _mm256_cvtpd_ps(b);
            _mm256_stream_ps(dst + i, _mm256_set_m128(fb, fa));
        }
        for (; i < hi; i++) dst[i] = (float)(src[i] - c);
    };
    std::vector<std::thread> th;
    size_t per = ((n / threads) + 7) & ~(size_t)7;
    for (int t = 0; t < threads; t++) {
        size_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (t == threads - 1) hi = n;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &t : th) t.join();
}
