// lazy_path.cu -- the general CUDA path: one thread per output pixel walks the integer
// shift grid and evaluates the model cost on demand, in FP64 and in the reference's
// summation order.  Covers everything the reference's cost() functions cover: all three
// models, masks, per-frame positions / ragged frames, reference_shift, strided ROIs.
//
// Reference behaviour restated here:
//   ModelNoDF::cost      UMPA/lib/Model.cpp:359-509
//   ModelDF::cost        UMPA/lib/Model.cpp:631-862
//   ModelDFKernel::cost  UMPA/lib/Model.cpp:997-1151, kernel ctor 88-117, Utils.cpp:46-50,85-117
//   ModelBase::coverage  UMPA/lib/Model.cpp:273-314
//   Model*::min          UMPA/lib/Model.cpp:562-578, 923-940, 1222-1238
//   pixel loop           UMPA/model.pyx:476-492
#include <algorithm>

#include "walk.cuh"

namespace {

__device__ __forceinline__ bool frame_reaches(const LazyView &m, int k, int i, int j)
{
    const int ri = i - m.pos[2 * k], rj = j - m.pos[2 * k + 1];
    return !(ri - m.padding < 0 || ri + m.padding > m.dim[2 * k] || rj - m.padding < 0 ||
             rj + m.padding > m.dim[2 * k + 1]);
}

__device__ __forceinline__ double mix_weights(double a, double b) { return a * b / (a + b + 1e-8); }

__device__ __forceinline__ int shift_status(int max_shift, int si, int sj)
{
    if (si <= -max_shift || si >= max_shift) return UMPA_ST_BOUND;
    if (sj <= -max_shift) return UMPA_ST_BOUND | UMPA_ST_DIM;
    if (sj >= max_shift) return UMPA_ST_BOUND | UMPA_ST_DIM | UMPA_ST_POS;
    return UMPA_ST_OK;
}

// 17x17 blur of frame `img` at (i,j); kernel normalised (Utils.cpp:85-97)
// The per-pixel kernel is stored interleaved over the threads of a launch (element n of
// this thread at kern[n*ks]) so that a warp's reads of "its" element n coalesce.
__device__ inline double blur_at(const double *__restrict__ img, int i, int j, int W,
                                 const double *__restrict__ kern, size_t ks)
{
    double out = 0.;
    for (int r = -UMPA_KWS; r <= UMPA_KWS; r++) {
        const double *row = img + (size_t)(i + r) * W + j;
        const double *kr = kern + (size_t)(UMPA_KSIDE * (r + UMPA_KWS) + UMPA_KWS) * ks;
        for (int q = -UMPA_KWS; q <= UMPA_KWS; q++) out += kr[(ptrdiff_t)q * (ptrdiff_t)ks] * row[q];
    }
    return out;
}

// The same blur for CB consecutive window elements (i, j .. j+CB-1) at once: every kernel tap is loaded once
// and used CB times (the per-pixel kernels live in global memory: one load per MAC made the DFKernel
// evaluation L2-bandwidth bound), the frame row segment sits in registers.  Each element's sum runs over
// the taps in the order of blur_at, so the results are bit-identical.
template <int CB>
__device__ inline void blur_row_at(const double *__restrict__ img, int i, int j, int W,
                                   const double *__restrict__ kern, size_t ks, double (&out)[CB])
{
#pragma unroll
    for (int c = 0; c < CB; c++) out[c] = 0.;
    for (int r = -UMPA_KWS; r <= UMPA_KWS; r++) {
        const double *row = img + (size_t)(i + r) * W + j - UMPA_KWS;
        const double *kr = kern + (size_t)(UMPA_KSIDE * (r + UMPA_KWS)) * ks;
        double seg[UMPA_KSIDE + CB - 1];
#pragma unroll
        for (int t = 0; t < UMPA_KSIDE + CB - 1; t++) seg[t] = row[t];
#pragma unroll
        for (int q = 0; q < UMPA_KSIDE; q++) {
            const double kv = kr[(size_t)q * ks];
#pragma unroll
            for (int c = 0; c < CB; c++) out[c] += kv * seg[q + c];
        }
    }
}

// mask-weighted blur (Utils.cpp:103-117)
__device__ inline double weighted_blur_at(const double *__restrict__ img, const double *__restrict__ wgt,
                                          int i, int j, int W, const double *__restrict__ kern, size_t ks)
{
    double out = 0., w = 0.;
    for (int r = -UMPA_KWS; r <= UMPA_KWS; r++) {
        const size_t o = (size_t)(i + r) * W + j;
        const double *kr = kern + (size_t)(UMPA_KSIDE * (r + UMPA_KWS) + UMPA_KWS) * ks;
        for (int q = -UMPA_KWS; q <= UMPA_KWS; q++) {
            const double kv = kr[(ptrdiff_t)q * (ptrdiff_t)ks];
            out += kv * img[o + q] * wgt[o + q];
            w += kv * wgt[o + q];
        }
    }
    return out / w;
}

// One cost evaluation at raw pixel (i,j).
// KIND: the model kind when known at compile time (the match kernel is instantiated per kind, so that the
// NoDF / DF variants do not carry the registers of the DFKernel blur), -1 = read m.kind at run time (hooks).
template <int KIND = -1>
struct LazyEval {
    const LazyView &m;
    int i, j;
    const double *kern;       // DFKernel: this pixel's normalised 17x17 kernel (interleaved)
    size_t ks;                // stride between its elements

    // DFKernel, unmasked: the window sums of one frame with the blur taken CB window columns at a time
    // (blur_row_at); the last chunk of a row is moved back inside the window, elements are added to the
    // sums in the reference's order (row-major), so nothing changes bitwise.
    template <int CB>
    __device__ void dfk_window(const double *__restrict__ R, const double *__restrict__ S, int r0, int rc0, int s0,
                               int sc0, int W, int K, double &t1, double &t3, double &t5) const
    {
        for (int a = 0; a < K; a++)
            for (int b0 = 0; b0 < K; b0 += CB) {
                double rb[CB];
                const int bs = min(b0, K - CB);
                blur_row_at<CB>(R, r0 + a, rc0 + bs, W, kern, ks, rb);
#pragma unroll
                for (int c = 0; c < CB; c++) {
                    const int b = bs + c;
                    if (b >= b0) {
                        const double w = m.win[a * K + b], s = S[(size_t)(s0 + a) * W + sc0 + b], r = rb[c];
                        t1 += w * s * s;
                        t3 += w * r * r;
                        t5 += w * r * s;
                    }
                }
            }
    }

    __device__ int operator()(int si, int sj, double &cost, FitArgs &args) const
    {
        const int st = shift_status(m.max_shift, si, sj);
        if (st != UMPA_ST_OK) return st;
        const int kind = KIND >= 0 ? KIND : m.kind;
        const int Nw = m.Nw, K = 2 * Nw + 1;
        int ri, rj, qi, qj;       // centres of the reference and of the sample window
        if (m.refshift) { ri = i; rj = j; qi = i - si; qj = j - sj; }
        else            { ri = i + si; rj = j + sj; qi = i; qj = j; }

        double t1 = 0., t2 = 0., t3 = 0., t4 = 0., t5 = 0., t6 = 0.;
        double wt = m.masked ? 0. : (double)m.Na;

        for (int k = 0; k < m.Na; k++) {
            if (!frame_reaches(m, k, i, j)) continue;
            const int W = m.dim[2 * k + 1];
            const int pi = m.pos[2 * k], pj = m.pos[2 * k + 1];
            const double *__restrict__ R = m.ref[k];
            const double *__restrict__ S = m.sam[k];
            const double *__restrict__ M = m.masked ? m.mask[k] : nullptr;
            const int r0 = ri - pi - Nw, rc0 = rj - pj - Nw;      // top-left of the reference window
            const int s0 = qi - pi - Nw, sc0 = qj - pj - Nw;      // top-left of the sample window

            if (kind == UMPA_DF) {
                // One pass over the window.  The reference takes the weighted mean of the reference window in a
                // loop of its own (Model.cpp:722-735) and then sums w*r again as s6 (unmasked branch): the two
                // accumulate the same products in the same order, so `mean` before its division IS that sum.
                double mean = 0., den = 0.;
                double s2 = 0., s4 = 0., s6 = 0.;
                for (int a = 0; a < K; a++)
                    for (int b = 0; b < K; b++) {
                        const size_t nr = (size_t)(r0 + a) * W + rc0 + b;
                        const size_t ns = (size_t)(s0 + a) * W + sc0 + b;
                        const double w = m.win[a * K + b], s = S[ns], r = R[nr];
                        mean += w * r;
                        den += w;
                        if (!m.masked) {
                            t1 += w * s * s;
                            t3 += w * r * r;
                            s4 += w * s;
                            t5 += w * r * s;
                        } else {
                            const double g = mix_weights(M[nr], M[ns]);
                            t1 += g * w * s * s;
                            s2 += g * w;
                            t3 += g * w * r * r;
                            s4 += g * w * s;
                            t5 += g * w * r * s;
                            s6 += g * w * r;
                            wt += g * w;
                        }
                    }
                if (!m.masked) s6 = mean;
                mean /= den;
                t2 += m.masked ? mean * mean * s2 : mean * mean;
                t4 += mean * s4;
                t6 += mean * s6;
            } else if (kind == UMPA_DFKERNEL && !m.masked && K >= 4) {
                if (K >= 7) dfk_window<7>(R, S, r0, rc0, s0, sc0, W, K, t1, t3, t5);
                else dfk_window<4>(R, S, r0, rc0, s0, sc0, W, K, t1, t3, t5);
            } else {
                for (int a = 0; a < K; a++)
                    for (int b = 0; b < K; b++) {
                        const size_t nr = (size_t)(r0 + a) * W + rc0 + b;
                        const size_t ns = (size_t)(s0 + a) * W + sc0 + b;
                        const double w = m.win[a * K + b], s = S[ns];
                        double r;
                        if (kind == UMPA_DFKERNEL)
                            r = m.masked ? weighted_blur_at(R, M, r0 + a, rc0 + b, W, kern, ks)
                                         : blur_at(R, r0 + a, rc0 + b, W, kern, ks);
                        else
                            r = R[nr];
                        if (!m.masked) {
                            t1 += w * s * s;
                            t3 += w * r * r;
                            t5 += w * r * s;
                        } else {
                            const double g = mix_weights(M[nr], M[ns]);
                            t1 += g * w * s * s;
                            t3 += g * w * r * r;
                            t5 += g * w * r * s;
                            wt += g * w;
                        }
                    }
            }
        }
        if (kind == UMPA_DF) {
            const double den = t2 * t3 - t6 * t6;
            const double Kc = (t2 * t5 - t4 * t6) / den;
            const double beta = (t3 * t4 - t5 * t6) / den;
            args.t = beta + Kc;
            args.v = Kc / args.t;
            cost = (t1 + beta * beta * t2 + Kc * Kc * t3 - 2. * beta * t4 - 2. * Kc * t5 + 2. * beta * Kc * t6) / wt;
        } else {
            args.t = t5 / t3;
            cost = (t1 - t5 * args.t) / wt;
        }
        return UMPA_ST_OK;
    }
};

__device__ inline void build_blur_kernel(double a, double b, double c, double *kern, size_t ks)
{
    double norm = 0.;
    for (int r = 0; r < UMPA_KSIDE; r++)
        for (int q = 0; q < UMPA_KSIDE; q++) {
            const int i = r - UMPA_KWS, j = q - UMPA_KWS;
            const double v = exp(-a * i * i - b * i * j - c * j * j);
            kern[(size_t)(r * UMPA_KSIDE + q) * ks] = v;
            norm += v;
        }
    for (int n = 0; n < UMPA_KSIDE * UMPA_KSIDE; n++) kern[(size_t)n * ks] /= norm;
}

// kern_ws: workspace for the per-pixel blur kernels (DFKernel), 289 doubles per launched thread.
// row0: first output row of this launch (DFKernel is launched in row bands to bound kern_ws).
template <int KIND>
__global__ void __launch_bounds__(128)
lazy_match_kernel(LazyView m, RoiView roi, umpa_outputs out, double *kern_ws, int row0)
{
    const int xj = blockIdx.x * blockDim.x + threadIdx.x;
    const int xi = row0 + blockIdx.y;
    if (xj >= roi.N1 || xi >= roi.N0) return;
    const size_t n = (size_t)xi * roi.N1 + xj;
    if (roi.cover && roi.cover[n] < roi.cover_threshold) return;   // model.pyx:480; outputs stay zero
    if (roi.dirty && (roi.dirty[n] != 0) != (roi.dirty_want != 0)) return;   // mixed path: the table kernels own this pixel

    double *kern = nullptr;
    const size_t ks = (size_t)gridDim.x * gridDim.y * blockDim.x;
    if (KIND == UMPA_DFKERNEL) {
        kern = kern_ws + ((size_t)blockIdx.y * gridDim.x * blockDim.x + xj);
        build_blur_kernel(roi.abc[3 * n], roi.abc[3 * n + 1], roi.abc[3 * n + 2], kern, ks);
    }
    LazyEval<KIND> eval{m, roi.off0 + roi.step0 * xi, roi.off1 + roi.step1 * xj, kern, ks};
    FitArgs args{0., 0.};
    double d[25], uv[2] = {roi.uv0[0], roi.uv0[1]}, f = 0.;
    int ncalls;
    WalkState ws;
    const int st = walk_search<false>(eval, args, f, uv, d, ncalls, ws);
    store_debug(out, n, d, ws);                    // (before the fit: it reuses the cache's cells)
    if (ws.finished) walk_refine(m.subpx, m.quad, d, ws, f, uv);
    store_pixel(out, n, m.kind, st, f, args, uv, ncalls);
}

__global__ void lazy_cost_kernel(LazyView m, int i, int j, int si, int sj, double a, double b, double c,
                                 double *res, double *kern)
{
    if (m.kind == UMPA_DFKERNEL) build_blur_kernel(a, b, c, kern, 1);
    LazyEval<> eval{m, i, j, kern, 1};
    FitArgs args{0., 0.};
    double cost = 0.;
    const int st = eval(si, sj, cost, args);
    res[0] = cost; res[1] = args.t; res[2] = args.v; res[3] = (double)st;
}

__global__ void lazy_min_kernel(LazyView m, int i, int j, double *io, double *kern)
{
    // io: [0..6] values, [7..8] uv, [9..33] d, [34..49] a, [50] ncalls, [51] status
    if (m.kind == UMPA_DFKERNEL) build_blur_kernel(io[4], io[5], io[6], kern, 1);
    LazyEval<> eval{m, i, j, kern, 1};
    FitArgs args{0., 0.};
    double d[25], uv[2] = {io[7], io[8]}, f = 0.;
    int ncalls;
    WalkState ws;
    const int st = walk_search<false>(eval, args, f, uv, d, ncalls, ws);
    for (int t = 0; t < 25; t++) io[9 + t] = walk_cache_get(d, ws, t);
    for (int t = 0; t < 16; t++) io[34 + t] = ws.finished ? walk_block_get(d, ws, t >> 2, t & 3) : 0.;
    if (ws.finished) walk_refine(m.subpx, m.quad, d, ws, f, uv);
    io[0] = f; io[1] = args.t; io[2] = uv[1]; io[3] = uv[0];
    if (m.kind == UMPA_DF) io[4] = args.v;
    io[7] = uv[0]; io[8] = uv[1];

    io[50] = ncalls; io[51] = st;
}

__global__ void coverage_kernel(LazyView m, RoiView roi, double *out)
{
    const int xj = blockIdx.x * blockDim.x + threadIdx.x;
    const int xi = blockIdx.y;
    if (xj >= roi.N1 || xi >= roi.N0) return;
    const int i = roi.off0 + roi.step0 * xi, j = roi.off1 + roi.step1 * xj;
    double wt = 0.;
    for (int k = 0; k < m.Na; k++) {
        if (!frame_reaches(m, k, i, j)) continue;
        if (!m.masked) wt += 1.;
        else wt += m.mask[k][(size_t)(i - m.pos[2 * k]) * m.dim[2 * k + 1] + (j - m.pos[2 * k + 1])];
    }
    out[(size_t)xi * roi.N1 + xj] = wt;
}

LazyView make_view(const umpa_model *m)
{
    LazyView v;
    v.kind = m->kind; v.Na = m->Na; v.Nw = m->Nw; v.max_shift = m->max_shift; v.padding = m->padding;
    v.subpx = m->subpx; v.refshift = m->refshift; v.masked = m->masked ? 1 : 0;
    v.sam = m->d_sam_ptrs; v.ref = m->d_ref_ptrs; v.mask = m->d_mask_ptrs;
    v.dim = m->d_dim; v.pos = m->d_pos; v.win = m->d_win; v.quad = m->d_quad;
    return v;
}

}  // namespace

int lazy_match(umpa_model *m, const RoiView &roi, const umpa_outputs &out, cudaStream_t st)
{
    const int threads = 128;
    const int gx = (roi.N1 + threads - 1) / threads;
    int band = roi.N0;
    double *ws = nullptr;
    if (m->kind == UMPA_DFKERNEL) {
        if (!roi.abc) { umpa_set_error("abc array has to be provided"); return UMPA_ERR_ARG; }
        const size_t per_row = (size_t)gx * threads * UMPA_KSIDE * UMPA_KSIDE * sizeof(double);
        band = (int)std::max<size_t>(1, std::min<size_t>(roi.N0, ((size_t)1 << 30) / per_row));
        int rc = scratch_reserve(m, m->tabX, per_row * band);
        if (rc) return rc;
        ws = (double *)m->tabX.p;
    }
    for (int row0 = 0; row0 < roi.N0; row0 += band) {
        dim3 grid(gx, std::min(band, roi.N0 - row0));
        if (m->kind == UMPA_NODF) lazy_match_kernel<UMPA_NODF><<<grid, threads, 0, st>>>(make_view(m), roi, out, ws, row0);
        else if (m->kind == UMPA_DF) lazy_match_kernel<UMPA_DF><<<grid, threads, 0, st>>>(make_view(m), roi, out, ws, row0);
        else lazy_match_kernel<UMPA_DFKERNEL><<<grid, threads, 0, st>>>(make_view(m), roi, out, ws, row0);
        UMPA_CUDA(cudaGetLastError());
        m->last_launches += 1;
    }
    return UMPA_OK;
}

int coverage_map(umpa_model *m, const RoiView &roi, double *out_dev, cudaStream_t st)
{
    const int threads = 128;
    dim3 grid((roi.N1 + threads - 1) / threads, roi.N0);
    coverage_kernel<<<grid, threads, 0, st>>>(make_view(m), roi, out_dev);
    UMPA_CUDA(cudaGetLastError());
    return UMPA_OK;
}

int lazy_cost(umpa_model *m, int i, int j, int si, int sj, const double abc[3], double values[3], int *status)
{
    double *buf = nullptr;
    UMPA_CUDA(cudaMalloc(&buf, (4 + UMPA_KSIDE * UMPA_KSIDE) * sizeof(double)));
    lazy_cost_kernel<<<1, 1>>>(make_view(m), i, j, si, sj, abc ? abc[0] : 0., abc ? abc[1] : 0.,
                               abc ? abc[2] : 0., buf, buf + 4);
    double h[4];
    cudaError_t e = cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    UMPA_CUDA(e);
    values[0] = h[0]; values[1] = h[1]; values[2] = h[2];
    if (status) *status = (int)h[3];
    return UMPA_OK;
}

int lazy_min(umpa_model *m, int i, int j, double *values, double uv[2], double *dd, double *da, int *ncalls, int *ok)
{
    const int np = m->kind == UMPA_NODF ? 4 : (m->kind == UMPA_DF ? 5 : 7);
    double h[52] = {0.};
    if (m->kind == UMPA_DFKERNEL) { h[4] = values[4]; h[5] = values[5]; h[6] = values[6]; }
    h[7] = uv[0]; h[8] = uv[1];
    double *buf = nullptr;
    UMPA_CUDA(cudaMalloc(&buf, (52 + UMPA_KSIDE * UMPA_KSIDE) * sizeof(double)));
    cudaError_t e = cudaMemcpy(buf, h, sizeof(h), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        lazy_min_kernel<<<1, 1>>>(make_view(m), i, j, buf, buf + 52);
        e = cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
    }
    cudaFree(buf);
    UMPA_CUDA(e);
    for (int t = 0; t < np; t++) if (!(m->kind == UMPA_DFKERNEL && t >= 4)) values[t] = h[t];
    uv[0] = h[7]; uv[1] = h[8];
    if (dd) for (int t = 0; t < 25; t++) dd[t] = h[9 + t];
    if (da) for (int t = 0; t < 16; t++) da[t] = h[34 + t];
    if (ncalls) *ncalls = (int)h[50];
    if (ok) *ok = ((int)h[51] & UMPA_ST_OK) ? 1 : 0;
    return UMPA_OK;
}
