// walk.cuh -- per-pixel integer-shift walk and sub-pixel refinement (device, FP64).
//
// Re-implements the behaviour of discrete_2d_minimizer (UMPA/lib/Optim.cpp:233-479),
// spmin (Optim.cpp:41-130) and spmin_quad (Optim.cpp:155-185) for one CUDA thread per
// pixel.  The cost function is a functor, so the same state machine runs on top of the
// lazy FP64 evaluator (lazy_path.cu) and of the FP32 shift tables (table_path.cu).
#pragma once
#include "common.cuh"

struct FitArgs {            // what the reference keeps in CostArgs{NoDF,DF,DFKernel}: t (and v)
    double t, v;
};

// ---- cubic B-spline surface through a 4x4 block -----------------------------------
// f(x,y) = sum_ij B_i(x) B_j(y) a[4i+j] / 36 with x along rows; B_* are the four uniform
// cubic B-spline pieces on [0,1] for samples at -1,0,1,2.  BSPL[n][s]: coefficient of t^n
// of piece s, times 6.
__device__ __forceinline__ double bspl_coef(int n, int s)
{
    // {1,4,1,0}, {-3,0,3,0}, {3,-6,3,0}, {-1,3,-3,1}
    const int tab[16] = {1, 4, 1, 0, -3, 0, 3, 0, 3, -6, 3, 0, -1, 3, -3, 1};
    return (double)tab[4 * n + s];
}

__device__ inline double subpixel_spline(const double *a, double *pos)
{
    // c[4m+n] multiplies x^n y^m
    double tmp[16], c[16];
#pragma unroll
    for (int n = 0; n < 4; n++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double s = 0.;
#pragma unroll
            for (int i = 0; i < 4; i++) s += bspl_coef(n, i) * a[4 * i + j];
            tmp[4 * n + j] = s;               // x-power n, column sample j
        }
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
        for (int n = 0; n < 4; n++) {
            double s = 0.;
#pragma unroll
            for (int j = 0; j < 4; j++) s += bspl_coef(m, j) * tmp[4 * n + j];
            c[4 * m + n] = s;
        }
    double x = pos[0], y = pos[1];
    for (int it = 0; it <= 20; it++) {
        double A[4], A1[4], A2[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const double c0 = c[4 * m], c1 = c[4 * m + 1], c2 = c[4 * m + 2], c3 = c[4 * m + 3];
            A[m] = c0 + x * (c1 + x * (c2 + x * c3));
            A1[m] = c1 + x * (2. * c2 + 3. * x * c3);
            A2[m] = 2. * c2 + 6. * x * c3;
        }
        const double fx = A1[0] + y * (A1[1] + y * (A1[2] + y * A1[3]));
        const double fxx = A2[0] + y * (A2[1] + y * (A2[2] + y * A2[3]));
        const double fy = A[1] + y * (2. * A[2] + 3. * y * A[3]);
        const double fxy = A1[1] + y * (2. * A1[2] + 3. * y * A1[3]);
        const double fyy = 2. * A[2] + 6. * y * A[3];
        const double det = fxx * fyy - fxy * fxy;
        const double dx = (fxy * fy - fyy * fx) / det;
        const double dy = (fxy * fx - fxx * fy) / det;
        x += dx;
        y += dy;
        if (dx * dx + dy * dy < 1e-8) break;          // absolute stop, Optim.cpp:85,123
    }
    pos[0] = x;
    pos[1] = y;
    double f = 0., yp = 1.;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        f += yp * (c[4 * m] + x * (c[4 * m + 1] + x * (c[4 * m + 2] + x * c[4 * m + 3])));
        yp *= y;
    }
    return f / 36.;
}

// ---- least-squares quadratic through a 4x4 block ------------------------------------
// p = 400 * pinv(A) a for the basis [1, i, j, i^2, ij, j^2] on i,j in {-1,0,1,2}
// (i along rows).  The 6x16 integer matrix is 400*(A^T A)^-1 A^T, generated on the host
// when a model is created (quad_matrix in capi.cu) and passed in as `quad` (device memory).
__device__ inline double subpixel_quadratic(const double *__restrict__ c_quad, const double *a, double *pos)
{
    double p[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double s = 0.;
#pragma unroll
        for (int n = 0; n < 16; n++) s += c_quad[16 * r + n] * a[n];
        p[r] = s;
    }
    const double det = 4. * p[3] * p[5] - p[4] * p[4];
    // reference quirk kept: pos[0] gets the column solution, pos[1] the row solution
    pos[0] = -(2. * p[3] * p[2] - p[4] * p[1]) / det;
    pos[1] = -(2. * p[5] * p[1] - p[4] * p[2]) / det;
    return (p[0] + .5 * (p[2] * pos[0] + p[1] * pos[1])) / 400.;
}

// ---- the walk --------------------------------------------------------------------------
// The reference keeps d, a 5x5 cache of costs centred on the current integer shift (row-major, centre 12,
// -1 = not evaluated), and physically shifts it by one row / column on every step.  Here the cache is a
// RING: the 25 storage cells are addressed modulo 5 from an origin (b0, b1) that moves with the centre, and
// which cells hold an evaluated cost is a 25-bit register mask in the reference's (logical) order -- a step
// rotates the mask and moves the origin, nothing is copied and nothing has to be initialised.
// walk_cache_get() reads the cache the way the reference would have left it.
// axis 0 scans the column shift, axis 1 the row shift.  `keep` is the reference's args_copy: the fit
// parameters of the best shift seen so far -- deliberately NOT refreshed on a restart (Optim.cpp:364-377),
// which the reference's outputs depend on.
//
// The reference calls the cost function from four places (centre, minus neighbour, plus
// neighbour, 4x4 fill).  On a GPU that would serialise the lanes of a warp that happen to be
// at different call sites, so the same control flow is written as a state machine with ONE
// evaluation site: each lane advances its state until it knows which shift it needs next,
// all lanes evaluate together, then each lane files the result.  The sequence of
// evaluations per pixel -- and therefore Ncalls, d, the 4x4 block and every tie decision --
// is exactly the reference's.
// Eval: int operator()(int si, int sj, double &cost, FitArgs &args) -> error_status bits.
// Grid: anything indexable by [int] yielding double& (a local array, or a shared-memory column).
struct WalkCache {
    unsigned known;                 // bit 5 r + c: logical entry (r, c) holds an evaluated cost
    int b0, b1;                     // storage row / column of logical row / column 0
};

__device__ __forceinline__ int walk_cell(int b0, int b1, int r, int c)
{
    int pr = b0 + r, pc = b1 + c;
    pr = pr >= 5 ? pr - 5 : pr;
    pc = pc >= 5 ? pc - 5 : pc;
    return 5 * pr + pc;
}

// logical entry n (0..24) of the cache as the reference holds it: the cost, or -1 when not evaluated
template <class Grid>
__device__ __forceinline__ double walk_cache_get(Grid d, const WalkCache &wc, int n)
{
    const int r = (n * 13) >> 6;                   // n / 5 for n < 64
    return ((wc.known >> n) & 1u) ? d[walk_cell(wc.b0, wc.b1, r, n - 5 * r)] : -1.;
}

template <class Eval, class Grid>
__device__ inline int walk_minimise(Eval &eval, int subpx, const double *quad, FitArgs &args, double &out,
                                    double *uv, Grid d, double *a, int &ncalls, WalkCache &wc)
{
    enum { R_CENTRE, R_LO, R_HI, R_FILL };         // what the pending evaluation is for
    const double tol = 1e-8;                       // absolute, Optim.cpp:243
    constexpr unsigned COL0 = 0x108421u, COL4 = 0x1084210u, ALL = 0x1ffffffu;
    int settled0 = 0, settled1 = 0, axis = 0, st = UMPA_ST_OK;
    int ip = 0, jp = 0, idle = 0, b0 = 0, b1 = 0;
    bool fill = false, skip_limit = false, finished = false;
    unsigned known = 0;
    FitArgs keep = args;
    ncalls = 0;
    int c0 = (int)round(uv[0]), c1 = (int)round(uv[1]);
    int req = R_CENTRE, sr = 2, sc = 2;            // the pending evaluation: logical cell (sr, sc) = shift (c0 + sr - 2, c1 + sc - 2)
    double dc = 0.;                                // d[12], the cost at the current centre

    while (true) {
        // ---- the one evaluation site ----
        double v;
        const int se = eval(c0 + sr - 2, c1 + sc - 2, v, args);
        ncalls++;
        idle = 0;
        if (se != UMPA_ST_OK) { st = se; break; }  // bound error: return at once (Optim.cpp:264,291,324,359)

        // ---- file the result ----
        if (req == R_FILL && v < dc) {             // 4x4 fill, lower value off-axis: hard restart there (Optim.cpp:364-377)
            c0 += sr - 2; c1 += sc - 2;
            sr = sc = 2;
            known = 0;
            args = keep;
            settled0 = settled1 = 0;
            skip_limit = true;
            fill = false;
            req = R_CENTRE;                        // (keep is NOT refreshed: see above)
        } else if (req == R_CENTRE || (req == R_LO && !(v > dc + tol)) || (req == R_HI && !(v > dc - tol)))
            keep = args;                           // Optim.cpp:262, 294-296, 325-327
        if (sr == 2 && sc == 2) dc = v;
        d[walk_cell(b0, b1, sr, sc)] = v;
        known |= 1u << (5 * sr + sc);

        // ---- advance this lane until it needs the next cost value (or is done) ----
        bool done = false;
        while (true) {
            if (fill) {                            // next missing entry of the 4x4 block (row-major)
                const unsigned pending = (0x7bdefu << (5 * ip + jp)) & ~known;    // 4 rows of 4 bits, 5 apart
                if (!pending) { finished = true; done = true; break; }
                const int slot = __ffs(pending) - 1;
                sr = (slot * 13) >> 6;             // slot / 5 for slot < 64
                sc = slot - 5 * sr;
                req = R_FILL;
                break;
            }
            // head of the reference's loop (Optim.cpp:267); a restart jumps past the test (goto start)
            if (!skip_limit && ncalls >= UMPA_MAX_CALLS) { st = 0; done = true; break; }   // Optim.cpp:267,477
            // Not in the reference: with a NaN cost next to finite ones (a non-finite input pixel) its loop can step
            // back and forth between two evaluated shifts for ever -- MAX_CALLS only counts evaluations.  Finite
            // costs never revisit (a few visits here between two evaluations at most), so this changes no result;
            // it turns a hung GPU into a failed pixel (err = 0).
            if (++idle > 16) { st = 0; done = true; break; }
            skip_limit = false;
            // minus / plus neighbour along the axis: logical cells (2, 1) / (2, 3) or (1, 2) / (3, 2)
            const int lo = axis ? 7 : 11, hi = axis ? 17 : 13;
            if (!((known >> lo) & 1u)) { sr = 2 - axis; sc = 1 + axis; req = R_LO; break; }
            if (!((known >> hi) & 1u)) { sr = 2 + axis; sc = 3 - axis; req = R_HI; break; }
            const double dl = d[walk_cell(b0, b1, 2 - axis, 1 + axis)], dh = d[walk_cell(b0, b1, 2 + axis, 3 - axis)];
            const bool up_m = dl > dc + tol, up_p = dh > dc - tol;
            if (up_m && up_p) {                    // bracketed on this axis
                const int dir = dl < dh ? -1 : 1;
                if (axis) settled1 = dir; else settled0 = dir;
                if ((axis ? settled0 : settled1) == 0) { axis = 1 - axis; continue; }
                ip = d[walk_cell(b0, b1, 3, 2)] < d[walk_cell(b0, b1, 1, 2)] ? 1 : 0;
                jp = d[walk_cell(b0, b1, 2, 3)] < d[walk_cell(b0, b1, 2, 1)] ? 1 : 0;
                fill = true;
                continue;
            }
            uv[0] = c0; uv[1] = c1;                // best so far, Optim.cpp:421-423
            out = dc;
            bool plus = up_m;
            if (!up_p && !up_m) plus = dh < dl;    // local maximum: go downhill
            // one step along the axis (Optim.cpp:431-474): the cache moves with the centre
            if (plus) {
                dc = dh;
                if (axis) { c0 += 1; b0 = b0 == 4 ? 0 : b0 + 1; known >>= 5; }
                else { c1 += 1; b1 = b1 == 4 ? 0 : b1 + 1; known = (known >> 1) & ~COL4; }
            } else {
                dc = dl;
                if (axis) { c0 -= 1; b0 = b0 == 0 ? 4 : b0 - 1; known = (known << 5) & ALL; }
                else { c1 -= 1; b1 = b1 == 0 ? 4 : b1 - 1; known = (known << 1) & ~COL0 & ALL; }
            }
            if (axis) settled0 = 0; else settled1 = 0;
        }
        if (done) break;
    }
    wc.known = known; wc.b0 = b0; wc.b1 = b1;

    if (finished) {                                // minimum bracketed on both axes: sub-pixel fit
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) a[4 * r + q] = d[walk_cell(b0, b1, ip + r, jp + q)];
        args = keep;
        uv[0] = 1. - ip;
        uv[1] = 1. - jp;
        if (subpx == 0) out = uv[0];                                   // reference quirk, Optim.cpp:399
        else if (subpx == 1) out = subpixel_quadratic(quad, a, uv);
        else out = subpixel_spline(a, uv);
        uv[0] += c0 + ip - 1.;
        uv[1] += c1 + jp - 1.;
        st = UMPA_ST_OK;
    }
    return st;
}

// Writes one pixel's results the way Model*::min packs `values` (Model.cpp:573-576, 934-938).
template <class Grid>
__device__ __forceinline__ void store_pixel(const umpa_outputs &o, size_t n, int kind, int st, double f,
                                            const FitArgs &args, const double *uv, Grid d, const WalkCache &wc,
                                            const double *a, int ncalls, bool have_a)
{
    if (o.f) o.f[n] = f;
    if (o.T) o.T[n] = args.t;
    if (o.dx) o.dx[n] = uv[1];
    if (o.dy) o.dy[n] = uv[0];
    if (o.df && kind == UMPA_DF) o.df[n] = args.v;
    if (o.err) o.err[n] = (st & UMPA_ST_OK) ? 1 : 0;
    if (o.ncalls) o.ncalls[n] = ncalls;
    if (o.debug_d)
        for (int t = 0; t < 25; t++) o.debug_d[25 * n + t] = walk_cache_get(d, wc, t);
    if (o.debug_a) {
#pragma unroll
        for (int t = 0; t < 16; t++) o.debug_a[16 * n + t] = have_a ? a[t] : 0.;
    }
}
