// walk.cuh -- per-pixel integer-shift walk and sub-pixel refinement (device, FP64).
//
// Re-implements the behaviour of discrete_2d_minimizer (UMPA/lib/Optim.cpp:233-479),
// spmin (Optim.cpp:41-130) and spmin_quad (Optim.cpp:155-185) for one CUDA thread per
// pixel.  The cost function is a functor, so the same state machine runs on top of the
// lazy FP64 evaluator (lazy_path.cu) and of the FP32 shift tables (table_path.cu).
//
// Two steps, so that a kernel can store the debug arrays in between and the refinement can reuse
// the cache's storage:
//   walk_search()   the integer walk; leaves the 5x5 cost cache and where the 4x4 block lies
//   walk_refine()   the sub-pixel fit on that block (spline / quadratic / none)
#pragma once
#include "common.cuh"

struct FitArgs {            // what the reference keeps in CostArgs{NoDF,DF,DFKernel}: t (and v)
    double t, v;
};

// ---- the cost cache -----------------------------------------------------------------------
// The reference keeps d, a 5x5 cache of costs centred on the current integer shift (row-major, centre 12,
// -1 = not evaluated), and physically shifts it by one row / column on every step.  Here the cache is a
// RING: the 25 storage cells are addressed modulo 5 from an origin (b0, b1) that moves with the centre, and
// which cells hold an evaluated cost is a 25-bit register mask in the reference's (logical) order -- a step
// rotates the mask and moves the origin, nothing is copied and nothing has to be initialised.
// walk_cache_get() reads the cache the way the reference would have left it.
// Grid: anything indexable by [int] yielding double& (a local array, or a shared-memory column).
struct WalkState {
    unsigned known;                 // bit 5 r + c: logical entry (r, c) holds an evaluated cost
    int b0, b1;                     // storage row / column of logical row / column 0
    int c0, c1;                     // the integer shift the walk ended on (row, column)
    int ip, jp;                     // corner of the 4x4 block inside the 5x5 cache (Optim.cpp:344-345)
    bool finished;                  // minimum bracketed on both axes: the 4x4 block is complete
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
__device__ __forceinline__ double walk_cache_get(Grid d, const WalkState &ws, int n)
{
    const int r = (n * 13) >> 6;                   // n / 5 for n < 64
    return ((ws.known >> n) & 1u) ? d[walk_cell(ws.b0, ws.b1, r, n - 5 * r)] : -1.;
}

// entry (r, q) of the 4x4 block the sub-pixel fit works on (minimizer_debug.a, Optim.cpp:349-384)
template <class Grid>
__device__ __forceinline__ double walk_block_get(Grid d, const WalkState &ws, int r, int q)
{
    return d[walk_cell(ws.b0, ws.b1, ws.ip + r, ws.jp + q)];
}

// ---- cubic B-spline surface through the 4x4 block -----------------------------------
// f(x,y) = sum_ij B_i(x) B_j(y) a[4i+j] / 36 with x along rows; B_* are the four uniform
// cubic B-spline pieces on [0,1] for samples at -1,0,1,2; times 6 their coefficients are
//     t^0: {1, 4, 1, 0}   t^1: {-3, 0, 3, 0}   t^2: {3, -6, 3, 0}   t^3: {-1, 3, -3, 1}
// The 16 polynomial coefficients c[4m+n] (of x^n y^m) are built column by column of the block -- four block
// entries in registers at a time -- and then live in the cache's own storage cells 0..15 (the block is not needed
// any more; a kernel that reports the cache stores it BEFORE the refinement): the Newton iteration reads them back
// row by row, so the fit needs ~50 registers instead of the ~100 of block + coefficients in registers.
template <class Grid>
__device__ inline double subpixel_spline(Grid d, const WalkState &ws, double *pos)
{
    double c[16];
#pragma unroll
    for (int n = 0; n < 16; n++) c[n] = 0.;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double a0 = walk_block_get(d, ws, 0, j), a1 = walk_block_get(d, ws, 1, j);
        const double a2 = walk_block_get(d, ws, 2, j), a3 = walk_block_get(d, ws, 3, j);
        // x-powers of column sample j
        const double t[4] = {a0 + 4. * a1 + a2, 3. * (a2 - a0), 3. * (a0 + a2) - 6. * a1, (a3 - a0) + 3. * (a1 - a2)};
        const double B[4][4] = {{1., 4., 1., 0.}, {-3., 0., 3., 0.}, {3., -6., 3., 0.}, {-1., 3., -3., 1.}};
#pragma unroll
        for (int m = 0; m < 4; m++) {
            if (B[m][j] == 0.) continue;
#pragma unroll
            for (int n = 0; n < 4; n++) c[4 * m + n] += B[m][j] * t[n];
        }
    }
#pragma unroll
    for (int n = 0; n < 16; n++) d[n] = c[n];
    double x = pos[0], y = pos[1];
    for (int it = 0; it <= 20; it++) {
        double A[4], A1[4], A2[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const double c0 = d[4 * m], c1 = d[4 * m + 1], c2 = d[4 * m + 2], c3 = d[4 * m + 3];
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
        f += yp * (d[4 * m] + x * (d[4 * m + 1] + x * (d[4 * m + 2] + x * d[4 * m + 3])));
        yp *= y;
    }
    return f / 36.;
}

// ---- least-squares quadratic through the 4x4 block ------------------------------------
// p = 400 * pinv(A) a for the basis [1, i, j, i^2, ij, j^2] on i,j in {-1,0,1,2}
// (i along rows).  The 6x16 integer matrix is 400*(A^T A)^-1 A^T, generated on the host
// when a model is created (quad_matrix in capi.cu) and passed in as `quad` (device memory).
template <class Grid>
__device__ inline double subpixel_quadratic(const double *__restrict__ c_quad, Grid d, const WalkState &ws, double *pos)
{
    double p[6] = {0., 0., 0., 0., 0., 0.};
#pragma unroll
    for (int n = 0; n < 16; n++) {
        const double a = walk_block_get(d, ws, n >> 2, n & 3);
#pragma unroll
        for (int r = 0; r < 6; r++) p[r] += c_quad[16 * r + n] * a;
    }
    const double det = 4. * p[3] * p[5] - p[4] * p[4];
    // reference quirk kept: pos[0] gets the column solution, pos[1] the row solution
    pos[0] = -(2. * p[3] * p[2] - p[4] * p[1]) / det;
    pos[1] = -(2. * p[5] * p[1] - p[4] * p[2]) / det;
    return (p[0] + .5 * (p[2] * pos[0] + p[1] * pos[1])) / 400.;
}

// ---- the walk --------------------------------------------------------------------------
// axis 0 scans the column shift, axis 1 the row shift.  `keep` is the reference's args_copy: the fit
// parameters of the best shift seen so far -- deliberately NOT refreshed on a restart (Optim.cpp:364-377),
// which the reference's outputs depend on.
//
// The reference calls the cost function from four places (centre, minus neighbour, plus
// neighbour, 4x4 fill).  On a GPU that would serialise the lanes of a warp that happen to be
// at different call sites, so the same control flow is written as a state machine whose loop has ONE
// evaluation site: each lane advances its state until it knows which shift it needs next,
// all lanes evaluate together, then each lane files the result (the centre, which every lane evaluates
// first and exactly once, may sit in front of the loop: PEEL).  The sequence of
// evaluations per pixel -- and therefore Ncalls, d, the 4x4 block and every tie decision --
// is exactly the reference's.
// Eval: int operator()(int si, int sj, double &cost, FitArgs &args) -> error_status bits.
// Returns the error_status bits; on UMPA_ST_OK with ws.finished the caller runs walk_refine().
// PEEL: the centre evaluation gets its own copy of the cost function in front of the loop (the table evaluators: a few
// dozen instructions, and the loop's filing code no longer knows the centre); PEEL = false sends it through the loop's
// evaluation site (the lazy evaluator, whose code is long and whose bookkeeping does not matter).
template <bool PEEL = true, class Eval, class Grid>
__device__ inline int walk_search(Eval &eval, FitArgs &args, double &out, double *uv, Grid d, int &ncalls, WalkState &ws)
{
    enum { R_LO = 2, R_HI = 4, R_FILL = 8 };       // what the pending evaluation is for
    const double tol = 1e-8;                       // absolute, Optim.cpp:243
    constexpr unsigned COL0 = 0x108421u, COL4 = 0x1084210u, ALL = 0x1ffffffu;
    int settled0 = 0, settled1 = 0, axis = 0, st = UMPA_ST_OK;
    int ip = 0, jp = 0, idle = 0, b0 = 0, b1 = 0;
    bool fill = false, skip_limit = false, finished = false;
    unsigned known = 0;
    FitArgs keep = args;
    ncalls = 0;
    int c0 = (int)round(uv[0]), c1 = (int)round(uv[1]);
    int req = R_LO, sr = 2, sc = 2;                // the pending evaluation: logical cell (sr, sc) = shift (c0 + sr - 2, c1 + sc - 2)
    double dc = 0.;                                // d[12], the cost at the current centre
    bool done = false;
    bool centre = !PEEL;                           // (!PEEL) the loop's first evaluation is the centre's

    // The centre (Optim.cpp:262) is evaluated exactly once -- a restart takes its new centre's cost from the fill
    // evaluation that triggered it -- and by all lanes at the same time, so it sits before the loop: the loop's
    // filing code then only knows neighbours and fill entries.
    if (PEEL) {
        double v;
        const int se = eval(c0, c1, v, args);
        ncalls = 1;
        if (se != UMPA_ST_OK) { st = se; done = true; }      // bound error: return at once (Optim.cpp:264)
        else {
            keep = args;
            dc = v;
            d[walk_cell(b0, b1, 2, 2)] = v;
            known = 1u << 12;
        }
    }

    while (!done) {
        if (PEEL || !centre) {
            // ---- advance this lane until it needs the next cost value (or is done) ----
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

        // ---- the one evaluation site of the loop ----
        double v;
        const int se = eval(c0 + sr - 2, c1 + sc - 2, v, args);
        ncalls++;
        idle = 0;
        if (se != UMPA_ST_OK) { st = se; break; }  // bound error: return at once (Optim.cpp:291,324,359)

        // ---- file the result ----
        if (!PEEL && centre) {
            centre = false;
            keep = args;                           // Optim.cpp:262
            dc = v;
        } else if (req == R_FILL) {
            // 4x4 fill, lower value off-axis: hard restart at that shift (Optim.cpp:364-377) -- the walk continues as if
            // it had started there, except that args / keep are NOT refreshed (see above) and the loop-head test is
            // skipped.  Rare, so it sits behind a branch: the eleven fill evaluations of a pixel pay one compare for it.
            if (v < dc) {
                args = keep;
                c0 += sr - 2; c1 += sc - 2;
                sr = 2; sc = 2;
                known = 0u;
                settled0 = 0; settled1 = 0;
                skip_limit = true;
                fill = false;
                dc = v;
            }
        } else if (!(v > dc + (req == R_LO ? tol : -tol))) {
            // a minus / plus neighbour that is not higher than the centre (Optim.cpp:294-296, 325-327; the two tests
            // are not symmetric)
            keep = args;
        }
        d[walk_cell(b0, b1, sr, sc)] = v;
        known |= 1u << (5 * sr + sc);
    }
    ws.known = known; ws.b0 = b0; ws.b1 = b1; ws.c0 = c0; ws.c1 = c1; ws.ip = ip; ws.jp = jp;
    ws.finished = finished;
    if (finished) { args = keep; st = UMPA_ST_OK; }                    // Optim.cpp:386
    return st;
}

// The sub-pixel fit on the completed 4x4 block (Optim.cpp:398-408).  May overwrite the cache's storage.
template <class Grid>
__device__ inline void walk_refine(int subpx, const double *quad, Grid d, const WalkState &ws, double &out, double *uv)
{
    uv[0] = 1. - ws.ip;
    uv[1] = 1. - ws.jp;
    if (subpx == 0) out = uv[0];                                       // reference quirk, Optim.cpp:399
    else if (subpx == 1) out = subpixel_quadratic(quad, d, ws, uv);
    else out = subpixel_spline(d, ws, uv);
    uv[0] += ws.c0 + ws.ip - 1.;
    uv[1] += ws.c1 + ws.jp - 1.;
}

// Writes one pixel's debug arrays (minimizer_debug.d / .a, model.pyx:488-491): before walk_refine().
template <class Grid>
__device__ __forceinline__ void store_debug(const umpa_outputs &o, size_t n, Grid d, const WalkState &ws)
{
    if (o.debug_d)
        for (int t = 0; t < 25; t++) o.debug_d[25 * n + t] = walk_cache_get(d, ws, t);
    if (o.debug_a)
        for (int t = 0; t < 16; t++) o.debug_a[16 * n + t] = ws.finished ? walk_block_get(d, ws, t >> 2, t & 3) : 0.;
}

// Writes one pixel's results the way Model*::min packs `values` (Model.cpp:573-576, 934-938).
__device__ __forceinline__ void store_pixel(const umpa_outputs &o, size_t n, int kind, int st, double f,
                                            const FitArgs &args, const double *uv, int ncalls)
{
    if (o.f) o.f[n] = f;
    if (o.T) o.T[n] = args.t;
    if (o.dx) o.dx[n] = uv[1];
    if (o.dy) o.dy[n] = uv[0];
    if (o.df && kind == UMPA_DF) o.df[n] = args.v;
    if (o.err) o.err[n] = (st & UMPA_ST_OK) ? 1 : 0;
    if (o.ncalls) o.ncalls[n] = ncalls;
}
