/*
 * umpa_oracle.c -- CPU restatement of the UMPA++ per-pixel window-matching path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in umpa_b200/csrc.  It is never linked into, imported by or called from the
 * product; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * --impl reference legs may load it.
 *
 * Parity pinning: the reference's own tests hold no golden vectors for this
 * path, so this restatement is pinned against
 *   (1) the UNMODIFIED reference compiled here into oracle/_ref (oracle/build.py)
 *       -- tests/test_oracle_vs_reference.py, and
 *   (2) golden vectors generated from that compiled reference and committed
 *       under tests/golden/ (tests/golden/make_golden.py).
 *
 * Everything is plain double arithmetic in the reference's summation order.
 * Citations are path:line under /root/reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define UO_KWS 8                       /* KERNEL_WINDOW_SIZE, UMPA/lib/Model.h:7 */
#define UO_KSIDE (2 * UO_KWS + 1)
#define UO_MAX_CALLS 500               /* UMPA/lib/Optim.cpp:14 */

enum { UO_NODF = 0, UO_DF = 1, UO_DFKERNEL = 2 };

/* status bits, same fields as error_status (UMPA/lib/Optim.h:7-12) */
enum { UO_OK = 1, UO_BOUND = 2, UO_DIM = 4, UO_POS = 8 };

typedef struct {
    int kind, Na, Nw, max_shift, padding;
    int subpx_func;        /* -1 spline (default), 0 none, 1 quadratic; Model.cpp:214 */
    int reference_shift;   /* 0: ref window moves (default); 1: sample window moves; Model.cpp:215 */
    const double **sam, **ref, **mask;   /* mask == NULL when no masks were given */
    int *dim;              /* Na x 2 (rows, cols) */
    int *pos;              /* Na x 2 */
    double *win;           /* (2Nw+1)^2 */
} uo_model;

/* what the reference keeps in CostArgs* (UMPA/lib/Model.h:16-71) */
typedef struct {
    double t, v;
    const double *kernel;  /* DFKernel only, 17x17 normalised */
} uo_args;

/* ------------------------------------------------------------------ model */

uo_model *uo_create(int kind, int Na, const int *dim, const int *pos,
                    const double **sam, const double **ref, const double **mask,
                    int Nw, const double *win, int max_shift, int padding)
{
    uo_model *m = (uo_model *)calloc(1, sizeof(uo_model));
    int K = 2 * Nw + 1;
    m->kind = kind; m->Na = Na; m->Nw = Nw; m->max_shift = max_shift; m->padding = padding;
    m->subpx_func = -1; m->reference_shift = 0;
    m->sam = (const double **)malloc(sizeof(double *) * Na);
    m->ref = (const double **)malloc(sizeof(double *) * Na);
    m->mask = mask ? (const double **)malloc(sizeof(double *) * Na) : NULL;
    m->dim = (int *)malloc(sizeof(int) * 2 * Na);
    m->pos = (int *)malloc(sizeof(int) * 2 * Na);
    m->win = (double *)malloc(sizeof(double) * K * K);
    for (int k = 0; k < Na; k++) {
        m->sam[k] = sam[k]; m->ref[k] = ref[k];
        if (mask) m->mask[k] = mask[k];
        m->dim[2 * k] = dim[2 * k]; m->dim[2 * k + 1] = dim[2 * k + 1];
        m->pos[2 * k] = pos[2 * k]; m->pos[2 * k + 1] = pos[2 * k + 1];
    }
    memcpy(m->win, win, sizeof(double) * K * K);
    return m;
}

void uo_destroy(uo_model *m)
{
    if (!m) return;
    free((void *)m->sam); free((void *)m->ref); free((void *)m->mask);
    free(m->dim); free(m->pos); free(m->win); free(m);
}

void uo_set_window(uo_model *m, int Nw, const double *win)   /* Model.cpp:239-246 */
{
    int K = 2 * Nw + 1;
    free(m->win);
    m->win = (double *)malloc(sizeof(double) * K * K);
    memcpy(m->win, win, sizeof(double) * K * K);
    m->Nw = Nw;
}

void uo_set_options(uo_model *m, int subpx_func, int reference_shift)
{
    m->subpx_func = subpx_func; m->reference_shift = reference_shift;
}

/* frame k contributes to pixel (i,j)?  Model.cpp:285-288 / 430-433 / 716-719 */
static int frame_in_reach(const uo_model *m, int k, int i, int j)
{
    int ri = i - m->pos[2 * k], rj = j - m->pos[2 * k + 1];
    if (ri - m->padding < 0) return 0;
    if (ri + m->padding > m->dim[2 * k]) return 0;
    if (rj - m->padding < 0) return 0;
    if (rj + m->padding > m->dim[2 * k + 1]) return 0;
    return 1;
}

/* Utils.cpp:125-130 */
static inline double mix_weights(double a, double b) { return a * b / (a + b + 1e-8); }

/* coverage of one pixel, Model.cpp:273-314 */
double uo_coverage(const uo_model *m, int i, int j)
{
    double wt = 0.;
    for (int k = 0; k < m->Na; k++) {
        if (!frame_in_reach(m, k, i, j)) continue;
        if (!m->mask) wt += 1.;
        else wt += m->mask[k][(i - m->pos[2 * k]) * m->dim[2 * k + 1] + (j - m->pos[2 * k + 1])];
    }
    return wt;
}

/* normalised 17x17 blur kernel exp(-a i^2 - b i j - c j^2), Model.cpp:88-117, Utils.cpp:46-50 */
void uo_make_kernel(double a, double b, double c, double *kernel)
{
    double norm = 0.;
    for (int r = 0; r < UO_KSIDE; r++)
        for (int q = 0; q < UO_KSIDE; q++) {
            int i = r - UO_KWS, j = q - UO_KWS;
            double v = exp(-a * i * i - b * i * j - c * j * j);
            kernel[r * UO_KSIDE + q] = v;
            norm += v;
        }
    for (int n = 0; n < UO_KSIDE * UO_KSIDE; n++) kernel[n] /= norm;
}

/* Utils.cpp:85-97 */
static double blur_at(const double *img, int i, int j, int W, const double *kernel)
{
    double out = 0.;
    for (int r = -UO_KWS; r <= UO_KWS; r++)
        for (int q = -UO_KWS; q <= UO_KWS; q++)
            out += kernel[UO_KSIDE * (r + UO_KWS) + q + UO_KWS] * img[(i + r) * W + (j + q)];
    return out;
}

/* Utils.cpp:103-117 */
static double weighted_blur_at(const double *img, const double *wgt, int i, int j, int W,
                               const double *kernel)
{
    double out = 0., w = 0.;
    for (int r = -UO_KWS; r <= UO_KWS; r++)
        for (int q = -UO_KWS; q <= UO_KWS; q++) {
            double kv = kernel[UO_KSIDE * (r + UO_KWS) + q + UO_KWS];
            out += kv * img[(i + r) * W + (j + q)] * wgt[(i + r) * W + (j + q)];
            w += kv * wgt[(i + r) * W + (j + q)];
        }
    return out / w;
}

/* -------------------------------------------------------------- cost()s */

/* shift range test shared by all three models, Model.cpp:372-399 / 654-681 / 1011-1038 */
static int shift_status(const uo_model *m, int si, int sj)
{
    if (si <= -m->max_shift) return UO_BOUND;
    if (si >= m->max_shift) return UO_BOUND;
    if (sj <= -m->max_shift) return UO_BOUND | UO_DIM;
    if (sj >= m->max_shift) return UO_BOUND | UO_DIM | UO_POS;
    return UO_OK;
}

/* One cost evaluation at pixel (i,j) (raw frame coordinates) and integer shift
 * (si,sj).  On success writes *out and args->t (and args->v for DF).
 * NoDF: Model.cpp:359-509, DF: 631-862, DFKernel: 997-1151. */
static int cost_eval(const uo_model *m, int i, int j, int si, int sj, uo_args *args, double *out)
{
    int st = shift_status(m, si, sj);
    if (st != UO_OK) return st;

    const int Nw = m->Nw, K = 2 * Nw + 1;
    /* (ri,rj): centre of the reference window, (qi,qj): centre of the sample window */
    int ri, rj, qi, qj;
    if (m->reference_shift) { ri = i; rj = j; qi = i - si; qj = j - sj; }
    else                    { ri = i + si; rj = j + sj; qi = i; qj = j; }

    const int masked = (m->mask != NULL);
    double t1 = 0., t2 = 0., t3 = 0., t4 = 0., t5 = 0., t6 = 0.;
    double wt = masked ? 0. : (double)m->Na;

    for (int k = 0; k < m->Na; k++) {
        if (!frame_in_reach(m, k, i, j)) continue;
        const int W = m->dim[2 * k + 1];
        const int pi = m->pos[2 * k], pj = m->pos[2 * k + 1];
        const double *R = m->ref[k], *S = m->sam[k];
        const double *M = masked ? m->mask[k] : NULL;

        if (m->kind == UO_DF) {
            /* window-weighted mean of the reference frame, Model.cpp:723-739 / 789-808 */
            double mean = 0., den = 0.;
            for (int a = 0; a < K; a++)
                for (int b = 0; b < K; b++) {
                    double w = m->win[a * K + b];
                    mean += w * R[(a - Nw + ri - pi) * W + (b - Nw + rj - pj)];
                    den += w;
                }
            mean /= den;
            double s2 = 0., s4 = 0., s6 = 0.;
            for (int a = 0; a < K; a++)
                for (int b = 0; b < K; b++) {
                    int nr = (a - Nw + ri - pi) * W + (b - Nw + rj - pj);
                    int ns = (a - Nw + qi - pi) * W + (b - Nw + qj - pj);
                    double w = m->win[a * K + b], s = S[ns], r = R[nr];
                    if (!masked) {                       /* Model.cpp:759-767 */
                        t1 += w * s * s;
                        t3 += w * r * r;
                        s4 += w * s;
                        t5 += w * r * s;
                        s6 += w * r;
                    } else {                             /* Model.cpp:827-840 */
                        double g = mix_weights(M[nr], M[ns]);
                        t1 += g * w * s * s;
                        s2 += g * w;
                        t3 += g * w * r * r;
                        s4 += g * w * s;
                        t5 += g * w * r * s;
                        s6 += g * w * r;
                        wt += g * w;
                    }
                }
            t2 += masked ? mean * mean * s2 : mean * mean;   /* Model.cpp:770 / 843 */
            t4 += mean * s4;
            t6 += mean * s6;
        } else {
            for (int a = 0; a < K; a++)
                for (int b = 0; b < K; b++) {
                    int fi = a - Nw + ri - pi, fj = b - Nw + rj - pj;
                    int nr = fi * W + fj;
                    int ns = (a - Nw + qi - pi) * W + (b - Nw + qj - pj);
                    double w = m->win[a * K + b], s = S[ns], r;
                    if (m->kind == UO_DFKERNEL)
                        r = masked ? weighted_blur_at(R, M, fi, fj, W, args->kernel)
                                   : blur_at(R, fi, fj, W, args->kernel);
                    else
                        r = R[nr];
                    if (!masked) {                       /* Model.cpp:454-456 / 1093-1095 */
                        t1 += w * s * s;
                        t3 += w * r * r;
                        t5 += w * r * s;
                    } else {                             /* Model.cpp:488-495 / 1127-1138 */
                        double g = mix_weights(M[nr], M[ns]);
                        t1 += g * w * s * s;
                        t3 += g * w * r * r;
                        t5 += g * w * r * s;
                        wt += g * w;
                    }
                }
        }
    }

    if (m->kind == UO_DF) {                              /* Model.cpp:849-858 */
        double den = t2 * t3 - t6 * t6;
        double Kc = (t2 * t5 - t4 * t6) / den;
        double beta = (t3 * t4 - t5 * t6) / den;
        args->t = beta + Kc;
        args->v = Kc / args->t;
        *out = (t1 + beta * beta * t2 + Kc * Kc * t3 - 2 * beta * t4 - 2 * Kc * t5 + 2 * beta * Kc * t6) / wt;
    } else {                                             /* Model.cpp:502-505 / 1144-1147 */
        args->t = t5 / t3;
        *out = (t1 - t5 * args->t) / wt;
    }
    return UO_OK;
}

/* cost_interface, Model.cpp:533-542 / 887-897 / 1181-1192.
 * values: [cost, t, v]; abc only read for DFKernel. returns status bits. */
int uo_cost(const uo_model *m, int i, int j, int si, int sj, const double *abc, double *values)
{
    double kernel[UO_KSIDE * UO_KSIDE];
    uo_args args = {0., 0., NULL};
    if (m->kind == UO_DFKERNEL) { uo_make_kernel(abc[0], abc[1], abc[2], kernel); args.kernel = kernel; }
    int st = cost_eval(m, i, j, si, sj, &args, &values[0]);
    values[1] = args.t; values[2] = args.v;
    return st;
}

/* --------------------------------------------------- sub-pixel refinement */

/* Power-basis coefficients (times 6) of the uniform cubic B-spline pieces on
 * [0,1] for the four samples at -1,0,1,2.  BSP[n][s] multiplies t^n of sample s. */
static const double BSP[4][4] = {
    { 1.,  4.,  1., 0.},
    {-3.,  0.,  3., 0.},
    { 3., -6.,  3., 0.},
    {-1.,  3., -3., 1.}};

/* spmin, Optim.cpp:41-130: the 4x4 block defines the tensor cubic-B-spline
 * surface  f(x,y) = sum_ij B_i(x) B_j(y) a[4i+j] / 36  (x along rows);
 * Newton iterations on its gradient from pos, at most 21, stop when the
 * squared step is below 1e-8; returns f at the final position. */
double uo_spmin(const double *a, double *pos)
{
    double c[16];                  /* c[4m+n] multiplies x^n y^m */
    for (int mm = 0; mm < 4; mm++)
        for (int n = 0; n < 4; n++) {
            double acc = 0.;
            for (int i = 0; i < 4; i++)
                for (int j = 0; j < 4; j++)
                    acc += BSP[n][i] * BSP[mm][j] * a[4 * i + j];
            c[4 * mm + n] = acc;
        }
    double x = pos[0], y = pos[1];
    const double tol = 1e-8;
    for (int it = 0; it <= 20; it++) {
        double xp[4] = {1., x, x * x, x * x * x}, yp[4] = {1., y, y * y, y * y * y};
        double fx = 0., fy = 0., fxx = 0., fxy = 0., fyy = 0.;
        for (int mm = 0; mm < 4; mm++)
            for (int n = 0; n < 4; n++) {
                double cc = c[4 * mm + n];
                if (n >= 1) fx += n * cc * xp[n - 1] * yp[mm];
                if (mm >= 1) fy += mm * cc * xp[n] * yp[mm - 1];
                if (n >= 2) fxx += n * (n - 1) * cc * xp[n - 2] * yp[mm];
                if (n >= 1 && mm >= 1) fxy += n * mm * cc * xp[n - 1] * yp[mm - 1];
                if (mm >= 2) fyy += mm * (mm - 1) * cc * xp[n] * yp[mm - 2];
            }
        double det = fxx * fyy - fxy * fxy;
        double dx = (fxy * fy - fyy * fx) / det;
        double dy = (fxy * fx - fxx * fy) / det;
        x += dx; y += dy;
        if (dx * dx + dy * dy < tol) break;
    }
    pos[0] = x; pos[1] = y;
    double xp[4] = {1., x, x * x, x * x * x}, yp[4] = {1., y, y * y, y * y * y};
    double f = 0.;
    for (int mm = 0; mm < 4; mm++)
        for (int n = 0; n < 4; n++) f += c[4 * mm + n] * xp[n] * yp[mm];
    return f / 36.;
}

/* spmin_quad, Optim.cpp:155-185: least-squares quadratic
 * p0 + p1 i + p2 j + p3 i^2 + p4 i j + p5 j^2 over the grid i,j in {-1,0,1,2}
 * (i along rows), all coefficients scaled by 400 (integers in the reference).
 * Quirk kept: pos[0] receives the COLUMN solution and pos[1] the ROW one. */
static double QUAD[6][16];
static int quad_ready = 0;

static void quad_init(void)
{
    double A[16][6], N[6][12];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double u = i - 1., v = j - 1.;
            double row[6] = {1., u, v, u * u, u * v, v * v};
            memcpy(A[4 * i + j], row, sizeof(row));
        }
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) {
            double s = 0.;
            for (int n = 0; n < 16; n++) s += A[n][r] * A[n][c];
            N[r][c] = s; N[r][6 + c] = (r == c);
        }
    for (int p = 0; p < 6; p++) {            /* Gauss-Jordan with partial pivoting */
        int best = p;
        for (int r = p + 1; r < 6; r++) if (fabs(N[r][p]) > fabs(N[best][p])) best = r;
        if (best != p) for (int c = 0; c < 12; c++) { double t = N[p][c]; N[p][c] = N[best][c]; N[best][c] = t; }
        double d = N[p][p];
        for (int c = 0; c < 12; c++) N[p][c] /= d;
        for (int r = 0; r < 6; r++) if (r != p) {
            double f = N[r][p];
            for (int c = 0; c < 12; c++) N[r][c] -= f * N[p][c];
        }
    }
    for (int r = 0; r < 6; r++)
        for (int n = 0; n < 16; n++) {
            double s = 0.;
            for (int c = 0; c < 6; c++) s += N[r][6 + c] * A[n][c];
            QUAD[r][n] = round(400. * s);    /* exact integers, see Optim.cpp:169-174 */
        }
    quad_ready = 1;
}

double uo_spmin_quad(const double *a, double *pos)
{
    if (!quad_ready) quad_init();
    double p[6];
    for (int r = 0; r < 6; r++) {
        double s = 0.;
        for (int n = 0; n < 16; n++) s += QUAD[r][n] * a[n];
        p[r] = s;
    }
    double det = 4 * p[3] * p[5] - p[4] * p[4];
    pos[0] = -(2 * p[3] * p[2] - p[4] * p[1]) / det;
    pos[1] = -(2 * p[5] * p[1] - p[4] * p[2]) / det;
    return (p[0] + .5 * (p[2] * pos[0] + p[1] * pos[1])) / 400.;
}

/* ------------------------------------------------------- integer walk */

typedef struct { double d[25], a[16]; int ncalls; } uo_debug;

static void grid_clear(double *d) { for (int n = 0; n < 25; n++) d[n] = -1.; }

/* discrete_2d_minimizer, Optim.cpp:233-479 (state machine described in
 * SURVEY.md 3.3).  d is a 5x5 cache of costs centred on the current integer
 * shift c=(c0,c1) (row, col); axis 0 scans columns, axis 1 scans rows. */
static int walk_minimise(const uo_model *m, int i, int j, uo_args *args, double *out, double *uv,
                         uo_debug *db)
{
    double *d = db->d, *a = db->a;
    const double tol = 1e-8;
    int settled[2] = {0, 0};
    int axis = 0, st;
    int c[2];
    uo_args keep;

    grid_clear(d);
    db->ncalls = 0;
    c[0] = (int)round(uv[0]);
    c[1] = (int)round(uv[1]);

    st = cost_eval(m, i, j, c[0], c[1], args, &d[12]);
    db->ncalls++;
    if (st != UO_OK) return st;
    keep = *args;

    while (db->ncalls < UO_MAX_CALLS) {
rescan:;
        /* neighbour on the minus / plus side along the current axis */
        const int lo = axis ? 7 : 11, hi = axis ? 17 : 13;
        const int dr = axis ? 1 : 0, dc = axis ? 0 : 1;
        int up_m, up_p;

        if (d[lo] < -.5) {
            st = cost_eval(m, i, j, c[0] - dr, c[1] - dc, args, &d[lo]);
            db->ncalls++;
            if (st != UO_OK) return st;
            up_m = d[lo] > d[12] + tol;
            if (!up_m) keep = *args;
        } else up_m = d[lo] > d[12] + tol;

        if (d[hi] < -.5) {
            st = cost_eval(m, i, j, c[0] + dr, c[1] + dc, args, &d[hi]);
            db->ncalls++;
            if (st != UO_OK) return st;
            up_p = d[hi] > d[12] - tol;
            if (!up_p) keep = *args;
        } else up_p = d[hi] > d[12] - tol;

        if (up_m & up_p) {
            settled[axis] = d[lo] < d[hi] ? -1 : 1;
            if (settled[1 - axis] == 0) { axis = 1 - axis; continue; }

            /* minimum along both axes: collect the 4x4 block around it */
            const int ip = d[17] < d[7] ? 1 : 0;
            const int jp = d[13] < d[11] ? 1 : 0;
            for (int r = 0; r < 4; r++)
                for (int q = 0; q < 4; q++) {
                    const int n = 5 * (ip + r) + jp + q;
                    if (d[n] < -.9) {
                        const int e0 = c[0] + ip + r - 2, e1 = c[1] + jp + q - 2;
                        st = cost_eval(m, i, j, e0, e1, args, &a[4 * r + q]);
                        db->ncalls++;
                        if (st != UO_OK) return st;
                        d[n] = a[4 * r + q];
                        if (a[4 * r + q] < d[12]) {
                            /* a lower value off-axis: restart from there */
                            const double v = a[4 * r + q];
                            c[0] = e0; c[1] = e1;
                            grid_clear(d);
                            d[12] = v;
                            *args = keep;
                            settled[0] = settled[1] = 0;
                            goto rescan;
                        }
                    } else a[4 * r + q] = d[n];
                }
            *args = keep;
            uv[0] = 1. - ip;
            uv[1] = 1. - jp;
            if (m->subpx_func == 0) *out = uv[0];
            else if (m->subpx_func == 1) *out = uo_spmin_quad(a, uv);
            else *out = uo_spmin(a, uv);
            uv[0] += c[0] + ip - 1.;
            uv[1] += c[1] + jp - 1.;
            return st;
        }

        /* best so far, Optim.cpp:421-423 */
        uv[0] = c[0]; uv[1] = c[1];
        *out = d[12];

        if (!up_p && !up_m) up_m = d[hi] < d[lo];       /* local maximum: go downhill */

        if (up_m) {                                     /* step towards plus */
            c[1 - axis] += 1;
            if (axis) { memmove(d, d + 5, 20 * sizeof(double)); for (int n = 20; n < 25; n++) d[n] = -1.; }
            else      { memmove(d, d + 1, 24 * sizeof(double)); for (int r = 0; r < 5; r++) d[5 * r + 4] = -1.; }
        } else {                                        /* step towards minus */
            c[1 - axis] -= 1;
            if (axis) { memmove(d + 5, d, 20 * sizeof(double)); for (int n = 0; n < 5; n++) d[n] = -1.; }
            else      { memmove(d + 1, d, 24 * sizeof(double)); for (int r = 0; r < 5; r++) d[5 * r] = -1.; }
        }
        settled[1 - axis] = 0;
    }
    return 0;   /* too many calls, Optim.cpp:477 */
}

/* Model*::min, Model.cpp:562-578 / 923-940 / 1222-1238.
 * values: Nparam doubles (4 NoDF, 5 DF, 7 DFKernel with abc in [4:7] on input).
 * returns error.ok */
int uo_min(const uo_model *m, int i, int j, double *values, double *uv,
           double *dbg_d, double *dbg_a, int *ncalls)
{
    double kernel[UO_KSIDE * UO_KSIDE];
    uo_args args = {0., 0., NULL};
    uo_debug db;
    double D = 0.;      /* uninitialised in the reference (Model.cpp:566) */
    memset(db.a, 0, sizeof(db.a));
    if (m->kind == UO_DFKERNEL) { uo_make_kernel(values[4], values[5], values[6], kernel); args.kernel = kernel; }
    int st = walk_minimise(m, i, j, &args, &D, uv, &db);
    values[0] = D;
    values[1] = args.t;
    values[2] = uv[1];
    values[3] = uv[0];
    if (m->kind == UO_DF) values[4] = args.v;
    if (dbg_d) memcpy(dbg_d, db.d, sizeof(db.d));
    if (dbg_a) memcpy(dbg_a, db.a, sizeof(db.a));
    if (ncalls) *ncalls = db.ncalls;
    return (st & UO_OK) ? 1 : 0;
}

/* pixel loop of UMPAModelBase._match, model.pyx:476-492.  offs* = padding+start.
 * cover may be NULL (no gating). dbg_* may be NULL. */
void uo_match(const uo_model *m, int offs0, int step0, int N0, int offs1, int step1, int N1,
              const double *cover, double cover_threshold, int nparam,
              double *values, double *uv, int *err,
              double *dbg_d, double *dbg_a, int *ncalls, int nthreads)
{
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#endif
    #pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (int xi = 0; xi < N0; xi++)
        for (int xj = 0; xj < N1; xj++) {
            size_t n = (size_t)xi * N1 + xj;
            if (cover && cover[n] < cover_threshold) continue;
            err[n] = uo_min(m, offs0 + step0 * xi, offs1 + step1 * xj, values + n * nparam, uv + 2 * n,
                            dbg_d ? dbg_d + 25 * n : NULL, dbg_a ? dbg_a + 16 * n : NULL,
                            ncalls ? ncalls + n : NULL);
        }
}

int uo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
