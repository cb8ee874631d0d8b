// table_path.cu -- the fast CUDA path for UMPAModelNoDF / UMPAModelDF .match().
//
// What the reference computes per pixel p and integer shift s (UMPA/lib/Model.cpp:359-509,
// 631-862; reference window moves: ia = i + shift, Model.cpp:415-421 / 695-701):
//     t1 = sum_k sum_u w(u) S_k(p+u)^2              t3 = sum_k sum_u w(u) R_k(p+s+u)^2
//     t5 = sum_k sum_u w(u) R_k(p+s+u) S_k(p+u)
//  DF: m_k = sum_u w R_k(p+s+u) / sum w,  t2 = sum_k m_k^2,  t6 = sum_k m_k sum_u w R_k(p+s+u),
//      t4 = sum_k m_k sum_u w S_k(p+u)
// and then a closed-form solve.  The reference evaluates this ~17 times per pixel with
// two (2Nw+1)^2 x Na gather loops each.  Here the same numbers come from an algebraic
// regrouping that removes the (2Nw+1)^2 factor from the shift-dependent work:
//   * the window sum is linear, so  t5(p,s) = [ w (*) C_s ](p)  with the UNWINDOWED
//     frame correlation  C_s(q) = sum_k R_k(q+s) S_k(q)  -- Na FMAs per (q,s) -- followed by
//     ONE separable Hamming filter per shift (2(2Nw+1) FMAs per (p,s));
//   * t4(p,s) = sum_k a_k(p+s) b_k(p) / sum w  with the per-frame filtered images
//     a_k = w (*) R_k, b_k = w (*) S_k  -- again Na FMAs per (p,s);
//   * t1 depends on p only, t2/t3/t6 on p+s only: they are images, computed once.
// Frames are stored mean-centred in FP32 (x' = x - mean_k, centring done in FP64); the
// uncentred sums are rebuilt in FP64 from the centred ones plus four cheap cross images,
// so FP32 cancellation scales with the speckle variance instead of the mean squared.
// The per-pixel solve, the reference's integer walk (Optim.cpp:233-479) and the spline
// refinement run in FP64 on those tables (walk.cuh).
//
// Kernels (all sm_100a CUDA-core kernels; the path is a stencil/correlation, no GEMM):
//   frame_partial_sums / finish_means / center_frames   upload-time conversion
//   moments_kernel     a_k, b_k stacks + aux images                (HBM bound)
//   shift_table_kernel C_s + filter -> cross table;  a_k,b_k -> mean table (FP32 FMA / smem bound)
//   table_walk_kernel  FP64 solve + walk + spline per pixel
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "walk.cuh"

namespace {

constexpr int TILE_W = 32;         // output tile width of shift_table_kernel (floats: one 128 B line)
constexpr int MAX_NT = 384;        // thread-block size cap of shift_table_kernel
constexpr int SMEM_CAP = 227 * 1024;

// ------------------------------------------------------------------ frame conversion

constexpr int SUM_BLOCKS = 64;

// grid (SUM_BLOCKS, 2*Na): partial sums of one FP64 frame
__global__ void frame_partial_sums(const double *sam, const double *ref, size_t frame_elems, int Na,
                                   double *partials)
{
    const int f = blockIdx.y;
    const double *src = (f < Na ? sam + (size_t)f * frame_elems : ref + (size_t)(f - Na) * frame_elems);
    const size_t chunk = (frame_elems + SUM_BLOCKS - 1) / SUM_BLOCKS;
    const size_t lo = (size_t)blockIdx.x * chunk, hi = min(frame_elems, lo + chunk);
    double s = 0.;
    for (size_t n = lo + threadIdx.x; n < hi; n += blockDim.x) s += src[n];
    __shared__ double red[256];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partials[(size_t)f * SUM_BLOCKS + blockIdx.x] = red[0];
}

__global__ void finish_means(const double *partials, int nframes, double inv_count, double *means64, float *mean_s,
                             float *mean_r, int Na)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    double s = 0.;
    for (int b = 0; b < SUM_BLOCKS; b++) s += partials[(size_t)f * SUM_BLOCKS + b];
    const double mu = s * inv_count;
    means64[f] = mu;
    if (f < Na) mean_s[f] = (float)mu; else mean_r[f - Na] = (float)mu;
}

// grid (ceil(W/256), H, 2*Na): x' = (float)(x - mean)
__global__ void center_frames(const double *sam, const double *ref, const double *means64, int Na, int H, int W,
                              int pitch, float *sam32, float *ref32)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= pitch) return;
    const bool is_s = f < Na;
    const int k = is_s ? f : f - Na;
    const double *src = (is_s ? sam : ref) + ((size_t)k * H + y) * W;
    float *dst = (is_s ? sam32 : ref32) + ((size_t)k * H + y) * pitch;
    dst[x] = x < W ? (float)(src[x] - means64[f]) : 0.f;
}

// ------------------------------------------------------------------ moments

struct MomentsParams {
    const float *sam, *ref;      // centred stacks [Na][H][pitch]
    float *fa, *fb;              // filtered stacks a_k (from ref), b_k (from sam); nullptr for NoDF
    float4 *auxS, *auxR;         // [H][pitch]
    const float *g;              // 1-D window factor, K
    const float *mean_s, *mean_r;
    int Na, Nw, H, W, pitch;
    int ty0, tx0;                // first tile (in tile units) of the bounding box
};

constexpr int MO_TH = 16, MO_TW = 64, MO_NT = 256;

// One block filters a MO_TH x MO_TW tile of every frame of both stacks.
// smem: inR, inS [EH][EW]; tmpR, tmpS [EH][MO_TW]; qR, qS [EH][EW]
__global__ void __launch_bounds__(MO_NT) moments_kernel(MomentsParams p)
{
    extern __shared__ float sm[];
    const int Nw = p.Nw, K = 2 * Nw + 1;
    const int EH = MO_TH + 2 * Nw, EW = MO_TW + 2 * Nw;
    float *inR = sm, *inS = inR + EH * EW;
    float *tmpR = inS + EH * EW, *tmpS = tmpR + EH * MO_TW;
    float *qR = tmpS + EH * MO_TW, *qS = qR + EH * EW;
    __shared__ float gs[UMPA_MAX_K];
    const int tid = threadIdx.x;
    if (tid < K) gs[tid] = p.g[tid];
    const int y0 = (blockIdx.y + p.ty0) * MO_TH, x0 = (blockIdx.x + p.tx0) * MO_TW;
    for (int n = tid; n < EH * EW; n += MO_NT) { qR[n] = 0.f; qS[n] = 0.f; }

    // each thread owns MO_TH*MO_TW/MO_NT = 4 output pixels: (oy[t], ox) with ox = tid % 64
    const int ox = tid % MO_TW, oyb = tid / MO_TW;      // rows oyb, oyb+4, oyb+8, oyb+12
    float m2[4] = {0, 0, 0, 0}, p3[4] = {0, 0, 0, 0}, uu[4] = {0, 0, 0, 0}, p1[4] = {0, 0, 0, 0}, vv[4] = {0, 0, 0, 0};
    const size_t fstride = (size_t)p.H * p.pitch;

    for (int k = 0; k < p.Na; k++) {
        const float *R = p.ref + k * fstride, *S = p.sam + k * fstride;
        __syncthreads();                                  // previous frame's passes are done with in*/tmp*
        for (int n = tid; n < EH * EW; n += MO_NT) {
            const int r = n / EW, c = n - r * EW;
            const int y = y0 - Nw + r, x = x0 - Nw + c;
            float rv = 0.f, sv = 0.f;
            if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                rv = __ldg(R + (size_t)y * p.pitch + x);
                sv = __ldg(S + (size_t)y * p.pitch + x);
            }
            inR[n] = rv; inS[n] = sv;
            qR[n] += rv * rv; qS[n] += sv * sv;
        }
        __syncthreads();
        for (int n = tid; n < EH * MO_TW; n += MO_NT) {    // row pass
            const int r = n / MO_TW, c = n - r * MO_TW;
            float ar = 0.f, as = 0.f;
            for (int v = 0; v < K; v++) {
                ar = fmaf(gs[v], inR[r * EW + c + v], ar);
                as = fmaf(gs[v], inS[r * EW + c + v], as);
            }
            tmpR[n] = ar; tmpS[n] = as;
        }
        __syncthreads();
        const float ck = p.mean_r[k], dk = p.mean_s[k];
#pragma unroll
        for (int t = 0; t < 4; t++) {                      // column pass
            const int oy = oyb + 4 * t;
            float a = 0.f, b = 0.f;
            for (int u = 0; u < K; u++) {
                a = fmaf(gs[u], tmpR[(oy + u) * MO_TW + ox], a);
                b = fmaf(gs[u], tmpS[(oy + u) * MO_TW + ox], b);
            }
            m2[t] = fmaf(a, a, m2[t]);
            p3[t] = fmaf(ck, a, p3[t]);
            uu[t] = fmaf(dk, a, uu[t]);
            p1[t] = fmaf(dk, b, p1[t]);
            vv[t] = fmaf(ck, b, vv[t]);
            const int y = y0 + oy, x = x0 + ox;
            if (p.fa && y < p.H && x < p.pitch) {
                p.fa[k * fstride + (size_t)y * p.pitch + x] = a;
                p.fb[k * fstride + (size_t)y * p.pitch + x] = b;
            }
        }
    }
    // window-filter the per-pixel sums of squares: T3 = w (*) sum_k R'^2, T1 = w (*) sum_k S'^2
    __syncthreads();
    for (int n = tid; n < EH * MO_TW; n += MO_NT) {
        const int r = n / MO_TW, c = n - r * MO_TW;
        float ar = 0.f, as = 0.f;
        for (int v = 0; v < K; v++) {
            ar = fmaf(gs[v], qR[r * EW + c + v], ar);
            as = fmaf(gs[v], qS[r * EW + c + v], as);
        }
        tmpR[n] = ar; tmpS[n] = as;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int oy = oyb + 4 * t;
        float t3 = 0.f, t1 = 0.f;
        for (int u = 0; u < K; u++) {
            t3 = fmaf(gs[u], tmpR[(oy + u) * MO_TW + ox], t3);
            t1 = fmaf(gs[u], tmpS[(oy + u) * MO_TW + ox], t1);
        }
        const int y = y0 + oy, x = x0 + ox;
        if (y < p.H && x < p.pitch) {
            p.auxR[(size_t)y * p.pitch + x] = make_float4(t3, p3[t], uu[t], m2[t]);
            p.auxS[(size_t)y * p.pitch + x] = make_float4(t1, p1[t], vv[t], 0.f);
        }
    }
}

// ------------------------------------------------------------------ shift tables

struct TableParams {
    const float *A, *B;          // stacks [Na][H][pitch]: A is read at q+s, B at q
    float *table;                // [S*S][rows_p][cols_p]
    const float *g;              // window factor (FILTER only)
    int Na, Nw, H, W, pitch;
    int oy, ox;                  // raw coordinates of table element (0,0)
    int rows_p, cols_p;          // padded table plane (multiples of the tile)
    int TH;                      // tile height
    int EH, EWs;                 // extended tile (B tile): rows, cols (multiple of 4)
    int AH, AP;                  // A tile rows, pitch
    int nslot, resident;         // frame slots in smem; 1 = whole stack stays resident
    int nt;                      // threads per block
};

__device__ __forceinline__ void cp_async4(float *dst, const float *src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int PF_DEPTH = 2;       // frames in flight

// S: shifts per axis (2*max_shift-1); SH: shift rows accumulated per pass;
// FILTER: apply the separable window to the accumulated correlation before storing.
template <int S, int SH, bool FILTER>
__global__ void __launch_bounds__(MAX_NT) shift_table_kernel(TableParams p)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int HS = (S - 1) / 2;                  // max |shift|
    constexpr int NA4 = (S + 3 + 3) / 4;             // float4 loads covering S+3 floats of an A row
    const int tid = threadIdx.x;
    const int slot_floats = p.AH * p.AP + p.EH * p.EWs;
    float *cbuf = sm + (size_t)p.nslot * slot_floats; // [S][EH][EWs] (FILTER only)
    __shared__ float gs[UMPA_MAX_K];
    const int K = 2 * p.Nw + 1;
    if (FILTER && tid < K) gs[tid] = p.g[tid];

    const int halo = FILTER ? p.Nw : 0;
    const int ty0 = blockIdx.y * p.TH, tx0 = blockIdx.x * TILE_W;       // table coords of the tile
    const int by = p.oy + ty0 - halo, bx = p.ox + tx0 - halo;           // raw origin of the B tile
    const int ay = by - HS, ax = bx - HS;                               // raw origin of the A tile
    const size_t fstride = (size_t)p.H * p.pitch;

    // this thread's strip of 4 consecutive extended-tile pixels
    const int spr = p.EWs / 4;                       // strips per row
    const int nstrips = p.EH * spr;
    const bool active = tid < nstrips;
    const int er = active ? tid / spr : 0, ec = active ? 4 * (tid - er * spr) : 0;

    auto issue_load = [&](int frame, int slot) {
        float *As = sm + (size_t)slot * slot_floats, *Bs = As + p.AH * p.AP;
        const float *Ag = p.A + frame * fstride, *Bg = p.B + frame * fstride;
        const int na = p.AH * p.AP;
        for (int n = tid; n < na; n += p.nt) {
            const int r = n / p.AP, c = n - r * p.AP;
            const int y = ay + r, x = ax + c;
            const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
            cp_async4(As + n, ok ? Ag + (size_t)y * p.pitch + x : Ag, ok);
        }
        const int nb = p.EH * p.EWs;
        for (int n = tid; n < nb; n += p.nt) {
            const int r = n / p.EWs, c = n - r * p.EWs;
            const int y = by + r, x = bx + c;
            const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
            cp_async4(Bs + n, ok ? Bg + (size_t)y * p.pitch + x : Bg, ok);
        }
    };

    constexpr int NPASS = (S + SH - 1) / SH;
    const int total = NPASS * p.Na;
    const int nload = p.resident ? p.Na : total;
    for (int g = 0; g < PF_DEPTH; g++) {
        if (g < nload) issue_load(g % p.Na, p.resident ? g : g % p.nslot);
        cp_async_commit();
    }

    float acc[SH][S][4];
    for (int g = 0; g < total; g++) {
        const int k = g % p.Na, pass = g / p.Na;
        if (g < nload + PF_DEPTH) {                  // loads may still be in flight
            cp_async_wait<PF_DEPTH - 1>();
            __syncthreads();
            const int gl = g + PF_DEPTH;
            if (gl < nload) issue_load(gl % p.Na, p.resident ? gl : gl % p.nslot);
            cp_async_commit();
        }
        if (k == 0) {
#pragma unroll
            for (int a = 0; a < SH; a++)
#pragma unroll
                for (int b = 0; b < S; b++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][b][c] = 0.f;
        }
        if (active) {
            const float *As = sm + (size_t)(p.resident ? k : g % p.nslot) * slot_floats;
            const float *Bs = As + p.AH * p.AP;
            const float4 b4 = *reinterpret_cast<const float4 *>(Bs + er * p.EWs + ec);
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int sh = 0; sh < SH; sh++) {
                const int si = pass * SH + sh;
                if (si < S) {
                    const float *arow = As + (er + si) * p.AP + ec;
                    float av[4 * NA4];
#pragma unroll
                    for (int v = 0; v < NA4; v++) {
                        const float4 t = *reinterpret_cast<const float4 *>(arow + 4 * v);
                        av[4 * v] = t.x; av[4 * v + 1] = t.y; av[4 * v + 2] = t.z; av[4 * v + 3] = t.w;
                    }
#pragma unroll
                    for (int sj = 0; sj < S; sj++)
#pragma unroll
                        for (int x = 0; x < 4; x++) acc[sh][sj][x] = fmaf(bv[x], av[sj + x], acc[sh][sj][x]);
                }
            }
        }
        if (k == p.Na - 1) {
            // ---- epilogue of this pass: one shift row at a time ----
#pragma unroll
            for (int sh = 0; sh < SH; sh++) {
                const int si = pass * SH + sh;
                if (si >= S) break;
                float *plane0 = p.table + (size_t)(si * S) * p.rows_p * p.cols_p;
                if (!FILTER) {
                    if (active) {
                        const int ty = ty0 + er, tx = tx0 + ec;
#pragma unroll
                        for (int sj = 0; sj < S; sj++)
                            *reinterpret_cast<float4 *>(plane0 + ((size_t)sj * p.rows_p + ty) * p.cols_p + tx) =
                                make_float4(acc[sh][sj][0], acc[sh][sj][1], acc[sh][sj][2], acc[sh][sj][3]);
                    }
                } else {
                    const int plane = p.EH * p.EWs;
                    if (active) {
#pragma unroll
                        for (int sj = 0; sj < S; sj++)
                            *reinterpret_cast<float4 *>(cbuf + sj * plane + er * p.EWs + ec) =
                                make_float4(acc[sh][sj][0], acc[sh][sj][1], acc[sh][sj][2], acc[sh][sj][3]);
                    }
                    __syncthreads();
                    // row pass: thread (row r of EH, output strip oc) -> 4 outputs per shift, kept in registers
                    const int ospr = TILE_W / 4;
                    const bool ract = tid < p.EH * ospr;
                    const int rr = ract ? tid / ospr : 0, oc = ract ? 4 * (tid - rr * ospr) : 0;
                    float rp[S][4];
                    if (ract) {
#pragma unroll
                        for (int sj = 0; sj < S; sj++) {
                            const float *src = cbuf + sj * plane + rr * p.EWs + oc;
                            float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                            float w0 = src[0], w1 = src[1], w2 = src[2], w3;
                            for (int v = 0; v < K; v++) {       // sliding 4-wide window over the row
                                w3 = src[v + 3];
                                const float gv = gs[v];
                                o0 = fmaf(gv, w0, o0); o1 = fmaf(gv, w1, o1);
                                o2 = fmaf(gv, w2, o2); o3 = fmaf(gv, w3, o3);
                                w0 = w1; w1 = w2; w2 = w3;
                            }
                            rp[sj][0] = o0; rp[sj][1] = o1; rp[sj][2] = o2; rp[sj][3] = o3;
                        }
                    }
                    __syncthreads();
                    if (ract) {
#pragma unroll
                        for (int sj = 0; sj < S; sj++)
                            *reinterpret_cast<float4 *>(cbuf + sj * plane + rr * p.EWs + oc) =
                                make_float4(rp[sj][0], rp[sj][1], rp[sj][2], rp[sj][3]);
                    }
                    __syncthreads();
                    // column pass: work item = (shift sj, output row y, strip oc)
                    const int items = S * p.TH * ospr;
                    for (int it = tid; it < items; it += p.nt) {
                        const int sj = it / (p.TH * ospr);
                        const int rem = it - sj * (p.TH * ospr);
                        const int y = rem / ospr, c4 = 4 * (rem - y * ospr);
                        const float *src = cbuf + sj * plane + y * p.EWs + c4;
                        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                        for (int u = 0; u < K; u++) {
                            const float4 t = *reinterpret_cast<const float4 *>(src + u * p.EWs);
                            const float gu = gs[u];
                            o0 = fmaf(gu, t.x, o0); o1 = fmaf(gu, t.y, o1);
                            o2 = fmaf(gu, t.z, o2); o3 = fmaf(gu, t.w, o3);
                        }
                        *reinterpret_cast<float4 *>(plane0 + ((size_t)sj * p.rows_p + ty0 + y) * p.cols_p + tx0 + c4) =
                            make_float4(o0, o1, o2, o3);
                    }
                    __syncthreads();
                }
            }
        }
    }
}

// ------------------------------------------------------------------ table-driven walk

struct WalkParams {
    const float *tabX, *tabM;       // cross / mean tables [S*S][rows_p][cols_p]; tabM nullptr for NoDF
    const float4 *auxS, *auxR;      // [H][pitch], raw coordinates
    int pitch;
    int rows_p, cols_p;
    int oy, ox;                     // raw coords of table (0,0)
    int kind, Na, max_shift, subpx;
    double sw, cd, cc, dd;          // sum of window; sum_k c_k d_k, c_k^2, d_k^2
    const double *quad;
};

struct TableEval {
    const WalkParams &w;
    int ty, tx;                     // table coords of this pixel
    double t1, V;                   // pixel-only terms

    __device__ int operator()(int si, int sj, double &cost, FitArgs &args) const
    {
        const int ms = w.max_shift;
        if (si <= -ms || si >= ms) return UMPA_ST_BOUND;
        if (sj <= -ms) return UMPA_ST_BOUND | UMPA_ST_DIM;
        if (sj >= ms) return UMPA_ST_BOUND | UMPA_ST_DIM | UMPA_ST_POS;
        const int S = 2 * ms - 1;
        const size_t e = ((size_t)((si + ms - 1) * S + (sj + ms - 1)) * w.rows_p + ty) * w.cols_p + tx;
        const float4 r = __ldg(w.auxR + (size_t)(w.oy + ty + si) * w.pitch + (w.ox + tx + sj));
        const double T3 = r.x, P3 = r.y, U = r.z, M2 = r.w;
        const double t3 = T3 + 2. * P3 + w.sw * w.cc;
        const double lin = U + V + w.sw * w.cd;
        const double t5 = (double)__ldg(w.tabX + e) + lin;
        if (w.kind == UMPA_DF) {
            const double t2 = M2 / (w.sw * w.sw) + 2. * P3 / w.sw + w.cc;
            const double t6 = w.sw * t2;
            const double t4 = (double)__ldg(w.tabM + e) / w.sw + lin;
            const double den = t2 * t3 - t6 * t6;
            const double Kc = (t2 * t5 - t4 * t6) / den;
            const double beta = (t3 * t4 - t5 * t6) / den;
            args.t = beta + Kc;
            args.v = Kc / args.t;
            cost = (t1 + beta * beta * t2 + Kc * Kc * t3 - 2. * beta * t4 - 2. * Kc * t5 + 2. * beta * Kc * t6) / w.Na;
        } else {
            args.t = t5 / t3;
            cost = (t1 - t5 * args.t) / w.Na;
        }
        return UMPA_ST_OK;
    }
};

__global__ void __launch_bounds__(128) table_walk_kernel(WalkParams w, RoiView roi, umpa_outputs out)
{
    const int xj = blockIdx.x * blockDim.x + threadIdx.x;
    const int xi = blockIdx.y;
    if (xj >= roi.N1 || xi >= roi.N0) return;
    const size_t n = (size_t)xi * roi.N1 + xj;
    if (roi.cover && roi.cover[n] < roi.cover_threshold) return;
    const int ty = roi.step0 * xi, tx = roi.step1 * xj;
    const float4 s = __ldg(w.auxS + (size_t)(w.oy + ty) * w.pitch + (w.ox + tx));
    TableEval eval{w, ty, tx, (double)s.x + 2. * (double)s.y + w.sw * w.dd, (double)s.z};
    FitArgs args{0., 0.};
    double d[25], a[16], uv[2] = {roi.uv0[0], roi.uv0[1]}, f = 0.;
    int ncalls;
#pragma unroll
    for (int t = 0; t < 16; t++) a[t] = 0.;
    const int st = walk_minimise(eval, w.subpx, w.quad, args, f, uv, d, a, ncalls);
    store_pixel(out, n, w.kind, st, f, args, uv, d, a, ncalls, true);
}

// ------------------------------------------------------------------ host side

template <int S, bool FILTER>
int launch_shift_table(const TableParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    constexpr int SH = (S <= 9) ? 1 : 1;
    auto kern = shift_table_kernel<S, SH, FILTER>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { umpa_set_error("cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e)); return UMPA_ERR_CUDA; }
    kern<<<grid, p.nt, smem, st>>>(p);
    UMPA_CUDA(cudaGetLastError());
    return UMPA_OK;
}

template <bool FILTER>
int dispatch_shift_table(int S, const TableParams &p, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (S) {
        case 3: return launch_shift_table<3, FILTER>(p, grid, smem, st);
        case 5: return launch_shift_table<5, FILTER>(p, grid, smem, st);
        case 7: return launch_shift_table<7, FILTER>(p, grid, smem, st);
        case 9: return launch_shift_table<9, FILTER>(p, grid, smem, st);
        case 11: return launch_shift_table<11, FILTER>(p, grid, smem, st);
        case 13: return launch_shift_table<13, FILTER>(p, grid, smem, st);
        case 15: return launch_shift_table<15, FILTER>(p, grid, smem, st);
        case 17: return launch_shift_table<17, FILTER>(p, grid, smem, st);
        case 19: return launch_shift_table<19, FILTER>(p, grid, smem, st);
    }
    umpa_set_error("table path: max_shift %d not instantiated", (S + 1) / 2);
    return UMPA_ERR_UNSUPPORTED;
}

// Fill in tile geometry for one table kernel; returns dynamic smem bytes (0 = does not fit).
size_t plan_tiles(TableParams &p, int S, bool filter)
{
    const int HS = (S - 1) / 2;
    const int halo = filter ? p.Nw : 0;
    const int NA4 = (S + 3 + 3) / 4;
    for (int TH = 16; TH >= 2; TH /= 2) {
        p.TH = TH;
        p.EH = TH + 2 * halo;
        p.EWs = 4 * ((TILE_W + 2 * halo + 3) / 4);
        p.AH = p.EH + 2 * HS;
        p.AP = p.EWs - 4 + 4 * NA4;
        const int nstrips = p.EH * (p.EWs / 4);
        if (nstrips > MAX_NT) continue;
        p.nt = std::max(128, 32 * ((nstrips + 31) / 32));
        const size_t slot = (size_t)(p.AH * p.AP + p.EH * p.EWs) * sizeof(float);
        const size_t cbuf = filter ? (size_t)S * p.EH * p.EWs * sizeof(float) : 0;
        const size_t budget = SMEM_CAP - 1024;
        if ((size_t)p.Na * slot + cbuf <= budget) {
            p.resident = 1; p.nslot = p.Na;
            return p.Na * slot + cbuf;
        }
        const int ns = PF_DEPTH + 1;
        if (ns * slot + cbuf <= budget) {
            p.resident = 0; p.nslot = ns;
            return ns * slot + cbuf;
        }
    }
    return 0;
}

}  // namespace

// FP64 device stacks -> per-frame means + centred FP32 stacks (pitch multiple of 4 floats)
int table_prepare_frames(umpa_model *m, cudaStream_t st)
{
    if (!m->uniform || m->masked || m->kind == UMPA_DFKERNEL) return UMPA_OK;
    const int Na = m->Na, H = m->H, W = m->W;
    m->pitch = 4 * ((W + 3) / 4);
    const size_t n32 = (size_t)Na * H * m->pitch;
    if (!m->d_sam32) {
        UMPA_CUDA(pool_malloc((void **)&m->d_sam32, n32 * sizeof(float)));
        UMPA_CUDA(pool_malloc((void **)&m->d_ref32, n32 * sizeof(float)));
        UMPA_CUDA(cudaMalloc(&m->d_mean_s, Na * sizeof(float)));
        UMPA_CUDA(cudaMalloc(&m->d_mean_r, Na * sizeof(float)));
        UMPA_CUDA(cudaMalloc(&m->d_means64, 2 * Na * sizeof(double)));
        UMPA_CUDA(cudaMalloc(&m->d_partials, (size_t)2 * Na * SUM_BLOCKS * sizeof(double)));
        m->dev_bytes += 2 * n32 * sizeof(float);
    }
    const size_t fe = (size_t)H * W;
    frame_partial_sums<<<dim3(SUM_BLOCKS, 2 * Na), 256, 0, st>>>(m->d_sam64, m->d_ref64, fe, Na, m->d_partials);
    finish_means<<<(2 * Na + 63) / 64, 64, 0, st>>>(m->d_partials, 2 * Na, 1. / (double)fe, m->d_means64,
                                                   m->d_mean_s, m->d_mean_r, Na);
    center_frames<<<dim3((m->pitch + 255) / 256, H, 2 * Na), 256, 0, st>>>(m->d_sam64, m->d_ref64, m->d_means64, Na,
                                                                          H, W, m->pitch, m->d_sam32, m->d_ref32);
    UMPA_CUDA(cudaGetLastError());
    std::vector<double> mu(2 * Na);
    UMPA_CUDA(cudaMemcpyAsync(mu.data(), m->d_means64, 2 * Na * sizeof(double), cudaMemcpyDeviceToHost, st));
    UMPA_CUDA(cudaStreamSynchronize(st));
    m->mean_s.assign(mu.begin(), mu.begin() + Na);
    m->mean_r.assign(mu.begin() + Na, mu.end());
    m->moments_valid = false;
    return UMPA_OK;
}

bool table_eligible(const umpa_model *m, const RoiView &roi, std::string *why)
{
    auto no = [&](const char *s) { if (why) *why = s; return false; };
    if (m->kind == UMPA_DFKERNEL) return no("DFKernel model");
    if (!m->uniform) return no("ragged frames or non-zero positions");
    if (m->masked) return no("masks");
    if (!m->separable) return no("window is not separable");
    if (m->refshift) return no("reference_shift=1");
    if (m->max_shift < 2 || m->max_shift > 10) return no("max_shift outside 2..10");
    if (m->Nw > 15) return no("window too large");
    if (roi.step0 * roi.step1 > 16) return no("sparse ROI (step product > 16)");
    if (!m->d_sam32) return no("FP32 stacks not prepared");
    return true;
}

int table_match(umpa_model *m, const RoiView &roi, const umpa_outputs &out, cudaStream_t st)
{
    const int S = 2 * m->max_shift - 1, HS = m->max_shift - 1;
    const bool df = m->kind == UMPA_DF;
    const int Na = m->Na, H = m->H, pitch = m->pitch;
    // dense table region in raw coordinates
    const int oy = roi.off0, ox = roi.off1;
    const int rows = (roi.N0 - 1) * roi.step0 + 1, cols = (roi.N1 - 1) * roi.step1 + 1;

    TableParams px{};
    px.Na = Na; px.Nw = m->Nw; px.H = H; px.W = m->W; px.pitch = pitch; px.oy = oy; px.ox = ox; px.g = m->d_g;
    TableParams pm = px;
    const size_t smx = plan_tiles(px, S, true);
    const size_t smm = df ? plan_tiles(pm, S, false) : 1;
    if (!smx || !smm) { umpa_set_error("table path: tile does not fit shared memory"); return UMPA_ERR_UNSUPPORTED; }
    // one table geometry for both kernels: pad to the larger tile height
    const int THmax = std::max(px.TH, df ? pm.TH : px.TH);
    const int rows_p = THmax * ((rows + THmax - 1) / THmax), cols_p = TILE_W * ((cols + TILE_W - 1) / TILE_W);
    px.rows_p = pm.rows_p = rows_p; px.cols_p = pm.cols_p = cols_p;
    const size_t tab_bytes = (size_t)S * S * rows_p * cols_p * sizeof(float);
    int rc;
    if ((rc = scratch_reserve(m, m->tabX, tab_bytes))) return rc;
    if (df && (rc = scratch_reserve(m, m->tabM, tab_bytes))) return rc;
    const size_t img = (size_t)H * pitch;
    if ((rc = scratch_reserve(m, m->auxS, img * sizeof(float4)))) return rc;
    if ((rc = scratch_reserve(m, m->auxR, img * sizeof(float4)))) return rc;
    if (df) {
        if ((rc = scratch_reserve(m, m->filtA, (size_t)Na * img * sizeof(float)))) return rc;
        if ((rc = scratch_reserve(m, m->filtB, (size_t)Na * img * sizeof(float)))) return rc;
    }
    if (m->profiling) {
        if (!m->ev[0]) for (int i = 0; i < 5; i++) UMPA_CUDA(cudaEventCreate(&m->ev[i]));
        UMPA_CUDA(cudaEventRecord(m->ev[0], st));
    }

    // 1. moments over the bounding box of everything the walk can touch
    {
        MomentsParams mp{};
        mp.sam = m->d_sam32; mp.ref = m->d_ref32;
        mp.fa = df ? (float *)m->filtA.p : nullptr; mp.fb = df ? (float *)m->filtB.p : nullptr;
        mp.auxS = (float4 *)m->auxS.p; mp.auxR = (float4 *)m->auxR.p;
        mp.g = m->d_g; mp.mean_s = m->d_mean_s; mp.mean_r = m->d_mean_r;
        mp.Na = Na; mp.Nw = m->Nw; mp.H = H; mp.W = m->W; mp.pitch = pitch;
        const int ylo = std::max(0, oy - HS), yhi = std::min(H, oy + rows + HS);
        const int xlo = std::max(0, ox - HS), xhi = std::min(m->W, ox + cols + HS);
        mp.ty0 = ylo / MO_TH; mp.tx0 = xlo / MO_TW;
        dim3 grid((xhi + MO_TW - 1) / MO_TW - mp.tx0, (yhi + MO_TH - 1) / MO_TH - mp.ty0);
        const int EH = MO_TH + 2 * m->Nw, EW = MO_TW + 2 * m->Nw;
        const size_t smem = (size_t)(4 * EH * EW + 2 * EH * MO_TW) * sizeof(float);
        UMPA_CUDA(cudaFuncSetAttribute(moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        moments_kernel<<<grid, MO_NT, smem, st>>>(mp);
        UMPA_CUDA(cudaGetLastError());
        m->last_launches++;
    }
    if (m->profiling) UMPA_CUDA(cudaEventRecord(m->ev[1], st));

    // 2. cross table: A = centred reference, B = centred sample, window-filtered
    px.A = m->d_ref32; px.B = m->d_sam32; px.table = (float *)m->tabX.p;
    {
        dim3 grid(cols_p / TILE_W, rows_p / px.TH);
        if ((rc = dispatch_shift_table<true>(S, px, grid, smx, st))) return rc;
        m->last_launches++;
    }
    if (m->profiling) UMPA_CUDA(cudaEventRecord(m->ev[2], st));

    // 3. mean table (DF): A = a_k, B = b_k, no filter
    if (df) {
        pm.A = (const float *)m->filtA.p; pm.B = (const float *)m->filtB.p; pm.table = (float *)m->tabM.p;
        dim3 grid(cols_p / TILE_W, rows_p / pm.TH);
        if ((rc = dispatch_shift_table<false>(S, pm, grid, smm, st))) return rc;
        m->last_launches++;
    }
    if (m->profiling) UMPA_CUDA(cudaEventRecord(m->ev[3], st));

    // 4. walk
    {
        WalkParams w{};
        w.tabX = (const float *)m->tabX.p; w.tabM = df ? (const float *)m->tabM.p : nullptr;
        w.auxS = (const float4 *)m->auxS.p; w.auxR = (const float4 *)m->auxR.p;
        w.pitch = pitch; w.rows_p = rows_p; w.cols_p = cols_p; w.oy = oy; w.ox = ox;
        w.kind = m->kind; w.Na = Na; w.max_shift = m->max_shift; w.subpx = m->subpx;
        w.sw = m->win_sum; w.quad = m->d_quad;
        double cd = 0., cc = 0., dd = 0.;
        for (int k = 0; k < Na; k++) {
            cd += m->mean_r[k] * m->mean_s[k];
            cc += m->mean_r[k] * m->mean_r[k];
            dd += m->mean_s[k] * m->mean_s[k];
        }
        w.cd = cd; w.cc = cc; w.dd = dd;
        const int threads = 128;
        dim3 grid((roi.N1 + threads - 1) / threads, roi.N0);
        table_walk_kernel<<<grid, threads, 0, st>>>(w, roi, out);
        UMPA_CUDA(cudaGetLastError());
        m->last_launches++;
    }
    if (m->profiling) { UMPA_CUDA(cudaEventRecord(m->ev[4], st)); m->ev_valid = true; }
    return UMPA_OK;
}
