// kernel_path.cu -- UMPAModelDFKernel on the fast path: the per-pixel 17x17 blur fused into
// the window pass.
//
// Reference (UMPA/lib/Model.cpp:997-1151, kernel ctor 88-117, Utils.cpp:46-50, 85-97): for
// every cost evaluation at pixel p and shift s it blurs each of the (2Nw+1)^2 window
// elements of each reference frame with the pixel's own Gaussian kernel k_p (289 MACs per
// element), then forms  t3 = sum w B^2,  t5 = sum w B S,  T = t5/t3,  cost = (t1 - t5 T)/Na.
// ~17 evaluations per pixel => ~6 M FP64 MACs per pixel.
//
// Here: k_p depends on the OUTPUT pixel only (it is built once per pixel, Model.cpp:1228-1229),
// so the blurred reference is needed on the (K+S-1)^2 patch around p only.  One warp slot owns
// one pixel; per frame it computes that patch  B'(q) = sum_v k_p(v) R'(q+v)  ONCE (the
// direct-form work SURVEY 8d counts: (K+S-1)^2 x 289 MACs per pixel and frame) and derives all
// S^2 shifts from it:
//     t3c(s) = [ w (*) sum_k (B'_k^2 + 2 sigma c_k B'_k) ](p+s)        (filtered once, after the frames)
//     t5c(s) = sum_k sum_u w(u) B'_k(p+s+u) S_k(p+u)                   (register partials per patch row)
// with R' = R - c_k the mean-centred FP32 reference, sigma = sum k_p.  The uncentred sums are
//     t3 = t3c + sigma^2 sw sum c_k^2,   t5 = t5c + sigma (V + sw sum c_k d_k)
// rebuilt in FP64 by the walk (table_path.cu, TableEval).  Output: one row of 2 S^2 + 1 floats
// per pixel (pixel-major table), consumed by table_walk_kernel.
//
// Mapping: CTA = 256 threads = 8 warps, NP = 8 * PPW pixels of ONE output column (consecutive
// output rows), so all patches share one reference region of ((NP-1) step + PS + 16) x (PS + 16)
// floats per frame, double-buffered with cp.async.  Lane (slot, y) owns patch row y of pixel
// `slot` of its warp: PS accumulators, the 17 kernel taps of a row are applied to a register
// row of PS+16 reference values (float4 LDS, conflict-free pitch), 17*PS FMAs per 13 LDS.128.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int KT_NT = 256;
constexpr int KSIDE = UMPA_KSIDE;          // 17
constexpr int KROW = 20;                   // kernel row padded to 5 float4
constexpr int KPIX = KSIDE * KROW;         // floats per pixel kernel in shared memory

struct KTableParams {
    const float *sam, *ref;                // centred stacks [Na][H][pitch]
    const float *mean_s, *mean_r;          // d_k, c_k
    const float *g;                        // 1-D window factor (K floats)
    const double *abc;                     // (N0, N1, 3)
    float *tab;                            // [N0*N1][TS]
    int Na, H, W, pitch;
    int off0, step0, N0, off1, step1, N1;
    int TS;
    int RR, SR;                            // rows of the reference / sample regions of a CTA
};

template <int NW, int S>
struct KTGeom {
    static constexpr int K = 2 * NW + 1, HS = (S - 1) / 2, PS = K + S - 1;
    static constexpr int REACH = HS + NW + UMPA_KWS;       // how far the reference is read from the pixel
    static constexpr int RW = PS + 2 * UMPA_KWS;           // region width
    static constexpr int PPW = 32 / PS;                    // pixels per warp
    static constexpr int SW = 32 / PPW;                    // lanes per pixel slot
    static constexpr int NP = (KT_NT / 32) * PPW;          // pixels per CTA
    static constexpr int NR4 = (RW + 3) / 4;
    static constexpr int RP = 4 * (NR4 | 1);               // region pitch: multiple of 4 floats, odd number of float4
    static constexpr int QP = PS | 1;                      // pitch of the final q3 patch
    static constexpr int WP = (K + 3) & ~3;                // pitch of the weighted sample window
    static constexpr int PSTR = (K * S) | 1;               // stride of one lane's t5 partials
    static_assert(PS <= 32, "patch does not fit a warp");
};

__device__ __forceinline__ void cp_async4(float *dst, const float *src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(sz));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NW, int S, bool RS>
size_t ktable_smem_floats(int RR, int SR)
{
    using G = KTGeom<NW, S>;
    auto a4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
    return 2 * (size_t)RR * G::RP + a4(2 * (size_t)SR * (RS ? G::PS : G::K)) + (size_t)G::NP * KPIX + (size_t)G::NP * G::K * G::WP +
           a4(G::K * G::K) + a4((size_t)G::NP * G::PS * G::QP) + (size_t)G::NP * G::PS * G::PSTR;
}

// RS = reference_shift (assign_coordinates = 'ref', Model.cpp:1060-1075): the blurred reference window stays at
// the pixel and the SAMPLE window moves by -s.  Same machinery with the roles swapped: the patch rows held by
// the lanes are rows of the (uncentred) sample, the K x K weights are w * B' (the blurred reference at the
// window, the centre of the blurred patch), t3 does not depend on the shift, and t5 comes out indexed by the
// negated shift -- which is the index TableEval<RS> looks up.
template <int NW, int S, bool RS>
__global__ void __launch_bounds__(KT_NT, 1) ktable_kernel(KTableParams p)
{
    using G = KTGeom<NW, S>;
    constexpr int HS = G::HS, SCW = RS ? G::PS : G::K;        // columns of the sample region
    constexpr int K = G::K, PS = G::PS, RW = G::RW, PPW = G::PPW, SW = G::SW, NP = G::NP;
    constexpr int NR4 = G::NR4, RP = G::RP, QP = G::QP, WP = G::WP, PSTR = G::PSTR, REACH = G::REACH;
    extern __shared__ __align__(16) float sm[];
    auto a4 = [](int n) { return (n + 3) & ~3; };
    float *Rbuf = sm;
    float *Sbuf = Rbuf + 2 * p.RR * RP;
    float *Ks = Sbuf + a4(2 * p.SR * SCW);
    float *Ws = Ks + NP * KPIX;
    float *W2 = Ws + NP * K * WP;
    float *Qs = W2 + a4(K * K);
    float *Ps = Qs + a4(NP * PS * QP);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = lane / SW, y = lane - slot * SW;
    const bool lane_on = slot < PPW && y < PS;
    const int yc = min(y, PS - 1), slotc = min(slot, PPW - 1);
    const int pl = warp * PPW + slotc;                 // this lane's pixel inside the CTA
    const int xj = blockIdx.x, xi0 = blockIdx.y * NP;
    const int i0 = p.off0 + p.step0 * xi0, j0 = p.off1 + p.step1 * xj;
    const size_t fstride = (size_t)p.H * p.pitch;

    auto load = [&](int k, int b) {
        const float *R = p.ref + (size_t)k * fstride, *Sg = p.sam + (size_t)k * fstride;
        float *dR = Rbuf + b * p.RR * RP, *dS = Sbuf + b * p.SR * SCW;
        for (int n = tid; n < p.RR * RW; n += KT_NT) {
            const int rr = n / RW, cc = n - rr * RW;
            const int gy = i0 - REACH + rr, gx = j0 - REACH + cc;
            const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
            cp_async4(dR + rr * RP + cc, R + (ok ? (size_t)gy * p.pitch + gx : 0), ok);
        }
        for (int n = tid; n < p.SR * SCW; n += KT_NT) {
            const int sr = n / SCW, cc = n - sr * SCW;
            const int gy = i0 - NW - (RS ? HS : 0) + sr, gx = j0 - NW - (RS ? HS : 0) + cc;
            const bool ok = gy >= 0 && gy < p.H && gx >= 0 && gx < p.W;
            cp_async4(dS + n, Sg + (ok ? (size_t)gy * p.pitch + gx : 0), ok);
        }
    };
    load(0, 0);
    asm volatile("cp.async.commit_group;\n" ::);

    for (int n = tid; n < K * K; n += KT_NT) W2[n] = __ldg(p.g + n / K) * __ldg(p.g + n % K);

    // ---- this warp's blur kernels: k = exp(-a i^2 - b i j - c j^2) / sum, FP64, stored as FP32 ----
    // Support of the FP32 kernels: taps below 1e-10 of the pixel's largest tap are set to zero (the Gaussian of a
    // typical (a, b, c) ~ 0.5 is that small beyond |i| = 6; what is dropped is far below the FP32 rounding of the
    // 289-term sum, and sigma is the sum of the taps that remain, so the centring identity stays exact).  RI / RJ:
    // largest |i| / |j| with a non-zero tap among this warp's pixels -- the blur loops stop there.  (A pixel's
    // result does not depend on which other pixel shares its warp: the extra taps it may be run over are zeros.)
    int RI = 0, RJ = 0;
    float dsig_mine = 0.f;                             // sigma - 1 of this lane's pixel
    for (int s = 0; s < PPW; s++) {
        const int pli = warp * PPW + s, xis = xi0 + pli;
        double a = 0., b = 0., c = 0.;
        if (xis < p.N0) {
            const double *q = p.abc + ((size_t)xis * p.N1 + xj) * 3;
            a = q[0]; b = q[1]; c = q[2];
        }
        constexpr int NE = (KSIDE * KSIDE + 31) / 32;
        double v[NE], part = 0.;
#pragma unroll
        for (int t = 0; t < NE; t++) {
            const int e = lane + 32 * t;
            const int ii = e / KSIDE - UMPA_KWS, jj = e % KSIDE - UMPA_KWS;
            v[t] = e < KSIDE * KSIDE ? exp(-a * ii * ii - b * ii * jj - c * jj * jj) : 0.;
            part += v[t];
        }
        const double norm = warp_sum(part);
        double sp = 0., vmax = 0.;
#pragma unroll
        for (int t = 0; t < NE; t++) vmax = fmax(vmax, v[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        float *kd = Ks + pli * KPIX;
        int ri = 0, rj = 0;
#pragma unroll
        for (int t = 0; t < NE; t++) {
            const int e = lane + 32 * t;
            if (e < KSIDE * KSIDE) {
                const bool keep = !(v[t] < 1e-10 * vmax);      // (NaN-safe: a broken kernel keeps its full support)
                const float kf = keep ? (float)(v[t] / norm) : 0.f;
                kd[(e / KSIDE) * KROW + e % KSIDE] = kf;
                sp += (double)kf;
                if (keep) {
                    ri = max(ri, abs(e / KSIDE - UMPA_KWS));
                    rj = max(rj, abs(e % KSIDE - UMPA_KWS));
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ri = max(ri, __shfl_xor_sync(0xffffffffu, ri, o));
            rj = max(rj, __shfl_xor_sync(0xffffffffu, rj, o));
        }
        RI = max(RI, ri); RJ = max(RJ, rj);
        for (int e = lane; e < KSIDE * (KROW - KSIDE); e += 32)
            kd[(e / (KROW - KSIDE)) * KROW + KSIDE + e % (KROW - KSIDE)] = 0.f;
        const double sig = warp_sum(sp);
        if (slotc == s) dsig_mine = (float)(sig - 1.);
    }
    const float sigf = 1.f + dsig_mine;
    __syncwarp();

    float q3[PS], P[K][S];
#pragma unroll
    for (int x = 0; x < PS; x++) q3[x] = 0.f;
#pragma unroll
    for (int a = 0; a < K; a++)
#pragma unroll
        for (int b = 0; b < S; b++) P[a][b] = 0.f;

    const int prow = pl * p.step0;
    const float *Kp = Ks + pl * KPIX;
    float *Wp = Ws + pl * K * WP;

    for (int k = 0; k < p.Na; k++) {
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();                               // frame k landed; everybody is done with frame k-1
        if (k + 1 < p.Na) load(k + 1, (k + 1) & 1);
        asm volatile("cp.async.commit_group;\n" ::);

        // weighted sample window of this pixel: ws(a,b) = w(a,b) * S_k(p + (a,b) - Nw)   (uncentred S)
        const float dk = __ldg(p.mean_s + k);
        if (!RS) {
            const float *Sb = Sbuf + (k & 1) * p.SR * SCW + prow * SCW;
            if (slot < PPW)
                for (int e = y; e < K * K; e += SW) {
                    const int a = e / K, b = e - a * K;
                    Wp[a * WP + b] = W2[e] * (Sb[e] + dk);
                }
            __syncwarp();
        }

        float acc[PS];
#pragma unroll
        for (int x = 0; x < PS; x++) acc[x] = 0.f;
        const float *Rb = Rbuf + (k & 1) * p.RR * RP + (prow + yc) * RP;
        // rows ii in [8-RI, 8+RI]; columns jj in [8-JR, 8+JR] with JR the compile-time bound >= RJ
        auto blur = [&](auto jrc) {
            constexpr int JR = decltype(jrc)::value;
#pragma unroll 1
            for (int ii = UMPA_KWS - RI; ii <= UMPA_KWS + RI; ii++) {
                float r[4 * NR4], kk[KROW];
#pragma unroll
                for (int v4 = 0; v4 < NR4; v4++) {
                    const float4 t = *reinterpret_cast<const float4 *>(Rb + ii * RP + 4 * v4);
                    r[4 * v4] = t.x; r[4 * v4 + 1] = t.y; r[4 * v4 + 2] = t.z; r[4 * v4 + 3] = t.w;
                }
#pragma unroll
                for (int v4 = 0; v4 < KROW / 4; v4++) {
                    const float4 t = *reinterpret_cast<const float4 *>(Kp + ii * KROW + 4 * v4);
                    kk[4 * v4] = t.x; kk[4 * v4 + 1] = t.y; kk[4 * v4 + 2] = t.z; kk[4 * v4 + 3] = t.w;
                }
#pragma unroll
                for (int jj = UMPA_KWS - JR; jj <= UMPA_KWS + JR; jj++)
#pragma unroll
                    for (int x = 0; x < PS; x++) acc[x] = fmaf(kk[jj], r[x + jj], acc[x]);
            }
        };
        if (RJ <= 4) blur(std::integral_constant<int, 4>());
        else if (RJ <= 6) blur(std::integral_constant<int, 6>());
        else blur(std::integral_constant<int, 8>());

        const float tsc = 2.f * sigf * __ldg(p.mean_r + k);
        float prod[PS];                                // what the K x K weights multiply: B' rows, or (RS) sample rows
        if (!RS) {
#pragma unroll
            for (int x = 0; x < PS; x++) { q3[x] = fmaf(acc[x], acc[x] + tsc, q3[x]); prod[x] = acc[x]; }
        } else {
            // the blurred reference at the window = rows / columns HS .. HS+K-1 of the blurred patch
            if (lane_on && y >= HS && y < HS + K) {
                const int a = y - HS;
#pragma unroll
                for (int b = 0; b < K; b++) {
                    const float wv = W2[a * K + b], bv = acc[HS + b];
                    Wp[a * WP + b] = wv * bv;
                    q3[0] = fmaf(wv * bv, bv + tsc, q3[0]);
                }
            }
            __syncwarp();
            const float *Sb = Sbuf + (k & 1) * p.SR * SCW + (prow + yc) * SCW;
#pragma unroll
            for (int x = 0; x < PS; x++) prod[x] = Sb[x] + dk;
        }
#pragma unroll
        for (int a = 0; a < K; a++) {
            float ws[WP];
#pragma unroll
            for (int v4 = 0; v4 < WP / 4; v4++) {
                const float4 t = *reinterpret_cast<const float4 *>(Wp + a * WP + 4 * v4);
                ws[4 * v4] = t.x; ws[4 * v4 + 1] = t.y; ws[4 * v4 + 2] = t.z; ws[4 * v4 + 3] = t.w;
            }
#pragma unroll
            for (int sj = 0; sj < S; sj++)
#pragma unroll
                for (int b = 0; b < K; b++) P[a][sj] = fmaf(ws[b], prod[sj + b], P[a][sj]);
        }
        __syncwarp();                                  // Wp is rewritten next frame
    }

    // ---- after the frames: window filter of q3, cross-lane sum of the t5 partials, table rows ----
    if (lane_on) {
        float *qd = Qs + (pl * PS + y) * QP;
#pragma unroll
        for (int x = 0; x < PS; x++) qd[x] = q3[x];
        float *pd = Ps + (size_t)(pl * PS + y) * PSTR;
#pragma unroll
        for (int a = 0; a < K; a++)
#pragma unroll
            for (int sj = 0; sj < S; sj++) pd[a * S + sj] = P[a][sj];
    }
    __syncwarp();
    for (int s = 0; s < PPW; s++) {
        const int pli = warp * PPW + s, xis = xi0 + pli;
        if (xis >= p.N0) continue;
        float *row = p.tab + ((size_t)xis * p.N1 + xj) * p.TS;
        const float *qb = Qs + pli * PS * QP;
        const float *pb = Ps + (size_t)pli * PS * PSTR;
        for (int n = lane; n < S * S; n += 32) {
            const int si = n / S, sj = n - si * S;
            float t3c = 0.f, t5c = 0.f;
#pragma unroll
            for (int a = 0; a < K; a++) {
                if (!RS) {
#pragma unroll
                    for (int b = 0; b < K; b++) t3c = fmaf(W2[a * K + b], qb[(si + a) * QP + sj + b], t3c);
                }
                t5c += pb[(size_t)(si + a) * PSTR + a * S + sj];
            }
            if (RS) {                                  // shift-independent: the lanes' window rows, summed
                for (int yy = HS; yy < HS + K; yy++) t3c += qb[yy * QP];
            }
            row[n] = t5c;
            row[S * S + n] = t3c;
        }
        const float ds = __shfl_sync(0xffffffffu, dsig_mine, s * SW);
        if (lane == 0) row[2 * S * S] = ds;
    }
}

typedef void (*KTableKernel)(KTableParams);

template <int NW, int S>
bool ktable_bind(int step0, bool rs, KTableKernel *k, size_t *smem, int *np, int *RR, int *SR)
{
    using G = KTGeom<NW, S>;
    *RR = (G::NP - 1) * step0 + G::PS + 2 * UMPA_KWS;
    *SR = (G::NP - 1) * step0 + (rs ? G::PS : G::K);
    *smem = (rs ? ktable_smem_floats<NW, S, true>(*RR, *SR) : ktable_smem_floats<NW, S, false>(*RR, *SR)) * sizeof(float);
    *np = G::NP;
    *k = rs ? ktable_kernel<NW, S, true> : ktable_kernel<NW, S, false>;
    return *smem <= (size_t)227 * 1024;
}

bool ktable_pick(int Nw, int S, int step0, bool rs, KTableKernel *k, size_t *smem, int *np, int *RR, int *SR)
{
#define KT_CASE(nw, s) if (Nw == nw && S == s) return ktable_bind<nw, s>(step0, rs, k, smem, np, RR, SR);
    KT_CASE(1, 3) KT_CASE(1, 5) KT_CASE(1, 7) KT_CASE(1, 9) KT_CASE(1, 11)
    KT_CASE(2, 3) KT_CASE(2, 5) KT_CASE(2, 7) KT_CASE(2, 9) KT_CASE(2, 11)
    KT_CASE(3, 3) KT_CASE(3, 5) KT_CASE(3, 7) KT_CASE(3, 9)
#undef KT_CASE
    return false;
}

}  // namespace

bool ktable_supported(int Nw, int max_shift, int step0, bool refshift)
{
    KTableKernel k; size_t smem; int np, RR, SR;
    return ktable_pick(Nw, 2 * max_shift - 1, step0, refshift, &k, &smem, &np, &RR, &SR);
}

int ktable_row_floats(int max_shift) { const int S = 2 * max_shift - 1; return (2 * S * S + 1 + 3) & ~3; }

int ktable_build(umpa_model *m, const RoiView &roi, float *tab, cudaStream_t st)
{
    KTableKernel kern = nullptr;
    size_t smem = 0;
    int np = 0, RR = 0, SR = 0;
    const int S = 2 * m->max_shift - 1;
    if (!ktable_pick(m->Nw, S, roi.step0, m->refshift != 0, &kern, &smem, &np, &RR, &SR)) {
        umpa_set_error("blur-table path: Nw=%d max_shift=%d step=%d not instantiated", m->Nw, m->max_shift, roi.step0);
        return UMPA_ERR_UNSUPPORTED;
    }
    KTableParams p{};
    p.sam = m->d_sam32; p.ref = m->d_ref32; p.mean_s = m->d_mean_s; p.mean_r = m->d_mean_r; p.g = m->d_g;
    p.abc = roi.abc; p.tab = tab;
    p.Na = m->Na; p.H = m->H; p.W = m->W; p.pitch = m->pitch;
    p.off0 = roi.off0; p.step0 = roi.step0; p.N0 = roi.N0; p.off1 = roi.off1; p.step1 = roi.step1; p.N1 = roi.N1;
    p.TS = ktable_row_floats(m->max_shift);
    p.RR = RR; p.SR = SR;
    UMPA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(roi.N1, (roi.N0 + np - 1) / np);
    kern<<<grid, KT_NT, smem, st>>>(p);
    UMPA_CUDA(cudaGetLastError());
    m->last_launches++;
    return UMPA_OK;
}
