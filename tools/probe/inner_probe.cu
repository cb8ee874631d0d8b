// Probe of the shift-table inner loop: smem-resident A/B tiles, accumulate acc[sh][sj][x] += B[x]*A[sj+x]
// for NF "frames", different thread tilings.  Reports cycles per frame per SM (all SMs busy, 1 CTA/SM
// unless stated).  FMA-issue bound for a 16x32 extended tile x 81 shifts = 324 cycles/frame.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int PW, int SH, int S, bool F2, int NT>
__global__ void __launch_bounds__(NT) probe(float *out, int NF, int G, int AP, int BP, int EH)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int NA = PW + S - 1;                 // A floats per row
    constexpr int NA4 = (NA + 3) / 4;
    const int tid = threadIdx.x;
    const int spr = 32 / PW, TG = EH * spr;
    const int grp = tid / TG, lt = tid % TG, er = lt / spr, ec = (lt % spr) * PW;
    float *As = sm, *Bs = sm + (EH + S - 1) * AP;
    for (int n = tid; n < (EH + S - 1) * AP + EH * BP; n += blockDim.x) sm[n] = (float)(n % 17) * 0.01f;
    __syncthreads();
    float acc[SH][S][PW];
#pragma unroll
    for (int a = 0; a < SH; a++)
#pragma unroll
        for (int b = 0; b < S; b++)
#pragma unroll
            for (int c = 0; c < PW; c++) acc[a][b][c] = 0.f;
    const int si0 = (grp * SH) % S;
    if (grp < G) {
        for (int f = 0; f < NF; f++) {
            float bv[PW];
#pragma unroll
            for (int v = 0; v < PW / 4; v++) {
                const float4 t = *reinterpret_cast<const float4 *>(Bs + er * BP + ec + 4 * v);
                bv[4 * v] = t.x; bv[4 * v + 1] = t.y; bv[4 * v + 2] = t.z; bv[4 * v + 3] = t.w;
            }
#pragma unroll
            for (int sh = 0; sh < SH; sh++) {
                const float *arow = As + (er + (si0 + sh) % S) * AP + ec;
                float av[4 * NA4];
#pragma unroll
                for (int v = 0; v < NA4; v++) {
                    const float4 t = *reinterpret_cast<const float4 *>(arow + 4 * v);
                    av[4 * v] = t.x; av[4 * v + 1] = t.y; av[4 * v + 2] = t.z; av[4 * v + 3] = t.w;
                }
                if (!F2) {
#pragma unroll
                    for (int sj = 0; sj < S; sj++)
#pragma unroll
                        for (int x = 0; x < PW; x++) acc[sh][sj][x] = fmaf(bv[x], av[sj + x], acc[sh][sj][x]);
                } else {
                    // packed FP32x2 FMA: pairs (x, x+1); odd sj needs the A row shifted by one register
                    float avo[4 * NA4];
#pragma unroll
                    for (int v = 0; v + 1 < 4 * NA4; v++) avo[v] = av[v + 1];
                    avo[4 * NA4 - 1] = 0.f;
#pragma unroll
                    for (int sj = 0; sj < S; sj++)
#pragma unroll
                        for (int x = 0; x < PW; x += 2) {
                            const float *src = (sj & 1) ? &avo[sj - 1 + x] : &av[sj + x];
                            unsigned long long a2, b2, c2;
                            asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(bv[x]), "f"(bv[x + 1]));
                            asm("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(src[0]), "f"(src[1]));
                            asm("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(acc[sh][sj][x]), "f"(acc[sh][sj][x + 1]));
                            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c2) : "l"(a2), "l"(b2));
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[sh][sj][x]), "=f"(acc[sh][sj][x + 1]) : "l"(c2));
                        }
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < SH; a++)
#pragma unroll
        for (int b = 0; b < S; b++)
#pragma unroll
            for (int c = 0; c < PW; c++) s += acc[a][b][c];
    out[blockIdx.x * blockDim.x + tid] = s;
}

template <int PW, int SH, int S, bool F2, int NT>
void run(const char *name, int G, int AP, int BP, int ctas_per_sm)
{
    const int EH = 16, NF = 2000;
    const int threads = G * EH * (32 / PW);
    const size_t smem = ((EH + S - 1) * AP + EH * BP) * sizeof(float);
    float *out;
    CK(cudaMalloc(&out, 148 * ctas_per_sm * threads * sizeof(float)));
    auto k = probe<PW, SH, S, F2, NT>;
    if (threads > NT) { printf("%s: threads %d > NT\n", name, threads); return; }
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<<<148 * ctas_per_sm, threads, smem>>>(out, NF, G, AP, BP, EH);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // shift rows covered per frame pass: G*SH (may exceed S); normalise to "81 shifts of a 16x32 tile"
    const double rows = (double)G * SH * ctas_per_sm;
    const double cyc = best * 1e-3 * 1.965e9 / NF;
    printf("%-34s threads %4d x%d  %7.1f cycles/frame  -> %6.1f cycles per 9 shift rows (FMA bound 324)\n", name, threads,
           ctas_per_sm, cyc, cyc * 9.0 / rows);
    cudaFree(out);
}

int main()
{
    run<4, 3, 9, false, 384>("PW4 SH3 G3 (current) AP40 BP32", 3, 40, 32, 1);
    run<4, 3, 9, true, 384>("PW4 SH3 G3 FFMA2", 3, 40, 32, 1);
    run<8, 1, 9, false, 576>("PW8 SH1 G9 AP44 BP36", 9, 44, 36, 1);
    run<8, 1, 9, false, 576>("PW8 SH1 G9 AP40 BP32 (conflicts)", 9, 40, 32, 1);
    run<8, 1, 9, true, 576>("PW8 SH1 G9 FFMA2 AP44 BP36", 9, 44, 36, 1);
    run<8, 2, 9, false, 320>("PW8 SH2 G5 AP44 BP36", 5, 44, 36, 1);
    run<8, 2, 9, true, 320>("PW8 SH2 G5 FFMA2", 5, 44, 36, 1);
    run<8, 1, 9, false, 320>("PW8 SH1 G5 x2 CTAs", 5, 44, 36, 2);
    run<4, 2, 9, false, 640>("PW4 SH2 G5 AP40 BP32", 5, 40, 32, 1);
    run<4, 1, 9, false, 1024>("PW4 SH1 G9(8) 1024thr", 8, 40, 32, 1);
    run<16, 1, 9, false, 288>("PW16 SH1 G9 AP44 BP36", 9, 44, 36, 1);
    return 0;
}
