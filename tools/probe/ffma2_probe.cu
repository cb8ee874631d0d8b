// ffma2_probe.cu -- what does Blackwell's packed FP32 FMA (fma.rn.f32x2 -> FFMA2) buy a kernel shaped like the
// shift-table main loop?  Per "frame" a thread does NL LDS.128 (conflict-free, as the strips of the table kernel)
// and 108 FMAs on 108 accumulators, either as 108 FFMA or as 54 FFMA2.  Prints cycles per frame per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int NL>     // MODE 0: FFMA, 1: FFMA2
__global__ void __launch_bounds__(384) probe(float *sink, int frames, long long *cycles)
{
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 24 * 1024; i += blockDim.x) sm[i] = 1e-3f * (i & 255);
    __syncthreads();
    float acc[108];
#pragma unroll
    for (int i = 0; i < 108; i++) acc[i] = 0.f;
    const float *base = sm + 4 * tid;
    const long long t0 = clock64();
    for (int f = 0; f < frames; f++) {
        float4 v[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) v[l] = *reinterpret_cast<const float4 *>(base + ((l * 1536 + f * 64) & 16383));
        const float *a = reinterpret_cast<const float *>(v);
        const float b0 = a[0], b1 = a[1], b2 = a[2], b3 = a[3];
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 108; i++) {
                const float bb = (i & 3) == 0 ? b0 : (i & 3) == 1 ? b1 : (i & 3) == 2 ? b2 : b3;
                acc[i] = fmaf(bb, a[4 + (i % (4 * NL - 4))], acc[i]);
            }
        } else {
            float2 *acc2 = reinterpret_cast<float2 *>(acc);
            const float2 *a2 = reinterpret_cast<const float2 *>(a);
            const float2 bb[4] = {make_float2(b0, b0), make_float2(b1, b1), make_float2(b2, b2), make_float2(b3, b3)};
#pragma unroll
            for (int i = 0; i < 54; i++) acc2[i] = __ffma2_rn(bb[i & 3], a2[2 + (i % (2 * NL - 2))], acc2[i]);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 108; i++) s += acc[i];
    if (s == 123.456f) sink[0] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int NL>
void run(const char *name, int sms)
{
    float *sink; long long *cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, sms * 8);
    const int frames = 20000;
    cudaFuncSetAttribute(probe<MODE, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<MODE, NL><<<sms, 384, 100 * 1024>>>(sink, frames, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    long long h[256]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < sms; i++) c += h[i];
    const double tf = 2. * 108 * 384 * (double)frames * sms / (best * 1e-3) / 1e12;
    printf("%-28s %2d LDS.128/frame: %7.1f cycles/frame/SM  %.3f ms  %.1f TFLOP/s  (%s)\n", name, NL, c / sms / frames, best, tf,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink); cudaFree(cyc);
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d; 384 threads per SM; 108 FMA per thread and frame (FMA-pipe floor 324 cycles, LDS floor 48*NL)\n", sms);
    run<0, 2>("FFMA", sms);  run<1, 2>("FFMA2", sms);
    run<0, 6>("FFMA", sms);  run<1, 6>("FFMA2", sms);
    run<0, 8>("FFMA", sms);  run<1, 8>("FFMA2", sms);
    run<0, 10>("FFMA", sms); run<1, 10>("FFMA2", sms);
    return 0;
}
