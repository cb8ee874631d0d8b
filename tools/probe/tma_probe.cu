// Standalone probe: one 3-D TMA box load (possibly with negative / unaligned coordinates)
// into shared memory, copied back to global and checked on the host.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int bw, int bh, int c0, int c1, int c2, int mode)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        if (mode & 1) asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        if (mode & 2) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::
                         "r"(smem_u32(sm)), "l"((uint64_t)&map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::
            "r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int n = threadIdx.x; n < bw * bh; n += blockDim.x) out[n] = sm[n];
}

int main(int argc, char **argv)
{
    const int W = 44, H = 40, Na = 5, pitch = 44;
    const int bw = argc > 1 ? atoi(argv[1]) : 40, bh = argc > 2 ? atoi(argv[2]) : 22;
    const int c0 = argc > 3 ? atoi(argv[3]) : -5, c1 = argc > 4 ? atoi(argv[4]) : -5, c2 = 1;
    const int mode = argc > 5 ? atoi(argv[5]) : 3;
    std::vector<float> h((size_t)Na * H * pitch);
    for (size_t n = 0; n < h.size(); n++) h[n] = (float)n;
    float *d, *out;
    CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&out, bw * bh * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Na};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d box=%dx%d coords=(%d,%d,%d) mode=%d\n", (int)r, bw, bh, c0, c1, c2, mode);
    probe<<<1, 128, bw * bh * 4>>>(map, out, bw, bh, c0, c1, c2, mode);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(bw * bh);
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++) {
            const int gy = c1 + y, gx = c0 + x;
            const float want = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[((size_t)c2 * H + gy) * pitch + gx] : 0.f;
            if (o[y * bw + x] != want) bad++;
        }
    printf("mismatches: %d of %d\n", bad, bw * bh);
    return 0;
}
