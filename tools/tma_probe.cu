// tma_probe.cu -- which box shapes does cp.async.bulk.tensor accept on this GPU?  (development aid)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, unsigned bytes, float *out, int n) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 65536);
    unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(d), "l"(&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main(int argc, char **argv) {
    // args: elemsize(4|8) NPX NPY NPZ boxw boxh rank(3|4)
    const int es_ = argc > 1 ? atoi(argv[1]) : 4, NPX = argc > 2 ? atoi(argv[2]) : 40, NPY = argc > 3 ? atoi(argv[3]) : 68,
              NPZ = argc > 4 ? atoi(argv[4]) : 36, w = argc > 5 ? atoi(argv[5]) : 68, rows = argc > 6 ? atoi(argv[6]) : 10,
              rank = argc > 7 ? atoi(argv[7]) : 4;
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    size_t bytes = (size_t)NPX * NPY * NPZ * 3 * es_;
    void *d;
    float *out;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 0, bytes);
    cudaMalloc(&out, 65536);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    cuuint64_t dims[4] = {(cuuint64_t)NPX, (cuuint64_t)NPY, (cuuint64_t)NPZ, 3};
    cuuint64_t str[3] = {(cuuint64_t)NPX * es_, (cuuint64_t)NPX * NPY * es_, (cuuint64_t)NPX * NPY * NPZ * es_};
    cuuint32_t box[4] = {(cuuint32_t)w, (cuuint32_t)rows, 1, 3}, es[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, es_ == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, d, dims, str, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("encode failed %d\n", (int)r);
        return 2;
    }
    const int n = w * rows * (rank == 4 ? 3 : 1);
    probe<<<1, 128, 65536 + 64>>>(tm, 2, 5, 1, 0, (unsigned)n * es_, out, 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("elem %d dims %dx%dx%d box %dx%d rank %d (inner %d bytes, tx %d): %s\n", es_, NPX, NPY, NPZ, w, rows, rank, w * es_, n * es_,
           cudaGetErrorString(e));
    return e != cudaSuccess;
}
