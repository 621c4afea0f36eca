// p2p_microbench.cu -- how fast can a kernel store to / load from a peer GPU over NVLink,
// as a function of the contiguous segment size of the access pattern?  Decides the layout of
// the transposes of the slab-decomposed FFT (poisson.cu): the strided FFT pass writes, per
// line element, one segment of `seg` bytes (lines-per-CTA x 8 B) into the destination.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_microbench tools/p2p_microbench.cu
// run  : tools/p2p_microbench            (needs 2 GPUs with peer access, single process)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__);  \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// One CTA = one tile column `c` of one plane: for every row r it moves SEG bytes at
//   plane * rows * rowbytes + r * rowbytes + c * SEG,   rowbytes = ntile * SEG.
// 1024 threads, float2 per thread and iteration, rows visited as r = t + P * m like the FFT
// kernel does (P = rows / 8).  src is read with the same pattern (local), dst may be a peer.
template <int SEG>
__global__ void __launch_bounds__(1024) seg_copy(const float2 *__restrict__ src, float2 *__restrict__ dst, int rows,
                                                 int ntile) {
    constexpr int E = SEG / 8;  // float2 per segment
    const int l = threadIdx.x % E, t = threadIdx.x / E;
    const int P = 1024 / E;  // rows covered per iteration
    const int c = blockIdx.x, plane = blockIdx.y;
    const long long rowe = (long long)ntile * E;
    const long long base = (long long)plane * rows * rowe + (long long)c * E + l;
    float2 v[8];
    const int iters = rows / P;  // 8 when rows * E == 8192
#pragma unroll
    for (int m = 0; m < 8; m++)
        if (m < iters) v[m] = src[base + (long long)(t + P * m) * rowe];
#pragma unroll
    for (int m = 0; m < 8; m++)
        if (m < iters) {
            v[m].x += 1.0f;
            dst[base + (long long)(t + P * m) * rowe] = v[m];
        }
}

__global__ void stream_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, long long n4) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i + 3 * step < n4; i += 4 * step) {
        float4 a = src[i], b = src[i + step], c = src[i + 2 * step], d = src[i + 3 * step];
        dst[i] = a;
        dst[i + step] = b;
        dst[i + 2 * step] = c;
        dst[i + 3 * step] = d;
    }
    for (; i < n4; i += step) dst[i] = src[i];
}

// bulk (TMA) variant: a CTA stages CH-byte chunks in shared memory with cp.async.bulk (local load, mbarrier)
// and stores them with cp.async.bulk.global.shared::cta in runs of RUN bytes (dst may be a peer); two buffers.
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int CH>
__global__ void __launch_bounds__(32) bulk_copy(const char *__restrict__ src, char *__restrict__ dst, long long nchunk, int run) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar[2];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + b)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    long long c = blockIdx.x;
    int it = 0;
    auto load = [&](long long cc, int b) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + b)), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + b * CH)),
                     "l"(src + cc * CH), "r"(CH), "r"(s32(bar + b))
                     : "memory");
    };
    if (c < nchunk) load(c, 0);
    for (; c < nchunk; c += gridDim.x, it++) {
        const int b = it & 1;
        const long long nx = c + gridDim.x;
        if (nx < nchunk) {
            // buffer b^1 was stored from in iteration it - 1: wait until that store has read it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            load(nx, b ^ 1);
        }
        unsigned ok = 0;
        const unsigned par = (it >> 1) & 1;
        while (!ok)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok)
                         : "r"(s32(bar + b)), "r"(par)
                         : "memory");
        for (int o = 0; o < CH; o += run)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * CH + o), "r"(s32(sm + b * CH + o)),
                         "r"(run)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

#include <functional>
static float time_ms(cudaStream_t st, int reps, const std::function<void()> &f) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) {
        printf("needs 2 GPUs, found %d\n", nd);
        return 0;
    }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    printf("devices %d, peer access 0->1: %d\n", nd, can);
    const size_t bytes = (size_t)512 << 20;
    float2 *a0, *b0, *a1;
    CK(cudaSetDevice(1));
    CK(cudaMalloc(&a1, bytes));
    CK(cudaMemset(a1, 0, bytes));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    CK(cudaMalloc(&a0, bytes));
    CK(cudaMalloc(&b0, bytes));
    CK(cudaMemset(a0, 0, bytes));
    CK(cudaMemset(b0, 0, bytes));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const int reps = 5;
    // --- segment patterns: bytes per CTA fixed at 64 KiB
    auto run_seg = [&](int seg, const float2 *src, float2 *dst, const char *what) {
        int rows = 65536 / seg;   // 1024 rows of 64 B, 512 of 128 B, 256 of 256 B ...
        int ntile = 65;           // like m0p / 8 = 65 tiles at n = 1024
        int planes = (int)(bytes / ((size_t)rows * ntile * seg));
        dim3 grid(ntile, planes);
        size_t moved = (size_t)planes * rows * ntile * seg;
        float ms = time_ms(st, reps, [&] {
            switch (seg) {
                case 64: seg_copy<64><<<grid, 1024, 0, st>>>(src, dst, rows, ntile); break;
                case 128: seg_copy<128><<<grid, 1024, 0, st>>>(src, dst, rows, ntile); break;
                case 256: seg_copy<256><<<grid, 1024, 0, st>>>(src, dst, rows, ntile); break;
                case 512: seg_copy<512><<<grid, 1024, 0, st>>>(src, dst, rows, ntile); break;
            }
        });
        printf("seg %4d B  %-28s %8.3f ms  %8.1f GB/s (one direction, %zu MB)\n", seg, what, ms,
               moved / ms * 1e-6, moved >> 20);
    };
    for (int seg : {64, 128, 256, 512}) {
        run_seg(seg, a0, b0, "local -> local");
        run_seg(seg, a0, a1, "local -> peer (store)");
        run_seg(seg, a1, b0, "peer -> local (load)");
    }
    // --- plain streaming copies
    for (int blocks : {37, 74, 148, 296, 592, 1184}) {
        float ms = time_ms(st, reps, [&] {
            stream_copy<<<blocks, 512, 0, st>>>((const float4 *)a0, (float4 *)a1, (long long)(bytes / 16));
        });
        float ms2 = time_ms(st, reps, [&] {
            stream_copy<<<blocks, 512, 0, st>>>((const float4 *)a1, (float4 *)b0, (long long)(bytes / 16));
        });
        printf("stream copy %5d CTAs x 512: store to peer %8.1f GB/s, load from peer %8.1f GB/s\n", blocks,
               bytes / ms * 1e-6, bytes / ms2 * 1e-6);
    }
    {
        float ms = time_ms(st, reps, [&] { CK(cudaMemcpyAsync(a1, a0, bytes, cudaMemcpyDeviceToDevice, st)); });
        printf("cudaMemcpyAsync local -> peer (512 MB): %8.1f GB/s\n", bytes / ms * 1e-6);
        const size_t chunk = (size_t)16 << 20;
        ms = time_ms(st, reps, [&] {
            for (size_t o = 0; o < bytes; o += chunk)
                CK(cudaMemcpyAsync((char *)a1 + o, (char *)a0 + o, chunk, cudaMemcpyDeviceToDevice, st));
        });
        printf("cudaMemcpyAsync local -> peer (32 x 16 MB): %8.1f GB/s\n", bytes / ms * 1e-6);
        ms = time_ms(st, reps, [&] { CK(cudaMemcpyAsync(b0, a0, bytes, cudaMemcpyDeviceToDevice, st)); });
        printf("cudaMemcpyAsync local -> local (512 MB): %8.1f GB/s (read + write %8.1f)\n", bytes / ms * 1e-6,
               2 * bytes / ms * 1e-6);
    }
    // --- bulk (TMA) copies: 16 KB chunks staged in shared memory, stored in runs of 256 B .. 16 KB
    {
        constexpr int CH = 16384;
        CK(cudaFuncSetAttribute(bulk_copy<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CH));
        for (int ctas : {148, 296, 592, 888})
            for (int run : {256, 1024, 4096, 16384}) {
                float ms = time_ms(st, reps, [&] {
                    bulk_copy<CH><<<ctas, 32, 2 * CH, st>>>((const char *)a0, (char *)a1, (long long)(bytes / CH), run);
                });
                float ms2 = time_ms(st, reps, [&] {
                    bulk_copy<CH><<<ctas, 32, 2 * CH, st>>>((const char *)a0, (char *)b0, (long long)(bytes / CH), run);
                });
                printf("bulk copy %4d CTAs, runs of %5d B: store to peer %8.1f GB/s, local %8.1f GB/s\n", ctas, run,
                       bytes / ms * 1e-6, bytes / ms2 * 1e-6);
            }
    }
    // --- bidirectional: both GPUs store to each other at the same time (segment 64 / 256 and stream)
    {
        cudaStream_t st1;
        float2 *b1;
        CK(cudaSetDevice(1));
        CK(cudaStreamCreate(&st1));
        CK(cudaMalloc(&b1, bytes));
        CK(cudaMemset(b1, 0, bytes));
        CK(cudaSetDevice(0));
        for (int seg : {64, 256}) {
            int rows = 65536 / seg, ntile = 65;
            int planes = (int)(bytes / ((size_t)rows * ntile * seg));
            dim3 grid(ntile, planes);
            size_t moved = (size_t)planes * rows * ntile * seg;
            float ms = time_ms(st, reps, [&] {
                CK(cudaSetDevice(1));
                if (seg == 64)
                    seg_copy<64><<<grid, 1024, 0, st1>>>(b1, b0, rows, ntile);
                else
                    seg_copy<256><<<grid, 1024, 0, st1>>>(b1, b0, rows, ntile);
                CK(cudaSetDevice(0));
                if (seg == 64)
                    seg_copy<64><<<grid, 1024, 0, st>>>(a0, a1, rows, ntile);
                else
                    seg_copy<256><<<grid, 1024, 0, st>>>(a0, a1, rows, ntile);
            });
            CK(cudaSetDevice(1));
            CK(cudaStreamSynchronize(st1));
            CK(cudaSetDevice(0));
            printf("bidirectional seg %4d B stores: %8.1f GB/s per direction (timed on GPU 0)\n", seg,
                   moved / ms * 1e-6);
        }
    }
    return 0;
}
