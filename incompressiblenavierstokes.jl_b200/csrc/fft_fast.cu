// fft_fast.cu -- register/shared-memory FFT passes for power-of-two lengths (32..2048).
//
// One thread owns 8 complex values of a line of length N (P = N/8 threads per line) and runs
// a Stockham auto-sort FFT whose radix-8/4/2 butterflies live entirely in registers; values
// are exchanged between stages through padded, double-buffered shared memory holding whole
// complex numbers (one 64-/128-bit access per value).  In EVERY stage register m of thread t
// holds element t + P*m, and so does the final result -- which is also the input distribution
// of the next transform.  The middle pass of the Poisson solve exploits that: forward FFT,
// eigenvalue division (pressure.jl:360-372) and inverse FFT run back to back without leaving
// registers in between.  Float32 arithmetic is done on (re, im) pairs with the packed FP32
// instructions of sm_100 (FADD2/FMUL2/FFMA2), which halves the issue slots of the butterflies.
//
// Pass structure of psolver_spectral (pressure.jl:349-381) with these kernels, 3-D:
//   x: two real rows -> one complex FFT -> two half spectra (R2C), reads view(p, Ip) directly
//   y: strided forward FFT        z: strided forward FFT + divide + inverse FFT
//   y: strided inverse FFT        x: two half spectra -> two real rows (C2R), writes view(p, Ip)
// i.e. 5 passes of one read + one write each (10 w per cell per solve).
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "cx.cuh"
#include "fft_fast.h"

namespace insb {

__host__ __device__ constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n / 2); }
template <int N>
struct FftShape {
    static constexpr int E = ilog2c(N);
    static constexpr int REM = E % 3;
    static constexpr int NST = E / 3 + (REM ? 1 : 0);
    static constexpr int P = N / 8;
    __host__ __device__ static constexpr int radix(int s) { return (REM != 0 && s == 0) ? (1 << REM) : 8; }
    __host__ __device__ static constexpr int ns_before(int s) {
        int ns = 1;
        for (int i = 0; i < s; i++) ns *= radix(i);
        return ns;
    }
};

// padded shared-memory position of element p of a line: one pad element per 2^SH elements.
// SH = 3 is conflict-free when lanes run over lines (strided passes) and for 16-byte elements;
// 8-byte elements with lanes running along the line (x passes, Float32) want SH = 4
// (simulated bank conflicts per access 1.25 instead of 1.5-1.8).
template <int SH>
__host__ __device__ __forceinline__ constexpr int spx(int p) { return p + (p >> SH); }
__host__ __device__ __forceinline__ constexpr int sp(int p) { return spx<3>(p); }

// ----------------------------------------------------------------------------------------
// complex arithmetic: generic (double) and packed (float)
// ----------------------------------------------------------------------------------------
template <typename T>
struct Ops {
    typedef Cx<T> C;
    static __device__ __forceinline__ C add(C a, C b) { return C{a.x + b.x, a.y + b.y}; }
    static __device__ __forceinline__ C sub(C a, C b) { return C{a.x - b.x, a.y - b.y}; }
    // a + s*i*b, s = -1 (forward transform) or +1 (inverse)
    template <bool INV>
    static __device__ __forceinline__ C add_i(C a, C b) {
        return INV ? C{a.x - b.y, a.y + b.x} : C{a.x + b.y, a.y - b.x};
    }
    template <bool INV>
    static __device__ __forceinline__ C sub_i(C a, C b) {
        return INV ? C{a.x + b.y, a.y - b.x} : C{a.x - b.y, a.y + b.x};
    }
    static __device__ __forceinline__ C mul(C a, C w) { return C{a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
    static __device__ __forceinline__ C axpy(C a, T s, C b) { return C{fma(a.x, s, b.x), fma(a.y, s, b.y)}; }  // b + s a
    static __device__ __forceinline__ C scale(C a, T s) { return C{a.x * s, a.y * s}; }
};

template <>
struct Ops<float> {
    typedef Cx<float> C;
    static __device__ __forceinline__ float2 f(C a) { return make_float2(a.x, a.y); }
    static __device__ __forceinline__ C c(float2 a) { return C{a.x, a.y}; }
    static __device__ __forceinline__ C add(C a, C b) { return c(__fadd2_rn(f(a), f(b))); }
    static __device__ __forceinline__ C sub(C a, C b) { return c(__ffma2_rn(f(b), make_float2(-1.f, -1.f), f(a))); }
    template <bool INV>
    static __device__ __forceinline__ C add_i(C a, C b) {
        return c(__ffma2_rn(make_float2(b.y, b.x), INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f), f(a)));
    }
    template <bool INV>
    static __device__ __forceinline__ C sub_i(C a, C b) {
        return c(__ffma2_rn(make_float2(b.y, b.x), INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f), f(a)));
    }
    static __device__ __forceinline__ C mul(C a, C w) {
        float2 t = __fmul2_rn(make_float2(-a.y, a.y), make_float2(w.y, w.x));
        return c(__ffma2_rn(make_float2(a.x, a.x), f(w), t));
    }
    static __device__ __forceinline__ C axpy(C a, float s, C b) { return c(__ffma2_rn(f(a), make_float2(s, s), f(b))); }
    static __device__ __forceinline__ C scale(C a, float s) { return c(__fmul2_rn(f(a), make_float2(s, s))); }
};

// factor / den of the eigenvalue division: Float32 takes the fast reciprocal (two instructions, <= 2 ulp, no
// slow-path call per element); Float64 keeps the correctly rounded quotient
__device__ __forceinline__ float fast_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double fast_div(double a, double b) { return a / b; }

// radix-R butterfly on v[b + B*r], r = 0..R-1, natural-order output
template <typename T, int R, int B, bool INV>
__device__ __forceinline__ void butterfly(Cx<T> (&v)[8], int b) {
    typedef Ops<T> O;
    if constexpr (R == 2) {
        Cx<T> a = v[b], c = v[b + B];
        v[b] = O::add(a, c);
        v[b + B] = O::sub(a, c);
    } else if constexpr (R == 4) {
        Cx<T> a0 = O::add(v[b], v[b + 2 * B]), a1 = O::sub(v[b], v[b + 2 * B]);
        Cx<T> a2 = O::add(v[b + B], v[b + 3 * B]), a3 = O::sub(v[b + B], v[b + 3 * B]);
        v[b] = O::add(a0, a2);
        v[b + 2 * B] = O::sub(a0, a2);
        v[b + B] = O::template add_i<INV>(a1, a3);
        v[b + 3 * B] = O::template sub_i<INV>(a1, a3);
    } else {
        const T h = T(0.70710678118654752440);
        Cx<T> a0 = O::add(v[0], v[4]), a1 = O::sub(v[0], v[4]);
        Cx<T> a2 = O::add(v[2], v[6]), a3 = O::sub(v[2], v[6]);
        Cx<T> a4 = O::add(v[1], v[5]), a5 = O::sub(v[1], v[5]);
        Cx<T> a6 = O::add(v[3], v[7]), a7 = O::sub(v[3], v[7]);
        Cx<T> b0 = O::add(a0, a2), b2 = O::sub(a0, a2);
        Cx<T> b1 = O::template add_i<INV>(a1, a3), b3 = O::template sub_i<INV>(a1, a3);
        Cx<T> b4 = O::add(a4, a6), b6 = O::sub(a4, a6);
        Cx<T> b5 = O::template add_i<INV>(a5, a7), b7 = O::template sub_i<INV>(a5, a7);
        // w8^1 b5 = h (b5 + s i b5),  w8^3 b7 = -h (b7 - s i b7),  w8^2 b6 = s i b6
        Cx<T> r5 = O::template add_i<INV>(b5, b5), r7 = O::template sub_i<INV>(b7, b7);
        v[0] = O::add(b0, b4);
        v[4] = O::sub(b0, b4);
        v[1] = O::axpy(r5, h, b1);
        v[5] = O::axpy(r5, -h, b1);
        v[2] = O::template add_i<INV>(b2, b6);
        v[6] = O::template sub_i<INV>(b2, b6);
        v[3] = O::axpy(r7, -h, b3);
        v[7] = O::axpy(r7, h, b3);
    }
}

// Per-thread view of the exchange buffers: element p of this thread's line sits at
// buf[k][sp(p) * ES] (ES = lines per block when the lines are interleaved, 1 otherwise).
template <typename T>
struct LineSmem {
    Cx<T> *buf[2];
};

// Barrier of the threads that share a line's exchange buffer.  Strided passes (ES > 1): the lines of a block are
// interleaved over all its threads -- the whole block.  x passes (ES == 1): line lp is the P consecutive threads
// [lp P, (lp + 1) P) with their own buffer, so the lines of a block need not wait for one another: a named barrier
// per line (bar.sync id, P), or the warp barrier when a line lives inside one warp.
template <int ES, int P>
__device__ __forceinline__ void line_barrier() {
    if constexpr (ES > 1) {
        __syncthreads();
    } else if constexpr (P <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x / P)), "n"(P) : "memory");
    }
}

// TW7: all seven radix-8 twiddles from the table (strided passes) instead of three loads + four products
template <typename T, int N, bool INV, int S, int ES, int SH = 3, bool TW7 = (ES > 1)>
__device__ __forceinline__ void fft_stages(Cx<T> (&v)[8], int t, const LineSmem<T> &sm,
                                           const Cx<T> *__restrict__ tw, int &exch) {
    typedef Ops<T> O;
    using Sh = FftShape<N>;
    constexpr int R = Sh::radix(S);
    constexpr int B = 8 / R;
    constexpr int P = Sh::P;
    constexpr int NS = Sh::ns_before(S);
    // twiddles exp(-+2 pi i r k / (NS R)), k = j mod NS
    if constexpr (NS > 1) {
        constexpr int TSTEP = N / (NS * R);
#pragma unroll
        for (int b = 0; b < B; b++) {
            int k = (t + P * b) & (NS - 1);
            if constexpr (R == 2) {
                Cx<T> w1 = tw[k * TSTEP];
                if (INV) w1 = cconj(w1);
                v[b + B] = O::mul(v[b + B], w1);
            } else if constexpr (R == 4) {
                Cx<T> w1 = tw[k * TSTEP], w2 = tw[2 * k * TSTEP];
                if (INV) w1 = cconj(w1), w2 = cconj(w2);
                Cx<T> w3 = O::mul(w1, w2);
                v[b + B] = O::mul(v[b + B], w1);
                v[b + 2 * B] = O::mul(v[b + 2 * B], w2);
                v[b + 3 * B] = O::mul(v[b + 3 * B], w3);
            } else {
                Cx<T> w1, w2, w3, w4, w5, w6, w7;
                if constexpr (TW7) {
                    // strided passes (the lanes of a warp run over lines: k takes four values per warp, the loads
                    // are broadcasts): all seven powers straight from the table (r k TSTEP < N) -- four more L1
                    // hits instead of four complex products, every twiddle the correctly rounded table entry
                    const Cx<T> *twk = tw + k * TSTEP;
                    w1 = twk[0], w2 = twk[k * TSTEP], w3 = twk[2 * k * TSTEP], w4 = twk[3 * k * TSTEP];
                    w5 = twk[4 * k * TSTEP], w6 = twk[5 * k * TSTEP], w7 = twk[6 * k * TSTEP];
                    if (INV) {
                        w1 = cconj(w1), w2 = cconj(w2), w3 = cconj(w3), w4 = cconj(w4);
                        w5 = cconj(w5), w6 = cconj(w6), w7 = cconj(w7);
                    }
                } else {
                    // x passes (lanes along the line: 32 different k per warp): three loads, four products
                    w1 = tw[k * TSTEP], w2 = tw[2 * k * TSTEP], w4 = tw[4 * k * TSTEP];
                    if (INV) w1 = cconj(w1), w2 = cconj(w2), w4 = cconj(w4);
                    w3 = O::mul(w1, w2), w5 = O::mul(w1, w4), w6 = O::mul(w2, w4);
                    w7 = O::mul(w3, w4);
                }
                v[1] = O::mul(v[1], w1);
                v[2] = O::mul(v[2], w2);
                v[3] = O::mul(v[3], w3);
                v[4] = O::mul(v[4], w4);
                v[5] = O::mul(v[5], w5);
                v[6] = O::mul(v[6], w6);
                v[7] = O::mul(v[7], w7);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < B; b++) butterfly<T, R, B, INV>(v, b);
    if constexpr (S < Sh::NST - 1) {
        Cx<T> *buf = sm.buf[exch & 1];
#pragma unroll
        for (int b = 0; b < B; b++) {
            int j = t + P * b;
            int j0 = (j / NS) * (NS * R) + (j & (NS - 1));
            constexpr int PADN = 1 << SH;
            if constexpr (NS % PADN == 0 || (NS == 1 && R == 8)) {
                // sp(j0 + r NS) = sp(j0) + r (NS + NS/PADN)   (r < 8 when NS == 1, j0 a multiple of 8)
                constexpr int DS = NS == 1 ? 1 : NS + NS / PADN;
                Cx<T> *w = buf + spx<SH>(j0) * ES;
#pragma unroll
                for (int r = 0; r < R; r++) w[r * DS * ES] = v[b + B * r];
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) buf[spx<SH>(j0 + r * NS) * ES] = v[b + B * r];
            }
        }
        line_barrier<ES, P>();
        if constexpr (P % (1 << SH) == 0) {
            const Cx<T> *rd = buf + spx<SH>(t) * ES;
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = rd[m * (P + P / (1 << SH)) * ES];
        } else {
#pragma unroll
            for (int m = 0; m < 8; m++) v[m] = buf[spx<SH>(t + P * m) * ES];
        }
        exch++;
        fft_stages<T, N, INV, S + 1, ES, SH, TW7>(v, t, sm, tw, exch);
    }
}

// lines per block of the strided passes: 64-byte row segments in Float32/Float64 while the
// block stays at <= 512 threads (1024 for the longest Float32 lines)
// (narrower tiles -- 32-byte segments, twice the blocks per SM -- were measured: the middle pass
// drops from 2.6 to 1.9 TB/s at 512^3 Float32, so 64-byte segments stay)
template <typename T, int N>
struct StridedShape {
    static constexpr int P = FftShape<N>::P;
    static constexpr int L = sizeof(T) == 4 ? (N <= 1024 ? 8 : 4) : (N <= 1024 ? 4 : 2);
    static constexpr int NT = L * P;
    static constexpr int BUF = sp(N - 1) + 1;  // complex elements per line and buffer
    static constexpr size_t smem = (size_t)2 * L * BUF * sizeof(Cx<T>);
};

// row-grouped transpose layout (fft_fast.h): offset of row j / of column kx inside a plane
template <typename T>
__device__ __forceinline__ long long grp_row(int j, int m0p) {
    constexpr int U = 64 / (int)sizeof(Cx<T>);
    return (long long)(j >> 2) * (m0p * 4) + (j & 3) * U;
}
template <typename T>
__device__ __forceinline__ int grp_x(int kx) {
    constexpr int U = 64 / (int)sizeof(Cx<T>);
    return (kx / U) * (4 * U) + (kx % U);
}

// ----------------------------------------------------------------------------------------
// strided pass along dim `tdim` (1 or 2): MODE 0 forward, 1 inverse, 2 forward+divide+inverse
// grid: x = tiles of L lines along kx, y = index along the remaining dimension
// BLK: generalised (blocked) addressing of the slab-decomposed solve, see fft_fast.h
// ----------------------------------------------------------------------------------------
template <typename T, int N, int MODE, bool BLK>
__global__ void __launch_bounds__(StridedShape<T, N>::NT, (sizeof(T) == 4 && StridedShape<T, N>::NT <= 512) ? 2 : 1)
    fft_strided_kernel(Cx<T> *__restrict__ W, FastGeom g, int tdim, const Cx<T> *__restrict__ tw,
                       const T *__restrict__ e0, const T *__restrict__ e1, const T *__restrict__ e2, T factor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef StridedShape<T, N> SS;
    constexpr int P = SS::P, L = SS::L;
    const int l = threadIdx.x % L, t = threadIdx.x / L;
    const int kx = blockIdx.x * L + l;
    const int other = blockIdx.y;
    const int odim = tdim == 1 ? 2 : 1;
    const long long in_sd = g.in_sd >= 0 ? g.in_sd : g.ws[tdim], out_sd = g.out_sd >= 0 ? g.out_sd : g.ws[tdim];
    const long long in_os = g.in_os >= 0 ? g.in_os : g.ws[odim], out_os = g.out_os >= 0 ? g.out_os : g.ws[odim];
    // row-grouped transpose layouts (fft_fast.h): column kx sits at grp_x(kx) inside its row group
    const long long xin = (BLK && g.in_grp) ? grp_x<T>(kx) : kx, xout = (BLK && g.out_grp) ? grp_x<T>(kx) : kx;
    const long long oin = (BLK && g.in_grp == 2) ? grp_row<T>(other, g.m0p) : other * in_os;
    const long long oout = (BLK && g.out_grp == 2) ? grp_row<T>(other, g.m0p) : other * out_os;
    const Cx<T> *ibase = (g.in ? (const Cx<T> *)g.in : W) + xin + (g.D == 3 ? oin : 0);
    Cx<T> *obase = (g.out ? (Cx<T> *)g.out : W) + xout + (g.D == 3 ? oout : 0);
    LineSmem<T> sm;
    sm.buf[0] = reinterpret_cast<Cx<T> *>(smem_raw) + l;
    sm.buf[1] = sm.buf[0] + L * SS::BUF;
    Cx<T> v[8];
    if (BLK) {
        const int in_mask = (1 << g.in_sh) - 1;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int k = t + P * m;
            const int kl = k & in_mask;
            v[m] = ibase[(long long)(k >> g.in_sh) * g.in_blk +
                         (g.in_grp == 1 ? grp_row<T>(kl, g.m0p) : (long long)kl * in_sd)];
        }
    } else {
        const Cx<T> *ip = ibase + t * in_sd;
        const long long step = P * in_sd;
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = ip[m * step];
    }
    int exch = 0;
    // (the middle pass of the 1024-thread blocks spills at its 64-register cap with seven live twiddles: it keeps the
    // products.  The rule depends on (N, MODE) only -- the plain, the generalised and the TMA variant of a pass must
    // round alike, or P GPUs would not reproduce one GPU bit for bit.)
    constexpr bool TW7 = !(MODE == 2 && N > 512);
    if (MODE == 0 || MODE >= 2) fft_stages<T, N, false, 0, L, 3, TW7>(v, t, sm, tw, exch);
    if (MODE >= 2) {
        // phat = -phat / ((ax + ay) + az), phat[1] = 0 (pressure.jl:360-372); `factor` carries
        // the sign and the 1/prod(Np) of the inverse transform.
        const T e01 = g.D == 3 ? e0[kx] + (tdim == 1 ? T(0) : e1[other + g.eig_off]) : e0[kx];
        const T eo2 = g.D == 3 && tdim == 1 ? e2[other + g.eig_off] : T(0);
        const T *el = (g.D == 3 && tdim == 2) ? e2 : e1;
        const bool zero_rest = g.D == 3 ? (kx | (other + g.eig_off)) == 0 : kx == 0;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int k = t + P * m;
            // (ax + ay) + az in the reference's order: the line runs along y (tdim == 1) or z
            const T den = (g.D == 3 && tdim == 1) ? (e01 + el[k]) + eo2 : e01 + el[k];
            const T f = (zero_rest && k == 0) ? T(0) : fast_div(factor, den);
            v[m] = Ops<T>::scale(v[m], f);
        }
    }
    if (MODE >= 1) fft_stages<T, N, true, 0, L, 3, TW7>(v, t, sm, tw, exch);
    if (BLK) {
        const int out_mask = (1 << g.out_sh) - 1;
        if (g.use_outp) {  // tile-by-tile stores straight into the peers' buffers (NVLink P2P)
            const long long orel = xout + oout;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                int k = t + P * m;
                const int kl = k & out_mask;
                Cx<T> *dst = reinterpret_cast<Cx<T> *>(g.outp[k >> g.out_sh]);
                dst[orel + (g.out_grp == 1 ? grp_row<T>(kl, g.m0p) : (long long)kl * out_sd)] = v[m];
            }
        } else {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                int k = t + P * m;
                const int kl = k & out_mask;
                obase[(long long)(k >> g.out_sh) * g.out_blk +
                      (g.out_grp == 1 ? grp_row<T>(kl, g.m0p) : (long long)kl * out_sd)] = v[m];
            }
        }
    } else {
        Cx<T> *op = obase + t * out_sd;
        const long long step = P * out_sd;
#pragma unroll
        for (int m = 0; m < 8; m++) op[m * step] = v[m];
    }
}

// ----------------------------------------------------------------------------------------
// TMA-pipelined variant of the strided pass (3-D, in place on W): a persistent block walks over its tiles of L lines;
// the [N rows][64 bytes] tile arrives as cp.async.bulk.tensor boxes (<= 256 rows each) with mbarrier completion and is
// read into registers at once, after which the NEXT tile's boxes are issued -- they land while this tile's FFTs run;
// the result goes through shared memory and one bulk tensor store per box.  The plain kernel above keeps every byte in
// flight in registers, so a block alternates between waiting for HBM and computing (3.2 us per tile at 512^3 where the
// memory system needs 1.5 us and the issue slots 0.9 us); here the memory traffic of a block overlaps its own compute.
// Shared memory: input tile A, exchange buffers X and B (the store reads whichever the last exchange left free).
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned tma_s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <typename T, int N>
struct TmaStridedShape {
    typedef StridedShape<T, N> SS;
    static constexpr int ROWB = SS::L * (int)sizeof(Cx<T>);             // bytes per tile row (64)
    static constexpr int BOXR = N < 256 ? N : 256;                      // rows per TMA box
    static constexpr size_t tile_bytes = (size_t)N * ROWB;
    static constexpr size_t ex_bytes = ((size_t)SS::L * SS::BUF * sizeof(Cx<T>) + 127) / 128 * 128;
    static constexpr size_t smem = tile_bytes + 2 * ex_bytes + 1024;    // (+ alignment slack)
};

template <typename T, int N, int MODE>
__global__ void __launch_bounds__(StridedShape<T, N>::NT, (sizeof(T) == 4 && StridedShape<T, N>::NT <= 512) ? 2 : 1)
    fft_strided_tma_kernel(const __grid_constant__ CUtensorMap tm, FastGeom g, int tdim, int ntx, int ntiles,
                           const Cx<T> *__restrict__ tw, const T *__restrict__ e0, const T *__restrict__ e1,
                           const T *__restrict__ e2, T factor) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    typedef StridedShape<T, N> SS;
    typedef TmaStridedShape<T, N> TS;
    constexpr int P = SS::P, L = SS::L;
    __shared__ __align__(8) unsigned long long bar;
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    Cx<T> *A = reinterpret_cast<Cx<T> *>(base);
    Cx<T> *X = reinterpret_cast<Cx<T> *>(base + TS::tile_bytes);
    Cx<T> *B = reinterpret_cast<Cx<T> *>(base + TS::tile_bytes + TS::ex_bytes);
    const int tid = threadIdx.x, l = tid % L, t = tid / L;
    constexpr int FPR = TS::ROWB / 4;  // box extent along x in 4-byte words
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tma_s32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // box r of tile `tile`: rows r * BOXR .. of the line direction, column block c, companion index `other`
    auto coords = [&](int tile, int r, int &cx, int &cy, int &cz) {
        const int c = tile % ntx, other = tile / ntx;
        cx = c * FPR;
        cy = tdim == 1 ? r * TS::BOXR : other;
        cz = tdim == 1 ? other : r * TS::BOXR;
    };
    auto issue_load = [&](int tile) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_s32(&bar)), "r"((unsigned)TS::tile_bytes)
                     : "memory");
#pragma unroll
        for (int r = 0; r < N / TS::BOXR; r++) {
            int cx, cy, cz;
            coords(tile, r, cx, cy, cz);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(tma_s32(reinterpret_cast<unsigned char *>(A) + (size_t)r * TS::BOXR * TS::ROWB)), "l"(&tm), "r"(tma_s32(&bar)),
                "r"(cx), "r"(cy), "r"(cz)
                : "memory");
        }
    };
    const int first = blockIdx.x;
    if (tid == 0 && first < ntiles) issue_load(first);
    int it = 0;
    for (int tile = first; tile < ntiles; tile += gridDim.x, it++) {
        const int c = tile % ntx, other = tile / ntx;
        const int kx = c * L + l;
        {  // wait for this tile's boxes
            unsigned ok = 0, spins = 0;
            const unsigned par = it & 1;
            while (!ok) {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok)
                             : "r"(tma_s32(&bar)), "r"(par)
                             : "memory");
                if (!ok && ++spins > (1u << 22)) __trap();  // a transfer that never completes must fail loudly, not hang
            }
        }
        Cx<T> v[8];
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = A[(t + P * m) * L + l];
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the last store has read its buffer
        __syncthreads();  // A is consumed, X / B are free
        if (tid == 0 && tile + (int)gridDim.x < ntiles) issue_load(tile + gridDim.x);
        LineSmem<T> sm;
        sm.buf[0] = X + l;
        sm.buf[1] = B + l;
        int exch = 0;
        constexpr bool TW7 = !(MODE == 2 && N > 512);  // (64-register cap of the 1024-thread blocks)
        if (MODE == 0 || MODE >= 2) fft_stages<T, N, false, 0, L, 3, TW7>(v, t, sm, tw, exch);
        if (MODE >= 2) {
            const T e01 = e0[kx] + (tdim == 1 ? T(0) : e1[other + g.eig_off]);
            const T eo2 = tdim == 1 ? e2[other + g.eig_off] : T(0);
            const T *el = tdim == 2 ? e2 : e1;
            const bool zero_rest = (kx | (other + g.eig_off)) == 0;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int k = t + P * m;
                const T den = tdim == 1 ? (e01 + el[k]) + eo2 : e01 + el[k];
                const T f = (zero_rest && k == 0) ? T(0) : fast_div(factor, den);
                v[m] = Ops<T>::scale(v[m], f);
            }
        }
        if (MODE >= 1) fft_stages<T, N, true, 0, L, 3, TW7>(v, t, sm, tw, exch);
        Cx<T> *O = (exch & 1) ? B : X;  // the buffer the last exchange did not use
#pragma unroll
        for (int m = 0; m < 8; m++) O[(t + P * m) * L + l] = v[m];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int r = 0; r < N / TS::BOXR; r++) {
                int cx, cy, cz;
                coords(tile, r, cx, cy, cz);
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tm),
                             "r"(tma_s32(reinterpret_cast<unsigned char *>(O) + (size_t)r * TS::BOXR * TS::ROWB)), "r"(cx), "r"(cy),
                             "r"(cz)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// strided DCT pass along dim `tdim` for psolver_transform (pressure.jl:255-317): orthonormal
// DCT-II / DCT-III of the real and imaginary parts of every line through ONE complex FFT of the
// same length (the even/odd permutation + twiddle construction of manual_dct!/manual_idct!,
// src/utils.jl:89-155).  MODE 0 forward, 1 inverse, 2 forward + eigenvalue divide + inverse.
// ----------------------------------------------------------------------------------------
// v[m] <-> element k = t + P m; returns in vr[m] element (N - k) mod N
template <typename T, int N, int ES>
__device__ __forceinline__ void exchange_rev(const Cx<T> (&v)[8], Cx<T> (&vr)[8], int t, const LineSmem<T> &sm,
                                             int &exch) {
    constexpr int P = FftShape<N>::P;
    Cx<T> *buf = sm.buf[exch & 1];
#pragma unroll
    for (int m = 0; m < 8; m++) buf[sp(t + P * m) * ES] = v[m];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; m++) vr[m] = buf[sp((N - (t + P * m)) & (N - 1)) * ES];
    exch++;
}

// BLK: the generalised (blocked / grouped / peer-store) addressing of the slab-decomposed solve, exactly the one of
// fft_strided_kernel (fft_fast.h); element k of the line is read at in_at(k) and written at out_at(k).
template <typename T, int N, int MODE, bool BLK>
__global__ void __launch_bounds__(StridedShape<T, N>::NT, (sizeof(T) == 4 && StridedShape<T, N>::NT <= 512) ? 2 : 1)
    fft_strided_dct_kernel(Cx<T> *__restrict__ W, FastGeom g, int tdim, const Cx<T> *__restrict__ tw,
                           const T *__restrict__ e0, const T *__restrict__ e1, const T *__restrict__ e2, T factor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef StridedShape<T, N> SS;
    constexpr int P = SS::P, L = SS::L;
    const int l = threadIdx.x % L, t = threadIdx.x / L;
    const int kx = blockIdx.x * L + l;
    const int other = blockIdx.y;
    const int odim = tdim == 1 ? 2 : 1;
    const long long sd = g.ws[tdim];
    Cx<T> *base = W + kx + (g.D == 3 ? other * g.ws[odim] : 0);
    const long long in_sd = g.in_sd >= 0 ? g.in_sd : g.ws[tdim], out_sd = g.out_sd >= 0 ? g.out_sd : g.ws[tdim];
    const long long in_os = g.in_os >= 0 ? g.in_os : g.ws[odim], out_os = g.out_os >= 0 ? g.out_os : g.ws[odim];
    const long long xin = (BLK && g.in_grp) ? grp_x<T>(kx) : kx, xout = (BLK && g.out_grp) ? grp_x<T>(kx) : kx;
    const long long oin = (BLK && g.in_grp == 2) ? grp_row<T>(other, g.m0p) : other * in_os;
    const long long oout = (BLK && g.out_grp == 2) ? grp_row<T>(other, g.m0p) : other * out_os;
    const Cx<T> *ibase = (g.in ? (const Cx<T> *)g.in : W) + xin + (g.D == 3 ? oin : 0);
    Cx<T> *obase = (g.out ? (Cx<T> *)g.out : W) + xout + (g.D == 3 ? oout : 0);
    auto in_at = [&](int k) -> const Cx<T> * {
        if (!BLK) return base + (long long)k * sd;
        const int kl = k & ((1 << g.in_sh) - 1);
        return ibase + (long long)(k >> g.in_sh) * g.in_blk + (g.in_grp == 1 ? grp_row<T>(kl, g.m0p) : (long long)kl * in_sd);
    };
    auto out_at = [&](int k) -> Cx<T> * {
        if (!BLK) return base + (long long)k * sd;
        const int kl = k & ((1 << g.out_sh) - 1);
        const long long rel = g.out_grp == 1 ? grp_row<T>(kl, g.m0p) : (long long)kl * out_sd;
        if (g.use_outp) return reinterpret_cast<Cx<T> *>(g.outp[k >> g.out_sh]) + xout + oout + rel;
        return obase + (long long)(k >> g.out_sh) * g.out_blk + rel;
    };
    const Cx<T> *dw = (const Cx<T> *)g.dw, *dwi = (const Cx<T> *)g.dwi;
    LineSmem<T> sm;
    sm.buf[0] = reinterpret_cast<Cx<T> *>(smem_raw) + l;
    sm.buf[1] = sm.buf[0] + L * SS::BUF;
    Cx<T> v[8], vr[8];
    // position in the line of element p of the FFT sequence: z[p] = x[2p], z[N-1-p] = x[2p+1]
    auto perm = [](int p) { return p < N / 2 ? 2 * p : 2 * (N - 1 - p) + 1; };
    if (MODE == 1) {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = *in_at(t + P * m);
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = *in_at(perm(t + P * m));
    }
    T f[8];
    if (MODE == 2) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int k = t + P * m;
            T den;
            bool zero;
            if (g.D == 3) {
                int k1 = tdim == 1 ? k : other + g.eig_off, k2 = tdim == 1 ? other + g.eig_off : k;
                den = (e0[kx] + e1[k1]) + e2[k2];
                zero = (kx | k1 | k2) == 0;
            } else {
                den = e0[kx] + e1[k];
                zero = (kx | k) == 0;
            }
            f[m] = zero ? T(0) : factor / den;
        }
    }
    int exch = 0;
    if (MODE != 1) {  // DCT-II: FFT, then X_k = 2 Re(w_k F_k) for the real and the imaginary sequence
        fft_stages<T, N, false, 0, L, 3, MODE != 2>(v, t, sm, tw, exch);
        exchange_rev<T, N, L>(v, vr, t, sm, exch);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const Cx<T> zk = v[m], zc = cconj(vr[m]), wk = dw[t + P * m];
            const Cx<T> fa{(zk.x + zc.x) * T(0.5), (zk.y + zc.y) * T(0.5)};
            const Cx<T> d{(zk.x - zc.x) * T(0.5), (zk.y - zc.y) * T(0.5)};
            const Cx<T> fb{d.y, -d.x};
            v[m] = Cx<T>{T(2) * (wk.x * fa.x - wk.y * fa.y), T(2) * (wk.x * fb.x - wk.y * fb.y)};
        }
    }
    if (MODE == 2) {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = Ops<T>::scale(v[m], f[m]);
    }
    if (MODE >= 1) {  // DCT-III: H_k from X_k and X_{N-k}, inverse FFT, 1/N, undo the permutation
        exchange_rev<T, N, L>(v, vr, t, sm, exch);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int k = t + P * m;
            const Cx<T> xk = v[m], xm = vr[m];
            if (k == 0) {
                const T w0 = dwi[0].x;
                v[m] = Cx<T>{w0 * xk.x, w0 * xk.y};
            } else {
                const Cx<T> wk = dwi[k], wm = cconj(dwi[N - k]);
                const Cx<T> ha{(wk.x * xk.x + wm.x * xm.x) * T(0.5), (wk.y * xk.x + wm.y * xm.x) * T(0.5)};
                const Cx<T> hb{(wk.x * xk.y + wm.x * xm.y) * T(0.5), (wk.y * xk.y + wm.y * xm.y) * T(0.5)};
                v[m] = Cx<T>{ha.x - hb.y, ha.y + hb.x};
            }
        }
        fft_stages<T, N, true, 0, L, 3, MODE != 2>(v, t, sm, tw, exch);
        const T invn = T(1) / T(N);
#pragma unroll
        for (int m = 0; m < 8; m++) *out_at(perm(t + P * m)) = Ops<T>::scale(v[m], invn);
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) *out_at(t + P * m) = v[m];
    }
}

// ----------------------------------------------------------------------------------------
// x pass, real -> half spectrum: rows (2q, 2q+1) along y are transformed as one complex line
// ----------------------------------------------------------------------------------------
template <typename T, int N>
struct XShape {
    static constexpr int P = FftShape<N>::P;
    static constexpr int LP = P >= 256 ? 1 : 256 / P;  // row pairs per block
    static constexpr int NT = LP * P;
    static constexpr int SH = sizeof(T) == 4 ? 4 : 3;
    static constexpr int BUF = (spx<SH>(N - 1) + 1 + 3) / 4 * 4 + 2;  // line stride: staggers lines over the banks
    static constexpr size_t smem = (size_t)2 * LP * BUF * sizeof(Cx<T>);
};

template <typename T, int N, bool DIV>
__global__ void __launch_bounds__(XShape<T, N>::NT)
    fft_x_r2c_kernel(const T *__restrict__ R, Cx<T> *__restrict__ W, FastGeom g,
                     const Cx<T> *__restrict__ tw, long long npairs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef XShape<T, N> XS;
    constexpr int P = XS::P, LP = XS::LP;
    const int t = threadIdx.x % P, lp = threadIdx.x / P;
    long long q = (long long)blockIdx.x * LP + lp;
    const bool valid = q < npairs;
    if (!valid) q = npairs - 1;  // keep the thread for the barriers; it recomputes the last pair
    const int hp = g.n[1] / 2;
    const int jp = (int)(q % hp), kz = (int)(q / hp);
    const T *ra = R + (long long)(2 * jp) * g.rs[1] + (long long)kz * g.rs[2];
    const T *rb = ra + g.rs[1];
    LineSmem<T> sm;
    sm.buf[0] = reinterpret_cast<Cx<T> *>(smem_raw) + lp * XS::BUF;
    sm.buf[1] = sm.buf[0] + LP * XS::BUF;
    Cx<T> v[8];
    T fa = T(1), fb = T(1);
    if (g.pre) {
        const T *pre = (const T *)g.pre;
        fa = g.pre_dim == 1 ? pre[2 * jp] : pre[kz];
        fb = g.pre_dim == 1 ? pre[2 * jp + 1] : pre[kz];
    }
    if constexpr (DIV) {
        // rows a = 2 jp, b = 2 jp + 1 of the divergence, scaled with the cell volume
        // (divergence! + scalewithvolume!, src/operators.jl:218-247)
        const long long s1 = g.rs[1], s2 = g.rs[2];
        const T *U0 = (const T *)g.div_u + (long long)(2 * jp) * s1 + (long long)kz * s2;
        const T *U1 = U0 + g.div_cs, *U2 = U1 + g.div_cs;
        const long long ym = ((g.div_wrap & 2) && jp == 0) ? (long long)(g.n[1] - 1) * s1 : -s1;
        const long long zm = ((g.div_wrap & 4) && kz == 0) ? (long long)(g.n[2] - 1) * s2 : -s2;
        const T *r0 = (const T *)g.div_r[0], *d0 = (const T *)g.div_d[0];
        const T *r1 = (const T *)g.div_r[1], *d1 = (const T *)g.div_d[1];
        const T rya = r1[2 * jp], ryb = r1[2 * jp + 1], dya = d1[2 * jp], dyb = d1[2 * jp + 1];
        T rz = T(0), dz = T(1);
        if (g.D == 3) {
            rz = ((const T *)g.div_r[2])[kz];
            dz = ((const T *)g.div_d[2])[kz];
        }
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int i = t + P * m;
            const long long xm = ((g.div_wrap & 1) && i == 0) ? N - 1 : -1;
            const T ua = __ldg(U0 + i), uam = __ldg(U0 + i + xm);
            const T ub = __ldg(U0 + i + s1), ubm = __ldg(U0 + i + s1 + xm);
            const T v0 = __ldg(U1 + i + ym), v1 = __ldg(U1 + i), v2 = __ldg(U1 + i + s1);
            const T rx = __ldg(r0 + i), dx = __ldg(d0 + i);
            T da = (ua - uam) * rx + (v1 - v0) * rya;
            T db = (ub - ubm) * rx + (v2 - v1) * ryb;
            T oa = dx * dya, ob = dx * dyb;
            if (g.D == 3) {
                const T wa = __ldg(U2 + i), wam = __ldg(U2 + i + zm);
                const T wb = __ldg(U2 + i + s1), wbm = __ldg(U2 + i + s1 + zm);
                da = da + (wa - wam) * rz;
                db = db + (wb - wbm) * rz;
                oa = oa * dz;
                ob = ob * dz;
            }
            v[m] = Cx<T>{da * oa * fa, db * ob * fb};
        }
    } else {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = Cx<T>{ra[t + P * m] * fa, rb[t + P * m] * fb};
    }
    int exch = 0;
    constexpr int SH = XS::SH;
    fft_stages<T, N, false, 0, 1, SH>(v, t, sm, tw, exch);
    // exchange once more to reach Z[N-k]
    Cx<T> *buf = sm.buf[exch & 1];
#pragma unroll
    for (int m = 0; m < 8; m++) buf[spx<SH>(t + P * m)] = v[m];
    line_barrier<1, P>();
    Cx<T> *wa = W + (long long)(2 * jp) * g.ws[1] + (long long)kz * g.ws[2];
    Cx<T> *wb = wa + g.ws[1];
    if (valid) {
#pragma unroll
        for (int m = 0; m < 5; m++) {
            int k = t + P * m;
            if (m == 4 && t != 0) break;  // k = N/2 is handled by t = 0
            Cx<T> zk = m < 4 ? v[m] : buf[spx<SH>(k)];
            Cx<T> zc = cconj(buf[spx<SH>((N - k) & (N - 1))]);  // conj(Z[N-k])
            Cx<T> A = Ops<T>::scale(Ops<T>::add(zk, zc), T(0.5));
            Cx<T> d = Ops<T>::scale(Ops<T>::sub(zk, zc), T(0.5));
            if (g.x_kxl > 0) {  // blocks of modes go to their owners (peer stores)
                const int r = k / g.x_kxl;
                Cx<T> *o = reinterpret_cast<Cx<T> *>(g.outp[r]) +
                           ((long long)(kz + g.x_zoff) * g.n[1] + 2 * jp) * g.x_kxl + (k - r * g.x_kxl);
                o[0] = A;
                o[g.x_kxl] = Cx<T>{d.y, -d.x};
                continue;
            }
            wa[k] = A;
            wb[k] = Cx<T>{d.y, -d.x};  // d / i
        }
    }
}

// ----------------------------------------------------------------------------------------
// x pass with the divergence fused in, shared-memory staged variant.  The per-thread loads of
// the kernel above (11 scalar loads per element, 64 registers) leave it latency-bound at half
// the HBM rate; here the block first copies every velocity row it needs -- u_x rows y0..y0+2LP-1,
// u_y rows y0-1..y0+2LP-1, u_z the same rows of planes z and z-1, each with its left ghost --
// into shared memory with 8-byte cp.async copies (no registers held, everything in flight at
// once), evaluates divergence! + scalewithvolume! (src/operators.jl:218-247) from there and
// then reuses the same shared memory for the FFT exchanges.
// ----------------------------------------------------------------------------------------
template <typename T, int N>
struct XDivShape {
    static constexpr int P = FftShape<N>::P;
    static constexpr int NTMAX = sizeof(T) == 4 ? 256 : 128;
    static constexpr int LP = P >= NTMAX ? 1 : NTMAX / P;  // row pairs per block
    static constexpr int NT = LP * P;
    static constexpr int SH = XShape<T, N>::SH;
    static constexpr int BUF = XShape<T, N>::BUF;
    static constexpr int RW = N + 2;           // staged row: elements -1 .. N
    static constexpr int NR = 8 * LP + 1;      // staged rows
    static constexpr size_t stage_bytes = (size_t)NR * RW * sizeof(T);
    static constexpr size_t fft_bytes = (size_t)2 * LP * BUF * sizeof(Cx<T>);
    static constexpr size_t smem = stage_bytes > fft_bytes ? stage_bytes : fft_bytes;
};

template <typename T, int N>
__global__ void __launch_bounds__(XDivShape<T, N>::NT)
    fft_x_r2c_div_kernel(Cx<T> *__restrict__ W, FastGeom g, const Cx<T> *__restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef XDivShape<T, N> XS;
    constexpr int P = XS::P, LP = XS::LP, RW = XS::RW, NT = XS::NT;
    const int t = threadIdx.x % P, lp = threadIdx.x / P;
    const int hp = g.n[1] / 2;
    const long long q0 = (long long)blockIdx.x * LP;  // first pair of the block; LP divides hp
    const int jp0 = (int)(q0 % hp), kz = (int)(q0 / hp);
    const int y0 = 2 * jp0;
    const long long s1 = g.rs[1], s2 = g.rs[2];
    T *stage = reinterpret_cast<T *>(smem_raw);
    // ---- stage the rows: 8-byte copies starting at the left ghost of every row ----
    {
        const T *U0 = (const T *)g.div_u + (long long)kz * s2 - 1;  // element -1 of row y = 0, plane kz
        const T *U1 = U0 + g.div_cs, *U2 = U1 + g.div_cs;
        const long long ym = ((g.div_wrap & 2) && jp0 == 0) ? (long long)(g.n[1] - 1) * s1 : -s1;
        const long long zm = ((g.div_wrap & 4) && kz == 0) ? (long long)(g.n[2] - 1) * s2 : -s2;
        constexpr int EPC = 8 / (int)sizeof(T);        // elements per copy
        constexpr int CPR = (N + 1 + EPC - 1) / EPC;   // copies per row: elements -1 .. N-1 at least
        const int nrows = g.D == 3 ? 8 * LP + 1 : 4 * LP + 1;
        for (int c = threadIdx.x; c < nrows * CPR; c += NT) {
            const int r = c / CPR, e = (c - r * CPR) * EPC;
            const T *src;
            if (r < 2 * LP)
                src = U0 + (long long)(y0 + r) * s1;                                   // u_x rows
            else if (r < 4 * LP + 1)
                src = r == 2 * LP ? U1 + (long long)y0 * s1 + ym : U1 + (long long)(y0 + r - 2 * LP - 1) * s1;  // u_y rows y0-1 ..
            else if (r < 6 * LP + 1)
                src = U2 + (long long)(y0 + r - 4 * LP - 1) * s1;                      // u_z rows, plane z
            else
                src = U2 + (long long)(y0 + r - 6 * LP - 1) * s1 + zm;                 // u_z rows, plane z-1
            unsigned d = (unsigned)__cvta_generic_to_shared(stage + r * RW + e);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src + e) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // tables while the copies are in flight
    const int jp = jp0 + lp;
    const T *r0 = (const T *)g.div_r[0], *d0 = (const T *)g.div_d[0];
    const T *r1 = (const T *)g.div_r[1], *d1 = (const T *)g.div_d[1];
    const T rya = r1[2 * jp], ryb = r1[2 * jp + 1], dya = d1[2 * jp], dyb = d1[2 * jp + 1];
    T rz = T(0), dz = T(1);
    if (g.D == 3) {
        rz = ((const T *)g.div_r[2])[kz];
        dz = ((const T *)g.div_d[2])[kz];
    }
    T fa = T(1), fb = T(1);
    if (g.pre) {
        const T *pre = (const T *)g.pre;
        fa = g.pre_dim == 1 ? pre[2 * jp] : pre[kz];
        fb = g.pre_dim == 1 ? pre[2 * jp + 1] : pre[kz];
    }
    T rx[8], dx[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
        rx[m] = __ldg(r0 + t + P * m);
        dx[m] = __ldg(d0 + t + P * m);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    Cx<T> v[8];
    {
        const int a = 2 * lp;
        const T *UXa = stage + a * RW, *UXb = UXa + RW;
        const T *UY0 = stage + (2 * LP + a) * RW, *UY1 = UY0 + RW, *UY2 = UY1 + RW;
        const T *UZa = stage + (4 * LP + 1 + a) * RW, *UZb = UZa + RW;
        const T *UMa = stage + (6 * LP + 1 + a) * RW, *UMb = UMa + RW;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int i = t + P * m, ix = i + 1;
            const int im = ((g.div_wrap & 1) && i == 0) ? N : i;  // staged index of element i-1
            T da = (UXa[ix] - UXa[im]) * rx[m] + (UY1[ix] - UY0[ix]) * rya;
            T db = (UXb[ix] - UXb[im]) * rx[m] + (UY2[ix] - UY1[ix]) * ryb;
            T oa = dx[m] * dya, ob = dx[m] * dyb;
            if (g.D == 3) {
                da = da + (UZa[ix] - UMa[ix]) * rz;
                db = db + (UZb[ix] - UMb[ix]) * rz;
                oa = oa * dz;
                ob = ob * dz;
            }
            v[m] = Cx<T>{da * oa * fa, db * ob * fb};
        }
    }
    __syncthreads();  // the staging area becomes the exchange buffers
    LineSmem<T> sm;
    sm.buf[0] = reinterpret_cast<Cx<T> *>(smem_raw) + lp * XS::BUF;
    sm.buf[1] = sm.buf[0] + LP * XS::BUF;
    int exch = 0;
    constexpr int SH = XS::SH;
    fft_stages<T, N, false, 0, 1, SH>(v, t, sm, tw, exch);
    Cx<T> *buf = sm.buf[exch & 1];
#pragma unroll
    for (int m = 0; m < 8; m++) buf[spx<SH>(t + P * m)] = v[m];
    line_barrier<1, P>();
    Cx<T> *wa = W + (long long)(2 * jp) * g.ws[1] + (long long)kz * g.ws[2];
    Cx<T> *wb = wa + g.ws[1];
#pragma unroll
    for (int m = 0; m < 5; m++) {
        int k = t + P * m;
        if (m == 4 && t != 0) break;  // k = N/2 is handled by t = 0
        Cx<T> zk = m < 4 ? v[m] : buf[spx<SH>(k)];
        Cx<T> zc = cconj(buf[spx<SH>((N - k) & (N - 1))]);  // conj(Z[N-k])
        Cx<T> A = Ops<T>::scale(Ops<T>::add(zk, zc), T(0.5));
        Cx<T> d = Ops<T>::scale(Ops<T>::sub(zk, zc), T(0.5));
        if (g.x_kxl > 0) {  // blocks of modes go to their owners (peer stores)
            const int r = k / g.x_kxl;
            Cx<T> *o = reinterpret_cast<Cx<T> *>(g.outp[r]) +
                       ((long long)(kz + g.x_zoff) * g.n[1] + 2 * jp) * g.x_kxl + (k - r * g.x_kxl);
            o[0] = A;
            o[g.x_kxl] = Cx<T>{d.y, -d.x};
            continue;
        }
        wa[k] = A;
        wb[k] = Cx<T>{d.y, -d.x};  // d / i
    }
}

// ----------------------------------------------------------------------------------------
// x pass, half spectrum -> real: inverse of the above
// ----------------------------------------------------------------------------------------
template <typename T, int N>
__global__ void __launch_bounds__(XShape<T, N>::NT)
    fft_x_c2r_kernel(T *__restrict__ R, const Cx<T> *__restrict__ W, FastGeom g,
                     const Cx<T> *__restrict__ tw, long long npairs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef XShape<T, N> XS;
    constexpr int P = XS::P, LP = XS::LP;
    const int t = threadIdx.x % P, lp = threadIdx.x / P;
    long long q = (long long)blockIdx.x * LP + lp;
    const bool valid = q < npairs;
    if (!valid) q = npairs - 1;
    const int hp = g.n[1] / 2;
    const int jp = (int)(q % hp), kz = (int)(q / hp);
    const Cx<T> *wa = W + (long long)(2 * jp) * g.ws[1] + (long long)kz * g.ws[2];
    const Cx<T> *wb = wa + g.ws[1];
    LineSmem<T> sm;
    sm.buf[0] = reinterpret_cast<Cx<T> *>(smem_raw) + lp * XS::BUF;
    sm.buf[1] = sm.buf[0] + LP * XS::BUF;
    Cx<T> v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
        int p = t + P * m;
        if (m < 4) {  // Z[k] = A[k] + i B[k]
            Cx<T> a = wa[p], b = wb[p];
            v[m] = Cx<T>{a.x - b.y, a.y + b.x};
        } else {      // Z[p] = conj(A[N-p]) + i conj(B[N-p])
            Cx<T> a = wa[N - p], b = wb[N - p];
            v[m] = Cx<T>{a.x + b.y, b.x - a.y};
        }
    }
    int exch = 0;
    fft_stages<T, N, true, 0, 1, XS::SH>(v, t, sm, tw, exch);
    if (valid) {
        T *ra = R + (long long)(2 * jp) * g.rs[1] + (long long)kz * g.rs[2];
        T *rb = ra + g.rs[1];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            ra[t + P * m] = v[m].x;
            rb[t + P * m] = v[m].y;
        }
        if (g.img) {
            // two ghost layers of periodic images: rows 0, 1 -> n1, n1 + 1; rows n1 - 2, n1 - 1 -> -2, -1; the
            // same along x and (bit 1) z
            const int n1 = g.n[1], n2 = g.n[2];
            const long long yo = (jp == 0 ? (long long)n1 : (2 * jp + 2 == n1 ? -(long long)n1 : 0)) * g.rs[1];
            const long long zo = !(g.img & 2) ? 0 : (kz < 2 ? (long long)n2 : (kz >= n2 - 2 ? -(long long)n2 : 0)) * g.rs[2];
            for (int iz = 0; iz < (zo ? 2 : 1); iz++)
                for (int iy = 0; iy < (yo ? 2 : 1); iy++) {
                    T *a2 = ra + (iz ? zo : 0) + (iy ? yo : 0), *b2 = a2 + g.rs[1];
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const int i = t + P * m;
                        const int xo = i < 2 ? N : (i >= N - 2 ? -N : 0);
                        if (iz | iy) a2[i] = v[m].x, b2[i] = v[m].y;
                        if (xo) a2[i + xo] = v[m].x, b2[i + xo] = v[m].y;
                    }
                }
        }
    }
}

// ----------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------
template <typename K>
static int allow_smem(K kern, size_t smem) {
    if (smem > 48 * 1024) INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

// tensor maps of the half-spectrum workspace (cuTensorMapEncodeTiled through the runtime's driver entry point)
typedef CUresult (*FftEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct FftMapEntry {
    const void *ptr;
    long long key[6];
    CUtensorMap map;
};
static std::vector<FftMapEntry> g_fft_maps;  // (a tensor map depends on the address and the geometry only)
static std::mutex g_fft_maps_mu;

// W viewed as [n2][n1][m0p complex] of 4-byte words; box = rowb bytes (L lines) x BOXR rows along `tdim`
static int fft_tensor_map(const void *W, long long m0p, int n1, int n2, long long ws1, long long ws2, int esz_cx, int tdim,
                          int boxr, int rowb, CUtensorMap *out) {
    const long long key[6] = {m0p, n1, n2, ws1 * 1000003 + ws2, esz_cx * 16 + tdim, boxr * 1024 + rowb};
    std::lock_guard<std::mutex> lock(g_fft_maps_mu);
    for (auto &e : g_fft_maps)
        if (e.ptr == W && !std::memcmp(e.key, key, sizeof(key))) {
            *out = e.map;
            return 0;
        }
    static FftEncodeFn enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = reinterpret_cast<FftEncodeFn>(p);
    }
    if (!enc) return 1;
    const cuuint64_t wpe = (cuuint64_t)esz_cx / 4;  // 4-byte words per complex element
    cuuint64_t dims[3] = {(cuuint64_t)m0p * wpe, (cuuint64_t)n1, (cuuint64_t)n2};
    cuuint64_t strides[2] = {(cuuint64_t)ws1 * esz_cx, (cuuint64_t)ws2 * esz_cx};
    cuuint32_t box[3] = {(cuuint32_t)rowb / 4, tdim == 1 ? (cuuint32_t)boxr : 1u, tdim == 1 ? 1u : (cuuint32_t)boxr};
    cuuint32_t estr[3] = {1, 1, 1};
    FftMapEntry e;
    e.ptr = W;
    std::memcpy(e.key, key, sizeof(key));
    if (enc(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(W), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    if (g_fft_maps.size() > 64) g_fft_maps.clear();
    g_fft_maps.push_back(e);
    *out = e.map;
    return 0;
}

template <typename T, int N>
static int fast_pass_n(int kind, T *R, Cx<T> *W, const FastGeom &g, int tdim, const Cx<T> *tw, const T *e0,
                       const T *e1, const T *e2, T factor, Launch L) {
    if (kind <= 2) {  // strided
        typedef StridedShape<T, N> SS;
        if (g.m0p % SS::L) {
            set_error("fast FFT pass: padded half-spectrum extent must be a multiple of the line tile");
            return 1;
        }
        const bool blk = g.in_sh < 30 || g.out_sh < 30 || g.use_outp || g.in_grp || g.out_grp;
        // TMA-pipelined persistent variant: in place on the natural layout of W, 3-D, lines of >= 256 elements (measured
        // at 512^3 Float32: forward / inverse 0.228 / 0.214 -> 0.203 / 0.200 ms, middle pass 0.363 -> 0.338 ms; lines of
        // 1024: 0.364 -> 0.329 ms; lines of 128 lose, 0.36 -> 0.48 ms: the per-tile hand-shake outweighs an 8 KB tile).
        // INSB200_FFT_TMA=0 disables it, =1 restricts it to the middle pass.
        {
            const char *tma_env = getenv("INSB200_FFT_TMA");  // (read per call, like the other switches)
            const int mode = tma_env ? atoi(tma_env) : 2;
            typedef TmaStridedShape<T, N> TS;
            const bool plain = !blk && !g.in && !g.out && g.in_sd < 0 && g.out_sd < 0 && g.in_os < 0 && g.out_os < 0 && g.nother < 0;
            if (sizeof(T) == 4 && N >= 256 && g.D == 3 && plain && mode > 0 && (kind == 2 || mode >= 2) &&
                TS::smem <= 227 * 1024) {
                CUtensorMap tm;
                if (fft_tensor_map(W, g.m0p, g.n[1], g.n[2], g.ws[1], g.ws[2], (int)sizeof(Cx<T>), tdim, TS::BOXR, TS::ROWB, &tm) == 0) {
                    const int ntx = g.m0p / SS::L, ntiles = ntx * g.n[tdim == 1 ? 2 : 1];
                    auto kern = kind == 0 ? fft_strided_tma_kernel<T, N, 0> : kind == 1 ? fft_strided_tma_kernel<T, N, 1>
                                                                                        : fft_strided_tma_kernel<T, N, 2>;
                    INSB_TRY(allow_smem(kern, TS::smem));
                    INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    const int per_sm = SS::NT <= 512 ? 2 : 1;
                    int nsm = 148;
                    const int grid = ntiles < nsm * per_sm ? ntiles : nsm * per_sm;
                    kern<<<grid, SS::NT, TS::smem, L.stream>>>(tm, g, tdim, ntx, ntiles, tw, e0, e1, e2, factor);
                    if (L.counter) ++*L.counter;
                    INSB_CUDA(cudaGetLastError());
                    return 0;
                }
            }
        }
        dim3 grid(g.m0p / SS::L, g.D == 3 ? (g.nother >= 0 ? g.nother : g.n[tdim == 1 ? 2 : 1]) : 1);
        auto kern = kind == 0 ? (blk ? fft_strided_kernel<T, N, 0, true> : fft_strided_kernel<T, N, 0, false>)
                    : kind == 1 ? (blk ? fft_strided_kernel<T, N, 1, true> : fft_strided_kernel<T, N, 1, false>)
                                : (blk ? fft_strided_kernel<T, N, 2, true> : fft_strided_kernel<T, N, 2, false>);
        INSB_TRY(allow_smem(kern, SS::smem));
        kern<<<grid, SS::NT, SS::smem, L.stream>>>(W, g, tdim, tw, e0, e1, e2, factor);
    } else if (kind >= 5) {  // strided DCT
        typedef StridedShape<T, N> SS;
        if (g.m0p % SS::L || !g.dw || !g.dwi) {
            set_error("fast DCT pass: bad geometry");
            return 1;
        }
        const bool blk = g.in_sh < 30 || g.out_sh < 30 || g.use_outp || g.in_grp || g.out_grp || g.in || g.out;
        dim3 grid(g.m0p / SS::L, g.D == 3 ? (g.nother >= 0 ? g.nother : g.n[tdim == 1 ? 2 : 1]) : 1);
        auto kern = kind == 5   ? (blk ? fft_strided_dct_kernel<T, N, 0, true> : fft_strided_dct_kernel<T, N, 0, false>)
                    : kind == 6 ? (blk ? fft_strided_dct_kernel<T, N, 1, true> : fft_strided_dct_kernel<T, N, 1, false>)
                                : (blk ? fft_strided_dct_kernel<T, N, 2, true> : fft_strided_dct_kernel<T, N, 2, false>);
        INSB_TRY(allow_smem(kern, SS::smem));
        kern<<<grid, SS::NT, SS::smem, L.stream>>>(W, g, tdim, tw, e0, e1, e2, factor);
    } else {
        typedef XShape<T, N> XS;
        long long npairs = (long long)(g.n[1] / 2) * (g.D == 3 ? g.n[2] : 1);
        unsigned blocks = (unsigned)((npairs + XS::LP - 1) / XS::LP);
        typedef XDivShape<T, N> XD;
        const char *sd = getenv("INSB200_FUSED_DIVERGENCE");
        // staged variant (INSB200_FUSED_DIVERGENCE=2): whole blocks inside one z plane, even row length
        // (8-byte aligned row starts), 8-byte aligned base and component stride
        const bool staged = kind == 3 && g.div_u && sd != nullptr && atoi(sd) == 2 && XD::smem <= 200 * 1024 &&
                            (g.n[1] / 2) % XD::LP == 0 && g.rs[1] % 2 == 0 && (g.rs[2] % 2 == 0 || g.D == 2) &&
                            g.div_cs % 2 == 0 && ((uintptr_t)((const T *)g.div_u - 1)) % 8 == 0;
        if (staged) {
            INSB_TRY(allow_smem(fft_x_r2c_div_kernel<T, N>, XD::smem));
            INSB_CUDA(cudaFuncSetAttribute(fft_x_r2c_div_kernel<T, N>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
            fft_x_r2c_div_kernel<T, N><<<(unsigned)(npairs / XD::LP), XD::NT, XD::smem, L.stream>>>(W, g, tw);
        } else if (kind == 3 && g.div_u) {
            INSB_TRY(allow_smem(fft_x_r2c_kernel<T, N, true>, XS::smem));
            fft_x_r2c_kernel<T, N, true><<<blocks, XS::NT, XS::smem, L.stream>>>(R, W, g, tw, npairs);
        } else if (kind == 3) {
            INSB_TRY(allow_smem(fft_x_r2c_kernel<T, N, false>, XS::smem));
            fft_x_r2c_kernel<T, N, false><<<blocks, XS::NT, XS::smem, L.stream>>>(R, W, g, tw, npairs);
        } else {
            INSB_TRY(allow_smem(fft_x_c2r_kernel<T, N>, XS::smem));
            fft_x_c2r_kernel<T, N><<<blocks, XS::NT, XS::smem, L.stream>>>(R, W, g, tw, npairs);
        }
    }
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

// kind: 0 strided fwd, 1 strided inv, 2 strided mid, 3 x r2c, 4 x c2r, 5/6/7 strided DCT fwd/inv/mid
template <typename T>
int fast_pass(int kind, int N, T *R, void *W, const FastGeom &g, int tdim, const void *tw, const T *e0,
              const T *e1, const T *e2, T factor, int tile, Launch L) {
    Cx<T> *Wc = (Cx<T> *)W;
    const Cx<T> *twc = (const Cx<T> *)tw;
    (void)tile;
    switch (N) {
        case 32: return fast_pass_n<T, 32>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
        case 64: return fast_pass_n<T, 64>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
        case 128: return fast_pass_n<T, 128>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
        case 256: return fast_pass_n<T, 256>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
        case 512: return fast_pass_n<T, 512>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
        case 1024: return fast_pass_n<T, 1024>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
        case 2048: return fast_pass_n<T, 2048>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, L);
    }
    set_error("fast FFT pass: unsupported length");
    return 1;
}

bool fast_fft_supported(int n) { return n >= 32 && n <= 2048 && (n & (n - 1)) == 0; }

template int fast_pass<float>(int, int, float *, void *, const FastGeom &, int, const void *, const float *,
                              const float *, const float *, float, int, Launch);
template int fast_pass<double>(int, int, double *, void *, const FastGeom &, int, const void *, const double *,
                               const double *, const double *, double, int, Launch);

}  // namespace insb
