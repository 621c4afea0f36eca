// fft_fast.cu -- register/shared-memory FFT passes for power-of-two lengths (32..2048).
//
// One thread owns 8 complex values of a line of length N (P = N/8 threads per line) and runs
// a Stockham auto-sort FFT whose radix-8/4/2 butterflies live entirely in registers; values
// are exchanged between stages through padded, double-buffered shared memory.  In EVERY
// stage register m of thread t holds element t + P*m, and so does the final result -- which
// is also the input distribution of the next transform.  The middle pass of the Poisson
// solve exploits that: forward FFT, eigenvalue division (pressure.jl:360-372) and inverse
// FFT run back to back without leaving registers in between.
//
// Pass structure of psolver_spectral (pressure.jl:349-381) with these kernels, 3-D:
//   x: two real rows -> one complex FFT -> two half spectra (R2C), reads view(p, Ip) directly
//   y: strided forward FFT        z: strided forward FFT + divide + inverse FFT
//   y: strided inverse FFT        x: two half spectra -> two real rows (C2R), writes view(p, Ip)
// i.e. 5 passes of one read + one write each (10 w per cell per solve).
#include <cstdlib>

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

__host__ __device__ __forceinline__ constexpr int phys(int p) { return p + (p >> 3); }

template <typename T, bool INV>
__device__ __forceinline__ Cx<T> rot90(Cx<T> a) {  // multiply by -i (forward) or +i (inverse)
    return INV ? Cx<T>{-a.y, a.x} : Cx<T>{a.y, -a.x};
}

// radix-R butterfly on v[b + B*r], r = 0..R-1, natural-order output
template <typename T, int R, int B, bool INV>
__device__ __forceinline__ void butterfly(Cx<T> (&v)[8], int b) {
    if constexpr (R == 2) {
        Cx<T> a = v[b], c = v[b + B];
        v[b] = cadd(a, c);
        v[b + B] = csub(a, c);
    } else if constexpr (R == 4) {
        Cx<T> a0 = cadd(v[b], v[b + 2 * B]), a1 = csub(v[b], v[b + 2 * B]);
        Cx<T> a2 = cadd(v[b + B], v[b + 3 * B]), a3 = rot90<T, INV>(csub(v[b + B], v[b + 3 * B]));
        v[b] = cadd(a0, a2);
        v[b + B] = cadd(a1, a3);
        v[b + 2 * B] = csub(a0, a2);
        v[b + 3 * B] = csub(a1, a3);
    } else {
        const T h = T(0.70710678118654752440);
        Cx<T> a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
        Cx<T> a2 = cadd(v[2], v[6]), a3 = rot90<T, INV>(csub(v[2], v[6]));
        Cx<T> a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
        Cx<T> a6 = cadd(v[3], v[7]), a7 = rot90<T, INV>(csub(v[3], v[7]));
        Cx<T> b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = csub(a1, a3);
        Cx<T> b4 = cadd(a4, a6), b6 = csub(a4, a6), b5 = cadd(a5, a7), b7 = csub(a5, a7);
        // w8^1 = (1 -+ i)/sqrt2, w8^2 = -+ i, w8^3 = (-1 -+ i)/sqrt2
        Cx<T> t5 = INV ? Cx<T>{(b5.x - b5.y) * h, (b5.x + b5.y) * h} : Cx<T>{(b5.x + b5.y) * h, (b5.y - b5.x) * h};
        Cx<T> t6 = rot90<T, INV>(b6);
        Cx<T> t7 = INV ? Cx<T>{(-b7.x - b7.y) * h, (b7.x - b7.y) * h} : Cx<T>{(b7.y - b7.x) * h, (-b7.x - b7.y) * h};
        v[0] = cadd(b0, b4);
        v[4] = csub(b0, b4);
        v[1] = cadd(b1, t5);
        v[5] = csub(b1, t5);
        v[2] = cadd(b2, t6);
        v[6] = csub(b2, t6);
        v[3] = cadd(b3, t7);
        v[7] = csub(b3, t7);
    }
}

// Shared-memory view of one line: two buffers x (re, im), padded index phys(p).
template <typename T>
struct LineSmem {
    T *re[2], *im[2];
};

// Shared-memory index of element p for the exchange that follows a stage with (NS, R).
// LFAST kernels (lanes run over lines): p + p/8.  TFAST kernels (lanes run along the line):
// pad PADW words per BANKS elements, PADW chosen so that both the strided writes of this
// exchange and the unit-stride reads are bank-conflict free.
template <typename T, bool TFAST, int NS, int R>
__device__ __forceinline__ int sidx(int p) {
    if constexpr (!TFAST) {
        return p + (p >> 3);
    } else {
        constexpr int BANKS = 128 / (int)sizeof(T);
        constexpr int PADW = (NS * R >= BANKS) ? (BANKS / R > 0 ? BANKS / R : 1) : NS;
        return p + PADW * (p / BANKS);
    }
}

template <typename T, int N, bool INV, int S, bool TFAST = false>
__device__ __forceinline__ void fft_stages(Cx<T> (&v)[8], int t, const LineSmem<T> &sm,
                                           const Cx<T> *__restrict__ tw, int &exch) {
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
                v[b + B] = cmul(v[b + B], w1);
            } else if constexpr (R == 4) {
                Cx<T> w1 = tw[k * TSTEP], w2 = tw[2 * k * TSTEP];
                if (INV) w1 = cconj(w1), w2 = cconj(w2);
                Cx<T> w3 = cmul(w1, w2);
                v[b + B] = cmul(v[b + B], w1);
                v[b + 2 * B] = cmul(v[b + 2 * B], w2);
                v[b + 3 * B] = cmul(v[b + 3 * B], w3);
            } else {
                Cx<T> w1 = tw[k * TSTEP], w2 = tw[2 * k * TSTEP], w4 = tw[4 * k * TSTEP];
                if (INV) w1 = cconj(w1), w2 = cconj(w2), w4 = cconj(w4);
                Cx<T> w3 = cmul(w1, w2), w5 = cmul(w1, w4), w6 = cmul(w2, w4);
                Cx<T> w7 = cmul(w3, w4);
                v[1] = cmul(v[1], w1);
                v[2] = cmul(v[2], w2);
                v[3] = cmul(v[3], w3);
                v[4] = cmul(v[4], w4);
                v[5] = cmul(v[5], w5);
                v[6] = cmul(v[6], w6);
                v[7] = cmul(v[7], w7);
            }
        }
    }
#pragma unroll
    for (int b = 0; b < B; b++) butterfly<T, R, B, INV>(v, b);
    if constexpr (S < Sh::NST - 1) {
        T *re = sm.re[exch & 1], *im = sm.im[exch & 1];
#pragma unroll
        for (int b = 0; b < B; b++) {
            int j = t + P * b;
            int j0 = (j / NS) * (NS * R) + (j & (NS - 1));
#pragma unroll
            for (int r = 0; r < R; r++) {
                int q = sidx<T, TFAST, NS, R>(j0 + r * NS);
                re[q] = v[b + B * r].x;
                im[q] = v[b + B * r].y;
            }
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int q = sidx<T, TFAST, NS, R>(t + P * m);
            v[m] = Cx<T>{re[q], im[q]};
        }
        exch++;
        fft_stages<T, N, INV, S + 1, TFAST>(v, t, sm, tw, exch);
    }
}

template <typename T>
__device__ __forceinline__ LineSmem<T> line_smem(T *base, int l, int L, int LS) {
    LineSmem<T> s;
    s.re[0] = base + (size_t)l * LS;
    s.im[0] = base + (size_t)(L + l) * LS;
    s.re[1] = base + (size_t)(2 * L + l) * LS;
    s.im[1] = base + (size_t)(3 * L + l) * LS;
    return s;
}


// ----------------------------------------------------------------------------------------
// strided pass along dim `tdim` (1 or 2): MODE 0 forward, 1 inverse, 2 forward+divide+inverse
// grid: x = tiles of L lines along kx, y = index along the remaining dimension
// ----------------------------------------------------------------------------------------
// MODE 3 = MODE 2 compiled without the two-blocks-per-SM register cap (selected at run time
// with INSB200_FFT_MID_UNCAPPED=1 for A/B measurements)
template <typename T, int N, int MODE>
__global__ void __launch_bounds__(512, (sizeof(T) == 4 && MODE != 3) ? 2 : 1)
    fft_strided_kernel(Cx<T> *__restrict__ W, FastGeom g, int tdim, int L, int LS,
                       const Cx<T> *__restrict__ tw, const T *__restrict__ e0, const T *__restrict__ e1,
                       const T *__restrict__ e2, T factor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = FftShape<N>::P;
    const int l = threadIdx.x % L, t = threadIdx.x / L;
    const int kx = blockIdx.x * L + l;
    const int other = blockIdx.y;
    const int odim = tdim == 1 ? 2 : 1;
    const long long in_sd = g.in_sd >= 0 ? g.in_sd : g.ws[tdim], out_sd = g.out_sd >= 0 ? g.out_sd : g.ws[tdim];
    const long long in_os = g.in_os >= 0 ? g.in_os : g.ws[odim], out_os = g.out_os >= 0 ? g.out_os : g.ws[odim];
    const Cx<T> *ibase = (g.in ? (const Cx<T> *)g.in : W) + kx + (g.D == 3 ? other * in_os : 0);
    Cx<T> *obase = (g.out ? (Cx<T> *)g.out : W) + kx + (g.D == 3 ? other * out_os : 0);
    const int in_mask = (1 << g.in_sh) - 1, out_mask = (1 << g.out_sh) - 1;
    auto ioff = [&](int k) { return (long long)(k >> g.in_sh) * g.in_blk + (long long)(k & in_mask) * in_sd; };
    auto ooff = [&](int k) { return (long long)(k >> g.out_sh) * g.out_blk + (long long)(k & out_mask) * out_sd; };
    LineSmem<T> sm = line_smem<T>(reinterpret_cast<T *>(smem_raw), l, L, LS);
    Cx<T> v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = ibase[ioff(t + P * m)];
    T f[8];
    if (MODE >= 2) {
        // phat = -phat / ((ax + ay) + az), phat[1] = 0 (pressure.jl:360-372); `factor` carries
        // the sign and the 1/prod(Np) of the inverse transform.  Evaluated while the loads
        // above are in flight.
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
    if (MODE == 0 || MODE >= 2) fft_stages<T, N, false, 0>(v, t, sm, tw, exch);
    if (MODE >= 2) {
#pragma unroll
        for (int m = 0; m < 8; m++) v[m] = Cx<T>{v[m].x * f[m], v[m].y * f[m]};
    }
    if (MODE >= 1) fft_stages<T, N, true, 0>(v, t, sm, tw, exch);
#pragma unroll
    for (int m = 0; m < 8; m++) obase[ooff(t + P * m)] = v[m];
}

// ----------------------------------------------------------------------------------------
// x pass, real -> half spectrum: rows (2q, 2q+1) along y are transformed as one complex line
// ----------------------------------------------------------------------------------------
template <typename T, int N>
__global__ void __launch_bounds__(256)
    fft_x_r2c_kernel(const T *__restrict__ R, Cx<T> *__restrict__ W, FastGeom g, int LP, int LS,
                     const Cx<T> *__restrict__ tw, long long npairs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = FftShape<N>::P;
    constexpr int NST = FftShape<N>::NST;
    const int t = threadIdx.x % P, lp = threadIdx.x / P;
    long long q = (long long)blockIdx.x * LP + lp;
    const bool valid = q < npairs;
    if (!valid) q = npairs - 1;  // keep the thread for the barriers; it recomputes the last pair
    const int hp = g.n[1] / 2;
    const int jp = (int)(q % hp), kz = (int)(q / hp);
    const T *ra = R + (long long)(2 * jp) * g.rs[1] + (long long)kz * g.rs[2];
    const T *rb = ra + g.rs[1];
    LineSmem<T> sm = line_smem<T>(reinterpret_cast<T *>(smem_raw), lp, LP, LS);
    Cx<T> v[8];
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = Cx<T>{ra[t + P * m], rb[t + P * m]};
    int exch = 0;
    fft_stages<T, N, false, 0, true>(v, t, sm, tw, exch);
    // exchange once more to reach Z[N-k]
    T *re = sm.re[exch & 1], *im = sm.im[exch & 1];
    (void)NST;
    auto phys = [](int p) { return sidx<T, true, 1, 1>(p); };
#pragma unroll
    for (int m = 0; m < 8; m++) {
        int p = phys(t + P * m);
        re[p] = v[m].x;
        im[p] = v[m].y;
    }
    __syncthreads();
    Cx<T> *wa = W + (long long)(2 * jp) * g.ws[1] + (long long)kz * g.ws[2];
    Cx<T> *wb = wa + g.ws[1];
    if (valid) {
#pragma unroll
        for (int m = 0; m < 5; m++) {
            int k = t + P * m;
            if (m == 4 && t != 0) break;  // k = N/2 is handled by t = 0
            Cx<T> zk = m < 4 ? v[m] : Cx<T>{re[phys(k)], im[phys(k)]};
            int kc = phys((N - k) & (N - 1));
            Cx<T> zc{re[kc], -im[kc]};  // conj(Z[N-k])
            Cx<T> A{(zk.x + zc.x) * T(0.5), (zk.y + zc.y) * T(0.5)};
            Cx<T> d{(zk.x - zc.x) * T(0.5), (zk.y - zc.y) * T(0.5)};
            wa[k] = A;
            wb[k] = Cx<T>{d.y, -d.x};  // d / i
        }
    }
}

// ----------------------------------------------------------------------------------------
// x pass, half spectrum -> real: inverse of the above
// ----------------------------------------------------------------------------------------
template <typename T, int N>
__global__ void __launch_bounds__(256)
    fft_x_c2r_kernel(T *__restrict__ R, const Cx<T> *__restrict__ W, FastGeom g, int LP, int LS,
                     const Cx<T> *__restrict__ tw, long long npairs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int P = FftShape<N>::P;
    const int t = threadIdx.x % P, lp = threadIdx.x / P;
    long long q = (long long)blockIdx.x * LP + lp;
    const bool valid = q < npairs;
    if (!valid) q = npairs - 1;
    const int hp = g.n[1] / 2;
    const int jp = (int)(q % hp), kz = (int)(q / hp);
    const Cx<T> *wa = W + (long long)(2 * jp) * g.ws[1] + (long long)kz * g.ws[2];
    const Cx<T> *wb = wa + g.ws[1];
    LineSmem<T> sm = line_smem<T>(reinterpret_cast<T *>(smem_raw), lp, LP, LS);
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
    fft_stages<T, N, true, 0, true>(v, t, sm, tw, exch);
    if (valid) {
        T *ra = R + (long long)(2 * jp) * g.rs[1] + (long long)kz * g.rs[2];
        T *rb = ra + g.rs[1];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            ra[t + P * m] = v[m].x;
            rb[t + P * m] = v[m].y;
        }
    }
}

// ----------------------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------------------
static int line_stride(int N, int lines_per_txn, int elem_bytes) {
    int banks = 128 / elem_bytes;
    int want = banks / (lines_per_txn < banks ? (lines_per_txn < 1 ? 1 : lines_per_txn) : banks);
    int ls = phys(N - 1) + 1;
    while (ls % banks != want % banks) ls++;
    return ls;
}

template <typename T, int N>
static int fast_pass_n(int kind, T *R, Cx<T> *W, const FastGeom &g, int tdim, const Cx<T> *tw, const T *e0,
                       const T *e1, const T *e2, T factor, int tile, Launch L) {
    constexpr int P = FftShape<N>::P;
    static bool attr_done[5] = {false, false, false, false, false};
    if (kind <= 2) {  // strided
        int Lt = tile;
        while (Lt * P > 512 && Lt > 1) Lt >>= 1;
        int LS = line_stride(N, Lt, sizeof(T));
        size_t smem = (size_t)4 * Lt * LS * sizeof(T);
        dim3 grid(g.m0p / Lt, g.D == 3 ? (g.nother >= 0 ? g.nother : g.n[tdim == 1 ? 2 : 1]) : 1);
        static const bool uncapped = getenv("INSB200_FFT_MID_UNCAPPED") != nullptr;
        auto kern = kind == 0 ? fft_strided_kernel<T, N, 0> : kind == 1 ? fft_strided_kernel<T, N, 1>
                    : uncapped ? fft_strided_kernel<T, N, 3> : fft_strided_kernel<T, N, 2>;
        INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (smem > 200 * 1024) {
            set_error("fast FFT pass: tile does not fit in shared memory");
            return 1;
        }
        kern<<<grid, Lt * P, smem, L.stream>>>(W, g, tdim, Lt, LS, tw, e0, e1, e2, factor);
    } else {
        int LP = 256 / P;
        if (LP < 1) LP = 1;
        int LS = line_stride(N, 32 / (P < 32 ? P : 32), sizeof(T));
        size_t smem = (size_t)4 * LP * LS * sizeof(T);
        long long npairs = (long long)(g.n[1] / 2) * (g.D == 3 ? g.n[2] : 1);
        unsigned blocks = (unsigned)((npairs + LP - 1) / LP);
        if (kind == 3) {
            if (!attr_done[3]) {
                INSB_CUDA(cudaFuncSetAttribute(fft_x_r2c_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
                attr_done[3] = true;
            }
            fft_x_r2c_kernel<T, N><<<blocks, LP * P, smem, L.stream>>>(R, W, g, LP, LS, tw, npairs);
        } else {
            if (!attr_done[4]) {
                INSB_CUDA(cudaFuncSetAttribute(fft_x_c2r_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
                attr_done[4] = true;
            }
            fft_x_c2r_kernel<T, N><<<blocks, LP * P, smem, L.stream>>>(R, W, g, LP, LS, tw, npairs);
        }
    }
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

// kind: 0 strided fwd, 1 strided inv, 2 strided mid, 3 x r2c, 4 x c2r
template <typename T>
int fast_pass(int kind, int N, T *R, void *W, const FastGeom &g, int tdim, const void *tw, const T *e0,
              const T *e1, const T *e2, T factor, int tile, Launch L) {
    Cx<T> *Wc = (Cx<T> *)W;
    const Cx<T> *twc = (const Cx<T> *)tw;
    switch (N) {
        case 32: return fast_pass_n<T, 32>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
        case 64: return fast_pass_n<T, 64>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
        case 128: return fast_pass_n<T, 128>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
        case 256: return fast_pass_n<T, 256>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
        case 512: return fast_pass_n<T, 512>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
        case 1024: return fast_pass_n<T, 1024>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
        case 2048: return fast_pass_n<T, 2048>(kind, R, Wc, g, tdim, twc, e0, e1, e2, factor, tile, L);
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
