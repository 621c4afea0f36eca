// poisson.cu -- FFT-type Poisson solvers on the pressure DOFs, hand-written (no cuFFT).
//
// Replaces the closures returned by psolver_spectral (src/pressure.jl:321-382),
// psolver_transform (:215-318, with manual_dct!/manual_idct! of src/utils.jl:65-155) and
// psolver_tridiagonal (:401-545).  The solve is a sequence of "line passes": every pass
// moves a tile of lines (along one dimension) into shared memory, transforms it there
// (mixed-radix Stockham FFT, DCT-II/III through the FFT, or the Thomas recurrence) and
// writes it back, so each pass is one read + one write of the spectrum in HBM.  The
// eigenvalue division (pressure.jl:360-372 / :281-293) is fused into the middle pass
// together with the inverse transform of that dimension; pack/unpack of view(p, Ip)
// (pressure.jl:351,378) is fused into the first/last pass.
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "internal.h"
#include "cx.cuh"
#include "fft_fast.h"

namespace insb {

enum PassOp {
    P_FFT_FWD = 0,
    P_FFT_INV = 1,
    P_DCT_FWD = 2,
    P_DCT_INV = 3,
    P_MID_FFT = 4,
    P_MID_DCT = 5,
    P_THOMAS = 6
};

constexpr int MAXRAD = 16;

template <typename T>
struct PassArgs {
    int op;
    int dim;        // dimension the lines run along
    int n;          // transform length
    int D;
    int src_real;   // load from the ghosted real array (imag = 0)
    int dst_real;   // store real part to the ghosted real array
    int half_out;   // store only the first n/2+1 entries (R2C)
    int half_in;    // load n/2+1 entries and Hermitian-extend (C2R)
    Cx<T> *W;       // complex workspace
    long long ws[3];
    int m[3];       // complex extents of W
    T *R;           // ghosted real array, already offset to the first pressure DOF
    long long rs[3];
    int tile;       // lines per block
    int ld;         // smem leading dimension (complex elements)
    int nrad;
    int radix[MAXRAD];
    const Cx<T> *tw;    // exp(-2 pi i q / n), q = 0..n-1
    const Cx<T> *dw;    // DCT twiddles w_k    (utils.jl:68-75)
    const Cx<T> *dwi;   // DCT twiddles winv_k (utils.jl:76-83)
    const T *eig[3];    // eigenvalue tables (ahat / lam)
    T sign;             // -1 for spectral (phat = -phat/(...)), +1 for transform
    T scale;            // 1 / prod(n_periodic)
    const T *prescale;  // optional factor on load, indexed along pdim (invΩ, pressure.jl:463-465)
    int pdim;
    // Thomas
    const T *ta, *tb, *tc;
};

// ----------------------------------------------------------------------------------------
// small DFTs
// ----------------------------------------------------------------------------------------
template <typename T, int R>
__device__ __forceinline__ void dft_small(Cx<T> *v, const Cx<T> *tw, int n, bool inv) {
    if (R == 2) {
        Cx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if (R == 4) {
        Cx<T> a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        Cx<T> c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        // multiply d by -i (forward) or +i (inverse)
        Cx<T> dj = inv ? Cx<T>{-d.y, d.x} : Cx<T>{d.y, -d.x};
        v[0] = cadd(a, c);
        v[2] = csub(a, c);
        v[1] = cadd(b, dj);
        v[3] = csub(b, dj);
    } else {
        Cx<T> o[R];
        int step = n / R;
#pragma unroll
        for (int q = 0; q < R; q++) {
            Cx<T> acc = v[0];
#pragma unroll
            for (int r = 1; r < R; r++) {
                Cx<T> w = tw[((q * r) % R) * step];
                if (inv) w = cconj(w);
                acc = cadd(acc, cmul(v[r], w));
            }
            o[q] = acc;
        }
#pragma unroll
        for (int q = 0; q < R; q++) v[q] = o[q];
    }
}

template <typename T, int R>
__device__ __forceinline__ void stockham_pass(const Cx<T> *A, Cx<T> *B, int nlines, int n, int ld,
                                              int Ns, const Cx<T> *tw, bool inv) {
    int nb = n / R;
    int total = nlines * nb;
    int tstep = n / (Ns * R);
    for (int w = threadIdx.x; w < total; w += blockDim.x) {
        int l = w / nb, j = w - l * nb;
        int k = j % Ns;
        Cx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            Cx<T> x = A[l * ld + j + r * nb];
            if (r > 0 && k > 0) {
                Cx<T> t = tw[r * k * tstep];
                if (inv) t = cconj(t);
                x = cmul(x, t);
            }
            v[r] = x;
        }
        dft_small<T, R>(v, tw, n, inv);
        int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) B[l * ld + j0 + r * Ns] = v[r];
    }
}

// In-smem batched FFT of `nlines` lines of length n; returns the buffer holding the result.
template <typename T>
__device__ Cx<T> *fft_lines(Cx<T> *A, Cx<T> *B, int nlines, const PassArgs<T> &a, bool inv) {
    int Ns = 1;
    for (int p = 0; p < a.nrad; p++) {
        int R = a.radix[p];
        switch (R) {
            case 2: stockham_pass<T, 2>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            case 3: stockham_pass<T, 3>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            case 4: stockham_pass<T, 4>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            case 5: stockham_pass<T, 5>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            case 7: stockham_pass<T, 7>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            case 11: stockham_pass<T, 11>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            case 13: stockham_pass<T, 13>(A, B, nlines, a.n, a.ld, Ns, a.tw, inv); break;
            default: break;
        }
        __syncthreads();
        Cx<T> *t = A;
        A = B;
        B = t;
        Ns *= R;
    }
    return A;
}

// DCT-II of the (complex = two real) lines through one complex FFT (Makhoul; the reference
// does the same per real array, utils.jl:89-122).  Input in A in natural order.
template <typename T>
__device__ Cx<T> *dct_fwd_lines(Cx<T> *A, Cx<T> *B, int nlines, const PassArgs<T> &a) {
    int n = a.n, ld = a.ld, h = n / 2;
    // permutation (get_perminds, utils.jl:51-63): z[m] = x[2m], z[n-1-m] = x[2m+1]
    for (int w = threadIdx.x; w < nlines * n; w += blockDim.x) {
        int l = w / n, q = w - l * n;
        int src = q < h ? 2 * q : 2 * (n - 1 - q) + 1;
        B[l * ld + q] = A[l * ld + src];
    }
    __syncthreads();
    Cx<T> *Z = fft_lines<T>(B, A, nlines, a, false);
    Cx<T> *O = (Z == A) ? B : A;
    for (int w = threadIdx.x; w < nlines * n; w += blockDim.x) {
        int l = w / n, k = w - l * n;
        Cx<T> zk = Z[l * ld + k];
        Cx<T> zc = cconj(Z[l * ld + (k == 0 ? 0 : n - k)]);
        // FFT of the real and of the imaginary input sequence
        Cx<T> fa{(zk.x + zc.x) * T(0.5), (zk.y + zc.y) * T(0.5)};
        Cx<T> d{(zk.x - zc.x) * T(0.5), (zk.y - zc.y) * T(0.5)};
        Cx<T> fb{d.y, -d.x};  // d / i
        Cx<T> wk = a.dw[k];
        // X = 2 Re(w_k F_k)
        O[l * ld + k] = Cx<T>{T(2) * (wk.x * fa.x - wk.y * fa.y), T(2) * (wk.x * fb.x - wk.y * fb.y)};
    }
    __syncthreads();
    return O;
}

// DCT-III (inverse of the above), utils.jl:124-155.
template <typename T>
__device__ Cx<T> *dct_inv_lines(Cx<T> *A, Cx<T> *B, int nlines, const PassArgs<T> &a) {
    int n = a.n, ld = a.ld, h = n / 2;
    for (int w = threadIdx.x; w < nlines * n; w += blockDim.x) {
        int l = w / n, k = w - l * n;
        Cx<T> xk = A[l * ld + k];
        Cx<T> out;
        if (k == 0) {
            T w0 = a.dwi[0].x;
            out = Cx<T>{w0 * xk.x, w0 * xk.y};
        } else {
            Cx<T> xm = A[l * ld + n - k];
            Cx<T> wk = a.dwi[k], wm = cconj(a.dwi[n - k]);
            // H[k] = (winv_k x[k] + conj(winv_{n-k}) x[n-k]) / 2 for the two real sequences
            Cx<T> ha{(wk.x * xk.x + wm.x * xm.x) * T(0.5), (wk.y * xk.x + wm.y * xm.x) * T(0.5)};
            Cx<T> hb{(wk.x * xk.y + wm.x * xm.y) * T(0.5), (wk.y * xk.y + wm.y * xm.y) * T(0.5)};
            out = Cx<T>{ha.x - hb.y, ha.y + hb.x};  // ha + i hb
        }
        B[l * ld + k] = out;
    }
    __syncthreads();
    Cx<T> *Y = fft_lines<T>(B, A, nlines, a, true);
    Cx<T> *O = (Y == A) ? B : A;
    T invn = T(1) / T(n);  // ifft! is normalised (utils.jl:132)
    for (int w = threadIdx.x; w < nlines * n; w += blockDim.x) {
        int l = w / n, q = w - l * n;
        // perminv (utils.jl:59-60): out[2m] = y[m], out[2m+1] = y[n-1-m]
        int src = (q & 1) ? n - 1 - (q >> 1) : (q >> 1);
        Cx<T> y = Y[l * ld + src];
        O[l * ld + q] = Cx<T>{y.x * invn, y.y * invn};
    }
    __syncthreads();
    (void)h;
    return O;
}

// ----------------------------------------------------------------------------------------
// the pass kernel
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) line_pass_kernel(PassArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *A = reinterpret_cast<Cx<T> *>(smem_raw);
    Cx<T> *B = A + (size_t)a.tile * a.ld;

    const int dim = a.dim;
    const int da = dim == 0 ? 1 : 0;                 // fast companion dimension
    const int db = dim == 2 ? 1 : 2;                 // slow companion dimension
    const int ma = a.m[da], mb = (a.D == 3) ? a.m[db] : 1;
    const long long nlines_total = (long long)ma * mb;
    const long long line0 = (long long)blockIdx.x * a.tile;
    int nl = (int)min((long long)a.tile, nlines_total - line0);
    if (nl <= 0) return;

    const int n = a.n, ld = a.ld;
    const int nin = a.half_in ? n / 2 + 1 : n;

    // ---- load ----
    const int total_in = nl * nin;
    for (int w = threadIdx.x; w < total_in; w += blockDim.x) {
        int l, k;
        if (dim == 0) {
            l = w / nin;
            k = w - l * nin;
        } else {
            k = w / nl;
            l = w - k * nl;
        }
        long long line = line0 + l;
        int ia = (int)(line % ma), ib = (int)(line / ma);
        Cx<T> v;
        if (a.src_real) {
            long long off = k * a.rs[dim] + ia * a.rs[da] + (a.D == 3 ? ib * a.rs[db] : 0);
            T r = a.R[off];
            if (a.prescale) {
                int ip = a.pdim == dim ? k : (a.pdim == da ? ia : ib);
                r *= a.prescale[ip];
            }
            v = Cx<T>{r, T(0)};
        } else {
            long long off = k * a.ws[dim] + ia * a.ws[da] + (a.D == 3 ? ib * a.ws[db] : 0);
            v = a.W[off];
        }
        A[l * ld + k] = v;
    }
    __syncthreads();
    if (a.half_in) {  // Hermitian extension X[n-k] = conj(X[k])
        int hx = n - nin;
        for (int w = threadIdx.x; w < nl * hx; w += blockDim.x) {
            int l = w / hx, k = nin + (w - l * hx);
            A[l * ld + k] = cconj(A[l * ld + n - k]);
        }
        __syncthreads();
    }

    // ---- transform ----
    Cx<T> *Rbuf = A;
    Cx<T> *Obuf = B;
    switch (a.op) {
        case P_FFT_FWD: Rbuf = fft_lines<T>(A, B, nl, a, false); break;
        case P_FFT_INV: Rbuf = fft_lines<T>(A, B, nl, a, true); break;
        case P_DCT_FWD: Rbuf = dct_fwd_lines<T>(A, B, nl, a); break;
        case P_DCT_INV: Rbuf = dct_inv_lines<T>(A, B, nl, a); break;
        case P_MID_FFT:
        case P_MID_DCT: {
            Rbuf = a.op == P_MID_FFT ? fft_lines<T>(A, B, nl, a, false) : dct_fwd_lines<T>(A, B, nl, a);
            Obuf = (Rbuf == A) ? B : A;
            for (int w = threadIdx.x; w < nl * n; w += blockDim.x) {
                int l = w / n, k = w - l * n;
                long long line = line0 + l;
                int ia = (int)(line % ma), ib = (int)(line / ma);
                int K[3] = {0, 0, 0};
                K[dim] = k;
                K[da] = ia;
                if (a.D == 3) K[db] = ib;
                T den = a.eig[0][K[0]] + a.eig[1][K[1]];
                if (a.D == 3) den = den + a.eig[2][K[2]];
                bool zero = (K[0] | K[1] | K[2]) == 0;
                T f = zero ? T(0) : a.sign * a.scale / den;
                Cx<T> v = Rbuf[l * ld + k];
                Rbuf[l * ld + k] = Cx<T>{v.x * f, v.y * f};
            }
            __syncthreads();
            Rbuf = a.op == P_MID_FFT ? fft_lines<T>(Rbuf, Obuf, nl, a, true)
                                     : dct_inv_lines<T>(Rbuf, Obuf, nl, a);
            break;
        }
        case P_THOMAS: {
            // batched Thomas recurrence, one line per thread (pressure.jl:480-504), then the
            // zero-mode mean removal (:531-532) and the 1/fftscale of :538
            T *wl = reinterpret_cast<T *>(B);
            for (int l = threadIdx.x; l < nl; l += blockDim.x) {
                long long line = line0 + l;
                int ia = (int)(line % ma), ib = (int)(line / ma);
                T lam = T(0);
                bool zero = true;
                for (int be = 0; be < a.D; be++) {
                    if (be == dim) continue;
                    int kb = be == da ? ia : ib;
                    lam += a.eig[be][kb];
                    zero = zero && kb == 0;
                }
                Cx<T> *f = A + l * ld;
                T *w = wl + l * 2 * ld;
                if (lam == T(0)) {
                    w[0] = T(0);
                    f[0] = Cx<T>{T(0), T(0)};
                } else {
                    T bb = a.tb[0] + lam;
                    w[0] = a.tc[0] / bb;
                    f[0] = Cx<T>{f[0].x / bb, f[0].y / bb};
                }
                for (int j = 1; j < n; j++) {
                    T aj = a.ta[j];
                    T denom = a.tb[j] + lam - aj * w[j - 1];
                    w[j] = a.tc[j] / denom;
                    f[j] = Cx<T>{(f[j].x - aj * f[j - 1].x) / denom, (f[j].y - aj * f[j - 1].y) / denom};
                }
                for (int j = n - 2; j >= 0; j--) {
                    f[j] = Cx<T>{f[j].x - w[j] * f[j + 1].x, f[j].y - w[j] * f[j + 1].y};
                }
                if (zero) {
                    Cx<T> s{T(0), T(0)};
                    for (int j = 0; j < n; j++) s = cadd(s, f[j]);
                    s = Cx<T>{s.x / T(n), s.y / T(n)};
                    for (int j = 0; j < n; j++) f[j] = csub(f[j], s);
                }
                for (int j = 0; j < n; j++) f[j] = Cx<T>{f[j].x * a.scale, f[j].y * a.scale};
            }
            __syncthreads();
            Rbuf = A;
            break;
        }
    }

    // ---- store ----
    const int nout = a.half_out ? n / 2 + 1 : n;
    const int total_out = nl * nout;
    for (int w = threadIdx.x; w < total_out; w += blockDim.x) {
        int l, k;
        if (dim == 0) {
            l = w / nout;
            k = w - l * nout;
        } else {
            k = w / nl;
            l = w - k * nl;
        }
        long long line = line0 + l;
        int ia = (int)(line % ma), ib = (int)(line / ma);
        Cx<T> v = Rbuf[l * ld + k];
        if (a.dst_real) {
            long long off = k * a.rs[dim] + ia * a.rs[da] + (a.D == 3 ? ib * a.rs[db] : 0);
            a.R[off] = v.x;
        } else {
            long long off = k * a.ws[dim] + ia * a.ws[da] + (a.D == 3 ? ib * a.ws[db] : 0);
            a.W[off] = v;
        }
    }
}

// ----------------------------------------------------------------------------------------
// Thomas pass of psolver_tridiagonal for the register-FFT fast path (pressure.jl:480-504,
// 531-538): one thread per Fourier mode (kx, ko) marches along the wall-normal index j;
// consecutive threads own consecutive kx, so every access of the sweep is a coalesced row.
// w_j goes to a scratch array for the back substitution, like `thomas_w` in the reference.
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
    thomas_march_kernel(Cx<T> *__restrict__ W, T *__restrict__ wbuf, int m0, int n, long long sd, long long so,
                        const T *__restrict__ ta, const T *__restrict__ tb, const T *__restrict__ tc,
                        const T *__restrict__ eig0, const T *__restrict__ eigo, T scale, int kx_off, int m0_glob) {
    // kx: column inside this rank's block of kx modes (kx_off = first global mode of the block)
    const int kx = blockIdx.x * blockDim.x + threadIdx.x, ko = blockIdx.y;
    if (kx >= m0 || kx + kx_off >= m0_glob) return;
    const T lam = eig0[kx + kx_off] + eigo[ko];
    Cx<T> *f = W + kx + ko * so;
    T *w = wbuf + kx + ko * so;
    Cx<T> fp;
    T wp;
    {
        Cx<T> f0 = f[0];
        if (lam == T(0)) {
            wp = T(0);
            fp = Cx<T>{T(0), T(0)};
        } else {
            T bb = tb[0] + lam;
            wp = tc[0] / bb;
            fp = Cx<T>{f0.x / bb, f0.y / bb};
        }
        w[0] = wp;
        f[0] = fp;
    }
    // rows are processed in chunks of CH with the next chunk's loads in flight (the sweeps are
    // latency chains; bandwidth comes from the other modes and from this prefetch)
    constexpr int CH = 4;
    Cx<T> cur[CH], nxt[CH];
#pragma unroll
    for (int i = 0; i < CH; i++)
        if (1 + i < n) nxt[i] = f[(long long)(1 + i) * sd];
    for (int j0 = 1; j0 < n; j0 += CH) {
#pragma unroll
        for (int i = 0; i < CH; i++) cur[i] = nxt[i];
#pragma unroll
        for (int i = 0; i < CH; i++)
            if (j0 + CH + i < n) nxt[i] = f[(long long)(j0 + CH + i) * sd];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int j = j0 + i;
            if (j < n) {
                T aj = ta[j];
                T denom = tb[j] + lam - aj * wp;
                wp = tc[j] / denom;
                fp = Cx<T>{(cur[i].x - aj * fp.x) / denom, (cur[i].y - aj * fp.y) / denom};
                w[(long long)j * sd] = wp;
                f[(long long)j * sd] = fp;
            }
        }
    }
    const bool zero = kx + kx_off == 0 && ko == 0;
    // back substitution; fp holds f[n-1]
    if (!zero) f[(long long)(n - 1) * sd] = Cx<T>{fp.x * scale, fp.y * scale};
    T wc[CH], wn[CH];
#pragma unroll
    for (int i = 0; i < CH; i++)
        if (n - 2 - i >= 0) {
            nxt[i] = f[(long long)(n - 2 - i) * sd];
            wn[i] = w[(long long)(n - 2 - i) * sd];
        }
    for (int j0 = n - 2; j0 >= 0; j0 -= CH) {
#pragma unroll
        for (int i = 0; i < CH; i++) cur[i] = nxt[i], wc[i] = wn[i];
#pragma unroll
        for (int i = 0; i < CH; i++)
            if (j0 - CH - i >= 0) {
                nxt[i] = f[(long long)(j0 - CH - i) * sd];
                wn[i] = w[(long long)(j0 - CH - i) * sd];
            }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int j = j0 - i;
            if (j >= 0) {
                fp = Cx<T>{cur[i].x - wc[i] * fp.x, cur[i].y - wc[i] * fp.y};
                f[(long long)j * sd] = zero ? fp : Cx<T>{fp.x * scale, fp.y * scale};
            }
        }
    }
    if (zero) {  // remove the mean of the zero mode (pressure.jl:531-532), then the 1/fftscale
        Cx<T> s{T(0), T(0)};
        for (int j = 0; j < n; j++) s = cadd(s, f[(long long)j * sd]);
        s = Cx<T>{s.x / T(n), s.y / T(n)};
        for (int j = 0; j < n; j++) {
            Cx<T> v = csub(f[(long long)j * sd], s);
            f[(long long)j * sd] = Cx<T>{v.x * scale, v.y * scale};
        }
    }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
struct Poisson {
    int D = 0, kind = 0;
    bool f64 = false;
    int n[3] = {1, 1, 1};     // Np
    int ioff[3] = {0, 0, 0};  // first pressure DOF
    int Ng[3] = {1, 1, 1};
    int m[3] = {1, 1, 1};     // complex extents
    long long wstride[3] = {1, 1, 1};
    int rdim = -1, wdir = -1;
    bool per[3] = {false, false, false};
    void *W = nullptr;
    std::vector<void *> owned;
    long long bytes = 0;
    // per-dimension tables (device)
    void *tw[3] = {nullptr, nullptr, nullptr};
    void *dw[3] = {nullptr, nullptr, nullptr};
    void *dwi[3] = {nullptr, nullptr, nullptr};
    void *eig[3] = {nullptr, nullptr, nullptr};
    void *ta = nullptr, *tb = nullptr, *tc = nullptr, *invO = nullptr;
    double sign = 1, scale = 1;
    std::vector<int> radix[3];
    struct PassDesc {
        int op, dim, src_real, dst_real, half_out, half_in;
    };
    std::vector<PassDesc> passes;
    int max_smem = 0;
    bool fast = false;  // power-of-two all-periodic path (fft_fast.cu)
    bool fast_tri = false;  // tridiagonal: register FFTs in the periodic directions + marching Thomas pass
    bool fast_dct = false;  // transform: x periodic (R2C) + register FFT / DCT passes in y (and z)
    void *thomas_w = nullptr;
    int tile = 8;
    // z-slab decomposition (world > 1): W holds this rank's nzl planes in natural layout;
    // S / Rv are the send / receive sides of the all-to-all transposes (comm.cu), laid out as
    // [world blocks][nzl planes][n1/world ky rows][m0p] and, equivalently on the receiving
    // side, [n3 planes][n1/world ky rows][m0p].
    Comm *comm = nullptr;
    int nzl = 0, kyl = 0;
    int kxl = 0;  // decomposed psolver_tridiagonal: kx modes per rank after the transpose
    void *S = nullptr, *Rv = nullptr;
    // direct peer-store transposes (NVLink P2P through CUDA IPC): peers' S / Rv / flag buffers
    // mapped into this process; flags = 3 rows x 8 slots (row 0: forward transpose done,
    // row 1: backward transpose done, row 2: wait time-out marker)
    bool p2p = false;
    unsigned *flags = nullptr;
    void *peerS[8] = {}, *peerRv[8] = {};
    unsigned *peerFlags[8] = {};
    unsigned epoch = 0;
    CgSolver *cg = nullptr;  // kind == INSB200_PSOLVER_CG: everything else in this struct is unused
    size_t wbytes = 0;
    size_t halo_slot = 0;
};

static bool factorize(int n, std::vector<int> &rad) {
    rad.clear();
    while (n % 4 == 0) {
        rad.push_back(4);
        n /= 4;
    }
    const int primes[6] = {2, 3, 5, 7, 11, 13};
    for (int p : primes)
        while (n % p == 0) {
            rad.push_back(p);
            n /= p;
        }
    return n == 1 && (int)rad.size() <= MAXRAD;
}

static long double sinpi_ld(long double x) { return sinl(3.14159265358979323846264338327950288L * x); }
static long double cospi_ld(long double x) { return cosl(3.14159265358979323846264338327950288L * x); }

template <typename T>
static int upload(Poisson *ps, const std::vector<T> &h, void **dptr) {
    size_t b = h.size() * sizeof(T);
    INSB_CUDA(cudaMalloc(dptr, b ? b : sizeof(T)));
    ps->owned.push_back(*dptr);
    ps->bytes += (long long)b;
    INSB_CUDA(cudaMemcpy(*dptr, h.data(), b, cudaMemcpyHostToDevice));
    return 0;
}

template <typename T>
static int build_tables(Poisson *ps, const insb200_setup_t &s, const int Ip[3][2]) {
    const int D = ps->D;
    const T *dl[3], *du[3];
    for (int a = 0; a < D; a++) {
        dl[a] = (const T *)s.Delta[a];
        du[a] = (const T *)s.DeltaU[a];
    }
    // twiddles
    for (int a = 0; a < D; a++) {
        int n = ps->n[a];
        std::vector<Cx<T>> tw(n), dw(n), dwi(n);
        for (int q = 0; q < n; q++) {
            long double x = (long double)(2 * q) / n;
            tw[q] = Cx<T>{(T)cospi_ld(x), (T)(-sinpi_ld(x))};
            // w_k = exp(-i pi k / (2n)) / sqrt(2n), w_0 /= sqrt(2)       (utils.jl:68-75)
            // winv_k = exp(+i pi k / (2n)) * sqrt(2n), winv_0 *= sqrt(2)/2 (utils.jl:76-83)
            long double y = (long double)q / (2.0L * n);
            long double nrm = sqrtl(2.0L * n);
            long double c = cospi_ld(y), sn = sinpi_ld(y);
            long double f0 = q == 0 ? 1.0L / sqrtl(2.0L) : 1.0L;
            long double g0 = q == 0 ? sqrtl(2.0L) / 2.0L : 1.0L;
            dw[q] = Cx<T>{(T)(c / nrm * f0), (T)(-sn / nrm * f0)};
            dwi[q] = Cx<T>{(T)(c * nrm * g0), (T)(sn * nrm * g0)};
        }
        INSB_TRY(upload<Cx<T>>(ps, tw, &ps->tw[a]));
        if (!ps->per[a] && ps->kind == INSB200_PSOLVER_TRANSFORM) {
            INSB_TRY(upload<Cx<T>>(ps, dw, &ps->dw[a]));
            INSB_TRY(upload<Cx<T>>(ps, dwi, &ps->dwi[a]));
        }
    }
    // eigenvalues
    double nper = 1;
    for (int a = 0; a < D; a++)
        if (ps->per[a]) nper *= ps->n[a];
    ps->scale = 1.0 / nper;
    if (ps->kind == INSB200_PSOLVER_SPECTRAL) {
        // pressure.jl:327-342: Δx = first.(Δ) (ghost width), Float64 evaluation rounded into T
        T dx[3], om = T(1);
        for (int a = 0; a < D; a++) {
            dx[a] = dl[a][0];
            om = a == 0 ? dx[a] : om * dx[a];
        }
        for (int a = 0; a < D; a++) {
            int cnt = ps->n[a];
            std::vector<T> e(cnt);
            for (int k = 0; k < cnt; k++) {
                if (s.eig[a]) {
                    int kmax = a == 0 ? ps->n[a] / 2 + 1 : ps->n[a];
                    e[k] = k < kmax ? ((const T *)s.eig[a])[k] : T(0);
                    continue;
                }
                double sp = (double)sinpi_ld((long double)((double)k / (double)ps->n[a]));
                e[k] = (T)(((double)(T(4) * om) * (sp * sp)) / (double)(dx[a] * dx[a]));
            }
            INSB_TRY(upload<T>(ps, e, &ps->eig[a]));
        }
        ps->sign = -1;
    } else if (ps->kind == INSB200_PSOLVER_TRANSFORM) {
        // pressure.jl:221-246
        T dx[3], om = T(1);
        for (int a = 0; a < D; a++) {
            dx[a] = (T)(((T)s.xlims[a][1] - (T)s.xlims[a][0]) / (T)ps->n[a]);
            om = a == 0 ? dx[a] : om * dx[a];
        }
        for (int a = 0; a < D; a++) {
            int cnt = ps->n[a];
            std::vector<T> e(cnt);
            for (int k = 0; k < cnt; k++) {
                if (s.eig[a]) {
                    e[k] = ((const T *)s.eig[a])[k];
                    continue;
                }
                double x = (double)k / (double)ps->n[a];
                double h2 = (double)(dx[a] * dx[a]);
                if (ps->per[a]) {
                    double sp = (double)sinpi_ld((long double)x);
                    e[k] = (T)(((double)(T(-4) * om) * (sp * sp)) / h2);
                } else {
                    double cp = (double)cospi_ld((long double)x);
                    e[k] = (T)(((double)(T(2) * om) * (cp - 1.0)) / h2);
                }
            }
            INSB_TRY(upload<T>(ps, e, &ps->eig[a]));
        }
        ps->sign = 1;
    } else {  // tridiagonal, pressure.jl:429-466
        int w = ps->wdir, n = ps->n[w];
        T h[3];
        for (int a = 0; a < D; a++) h[a] = dl[a][Ip[a][0]];
        std::vector<T> ta(n, T(0)), tb(n, T(0)), tc(n, T(0)), io(n);
        T hper = T(1);
        for (int a = 0; a < D; a++) hper = hper * (a == w ? T(1) : h[a]);
        for (int j = 0; j < n; j++) {
            int i = Ip[w][0] + j;
            if (j != 0) ta[j] = T(1) / (dl[w][i] * du[w][i - 1]);
            if (j != n - 1) tc[j] = T(1) / (dl[w][i] * du[w][i]);
            tb[j] = -(ta[j] + tc[j]);
            io[j] = T(1) / (hper * dl[w][i]);
        }
        INSB_TRY(upload<T>(ps, ta, &ps->ta));
        INSB_TRY(upload<T>(ps, tb, &ps->tb));
        INSB_TRY(upload<T>(ps, tc, &ps->tc));
        INSB_TRY(upload<T>(ps, io, &ps->invO));
        for (int a = 0; a < D; a++) {
            int cnt = a == w ? 1 : ps->n[a];
            std::vector<T> e(cnt, T(0));
            if (a != w) {
                for (int k = 0; k < cnt; k++) {
                    if (s.eig[a]) {
                        e[k] = ((const T *)s.eig[a])[k];
                        continue;
                    }
                    // evaluated in T (pressure.jl:457-458): argument rounded to T first
                    T arg = (T)k / (T)ps->n[a];
                    T sp = (T)sinpi_ld((long double)arg);
                    e[k] = T(-4) * (sp * sp) / (h[a] * h[a]);
                }
            }
            INSB_TRY(upload<T>(ps, e, &ps->eig[a]));
        }
        ps->sign = 1;
    }
    return 0;
}

int poisson_create(const insb200_setup_t &s, const int Ip[3][2], int kind, Comm *comm, Poisson **out) {
    Poisson *ps = new Poisson();
    ps->comm = comm;
    ps->D = s.D;
    ps->kind = kind;
    ps->f64 = s.dtype == INSB200_F64;
    const int D = s.D;
    for (int a = 0; a < D; a++) {
        ps->Ng[a] = s.N[a];
        ps->n[a] = Ip[a][1] - Ip[a][0];
        ps->ioff[a] = Ip[a][0];
        ps->per[a] = s.bc_u[a][0] == INSB200_BC_PERIODIC && s.bc_u[a][1] == INSB200_BC_PERIODIC;
    }
    auto fail = [&](const std::string &m) {
        set_error(m);
        poisson_destroy(ps);
        return 1;
    };
    if (kind == INSB200_PSOLVER_CG) {
        if (comm_world(comm) > 1) return fail("psolver_cg is not decomposed: world > 1 needs an FFT-type solver");
        if (cg_create(s, Ip, s.cg_reltol, s.cg_abstol, s.cg_maxiter, &ps->cg)) {
            poisson_destroy(ps);
            return 1;
        }
        ps->bytes = cg_bytes(ps->cg);
        *out = ps;
        return 0;
    }
    // preconditions (utils.jl:4-15, pressure.jl:225-229, :406-435, utils.jl:53)
    for (int a = 0; a < D; a++) {
        bool wall = s.bc_u[a][0] == INSB200_BC_DIRICHLET && s.bc_u[a][1] == INSB200_BC_DIRICHLET;
        if (kind == INSB200_PSOLVER_SPECTRAL && !ps->per[a])
            return fail("Spectral psolver requires periodic boundary conditions.");
        if (kind == INSB200_PSOLVER_SPECTRAL && ps->n[a] % 2)
            return fail("Spectral psolver requires even number of volumes.");
        if (kind == INSB200_PSOLVER_TRANSFORM && !(ps->per[a] || wall))
            return fail("psolver_transform: only PeriodicBC / DirichletBC are supported");
        if (kind == INSB200_PSOLVER_TRANSFORM && ps->n[a] % 2)
            return fail("Only even grids supported");
    }
    if (kind == INSB200_PSOLVER_TRIDIAGONAL) {
        int w = s.tridiagonal_dir;
        if (w < 0) {
            for (int a = 0; a < D && w < 0; a++)
                if (s.bc_u[a][0] == INSB200_BC_DIRICHLET && s.bc_u[a][1] == INSB200_BC_DIRICHLET) w = a;
            if (w < 0)
                return fail("psolver_tridiagonal: no wall-bounded direction (DirichletBC on both sides) found");
        }
        if (w >= D || !(s.bc_u[w][0] == INSB200_BC_DIRICHLET && s.bc_u[w][1] == INSB200_BC_DIRICHLET))
            return fail("psolver_tridiagonal: direction is not wall-bounded (needs DirichletBC on both sides)");
        for (int a = 0; a < D; a++)
            if (a != w && !ps->per[a])
                return fail("psolver_tridiagonal: all directions except the wall-normal one must be periodic");
        ps->wdir = w;
    }
    // uniform spacing where the solver needs it
    for (int a = 0; a < D; a++) {
        if (kind == INSB200_PSOLVER_TRIDIAGONAL && a == ps->wdir) continue;
        double ref = ps->f64 ? ((const double *)s.Delta[a])[Ip[a][0]] : ((const float *)s.Delta[a])[Ip[a][0]];
        double tol = std::sqrt(ps->f64 ? 2.220446049250313e-16 : 1.1920929e-07);
        for (int i = Ip[a][0]; i < Ip[a][1]; i++) {
            double v = ps->f64 ? ((const double *)s.Delta[a])[i] : ((const float *)s.Delta[a])[i];
            if (std::fabs(v - ref) > tol * std::fmax(std::fabs(v), std::fabs(ref)))
                return fail("FFT-type psolver requires uniform grid spacing in the transformed directions.");
        }
    }
    // R2C dimension: first periodic one
    for (int a = 0; a < D && ps->rdim < 0; a++)
        if (ps->per[a]) ps->rdim = a;
    for (int a = 0; a < D; a++) {
        ps->m[a] = (a == ps->rdim) ? ps->n[a] / 2 + 1 : ps->n[a];
        if (a == ps->rdim && ps->n[a] % 2)
            return fail("FFT-type psolver requires an even number of volumes in the first periodic direction.");
        if (!(kind == INSB200_PSOLVER_TRIDIAGONAL && a == ps->wdir))
            if (!factorize(ps->n[a], ps->radix[a]))
                return fail("unsupported transform length (prime factors must be in {2,3,5,7,11,13})");
    }
    long long pitch = (ps->m[0] + 7) / 8 * 8;
    const int world = comm_world(comm);
    ps->nzl = D == 3 ? ps->n[2] / world : 1;
    ps->kyl = ps->n[1] / world;
    if (world > 1) {
        bool spec = kind == INSB200_PSOLVER_SPECTRAL && D == 3;
        for (int a = 0; a < D; a++) spec = spec && fast_fft_supported(ps->n[a]);
        // psolver_tridiagonal with walls in y: x R2C on the z-slab, transpose to blocks of kx
        // modes (all y, all z), z FFT + Thomas along y + inverse z FFT there, transpose back
        const bool tri = kind == INSB200_PSOLVER_TRIDIAGONAL && D == 3 && ps->wdir == 1 &&
                         fast_fft_supported(ps->n[0]) && fast_fft_supported(ps->n[2]) && ps->n[1] % 2 == 0;
        // psolver_transform (x periodic, DCT along the wall-bounded directions): the same pass structure and
        // transposes as psolver_spectral with DCT kernels in y and / or z
        bool dct = kind == INSB200_PSOLVER_TRANSFORM && D == 3 && ps->per[0];
        for (int a = 0; a < D; a++) dct = dct && fast_fft_supported(ps->n[a]);
        if (!spec && !tri && !dct)
            return fail("world > 1: only psolver_spectral / psolver_transform (x periodic) on 3-D power-of-two grids "
                        "(32..2048 per dim) and psolver_tridiagonal with walls in y (power-of-two x and z extents) "
                        "are decomposed");
        spec = spec || dct;  // (same divisibility requirements)
        if (ps->n[2] % world || (ps->nzl & (ps->nzl - 1)))
            return fail("world > 1: n3 must be divisible by the number of ranks (power-of-two planes per rank)");
        if (spec && (ps->n[1] % world || (ps->kyl & (ps->kyl - 1))))
            return fail("world > 1: n2 must be divisible by the number of ranks (power of two)");
        if (tri) {  // whole line tiles per rank: pad the half spectrum to a multiple of 8 * world
            pitch = (ps->m[0] + 8 * world - 1) / (8 * world) * (8 * world);
            ps->kxl = (int)(pitch / world);
        }
    }
    ps->wstride[0] = 1;
    ps->wstride[1] = pitch;
    ps->wstride[2] = pitch * ps->m[1];
    size_t esz = ps->f64 ? 16 : 8;
    size_t wbytes = (size_t)pitch * ps->m[1] * (D == 3 ? (world > 1 ? ps->nzl : ps->m[2]) : 1) * esz;
    if (cudaMalloc(&ps->W, wbytes) != cudaSuccess) return fail("cudaMalloc of the spectral workspace failed");
    cudaMemset(ps->W, 0, wbytes);
    ps->bytes += (long long)wbytes;
    if (world > 1) {
        if (cudaMalloc(&ps->S, wbytes) != cudaSuccess || cudaMalloc(&ps->Rv, wbytes) != cudaSuccess)
            return fail("cudaMalloc of the transpose buffers failed");
        ps->owned.push_back(ps->S);
        ps->owned.push_back(ps->Rv);
        cudaMemset(ps->S, 0, wbytes);
        cudaMemset(ps->Rv, 0, wbytes);
        ps->bytes += 2 * (long long)wbytes;
        ps->wbytes = wbytes;
        // flag words of the transposes + the window of the peer-store halo exchange (comm.cu)
        // (largest exchange: two padded planes of the three components of the fused path, fused.h)
        ps->halo_slot = (size_t)6 * ((ps->Ng[0] + 9) / 4 * 4) * (ps->Ng[1] + 2) * esz / 2;  // esz = complex size here
        const size_t fb = comm_halo_window_bytes(ps->halo_slot);
        if (cudaMalloc((void **)&ps->flags, fb) != cudaSuccess)
            return fail("cudaMalloc of the peer flags / halo window failed");
        ps->owned.push_back(ps->flags);
        cudaMemset(ps->flags, 0, fb);
        ps->bytes += (long long)fb;
    }
    int r = ps->f64 ? build_tables<double>(ps, s, Ip) : build_tables<float>(ps, s, Ip);
    if (r) {
        poisson_destroy(ps);
        return r;
    }
    // pass list: R2C(rdim) -> FFT over the other periodic dims -> wall dims; the last forward
    // pass is fused with the eigenvalue division and its own inverse; then mirror back.
    std::vector<int> order;
    if (ps->rdim >= 0) order.push_back(ps->rdim);
    for (int a = 0; a < D; a++)
        if (ps->per[a] && a != ps->rdim) order.push_back(a);
    for (int a = 0; a < D; a++)
        if (!ps->per[a]) order.push_back(a);
    int np = (int)order.size();
    for (int i = 0; i < np; i++) {
        int a = order[i];
        bool first = i == 0, last = i == np - 1;
        Poisson::PassDesc pd{};
        pd.dim = a;
        pd.src_real = first;
        pd.half_out = first && a == ps->rdim;
        if (last) {
            pd.op = kind == INSB200_PSOLVER_TRIDIAGONAL ? P_THOMAS : (ps->per[a] ? P_MID_FFT : P_MID_DCT);
        } else {
            pd.op = ps->per[a] ? P_FFT_FWD : P_DCT_FWD;
        }
        ps->passes.push_back(pd);
    }
    for (int i = np - 2; i >= 0; i--) {
        int a = order[i];
        Poisson::PassDesc pd{};
        pd.dim = a;
        pd.op = ps->per[a] ? P_FFT_INV : P_DCT_INV;
        pd.dst_real = i == 0;
        pd.half_in = i == 0 && a == ps->rdim;
        ps->passes.push_back(pd);
    }
    ps->fast = kind == INSB200_PSOLVER_SPECTRAL && getenv("INSB200_GENERIC_FFT") == nullptr;
    for (int a = 0; a < D; a++) ps->fast = ps->fast && fast_fft_supported(ps->n[a]);
    // psolver_tridiagonal fast path: 3-D, walls in y or z, power-of-two periodic extents
    ps->fast_tri = kind == INSB200_PSOLVER_TRIDIAGONAL && D == 3 && ps->wdir >= 1 &&
                   (world == 1 ? getenv("INSB200_GENERIC_FFT") == nullptr : ps->wdir == 1) && ps->n[1] % 2 == 0;
    for (int a = 0; a < D; a++)
        if (a != ps->wdir) ps->fast_tri = ps->fast_tri && fast_fft_supported(ps->n[a]);
    if (ps->fast_tri) {
        size_t wb = (size_t)pitch * ps->m[1] * (world > 1 ? ps->nzl : ps->m[2]) * (ps->f64 ? 8 : 4);
        if (cudaMalloc(&ps->thomas_w, wb) != cudaSuccess) return fail("cudaMalloc of thomas_w failed");
        ps->owned.push_back(ps->thomas_w);
        ps->bytes += (long long)wb;
    }
    // psolver_transform fast path: x periodic, every extent a supported power of two
    ps->fast_dct = kind == INSB200_PSOLVER_TRANSFORM && ps->per[0] && D >= 2 &&
                   (world > 1 || getenv("INSB200_GENERIC_FFT") == nullptr) && ps->n[1] % 2 == 0;
    for (int a = 0; a < D; a++) ps->fast_dct = ps->fast_dct && fast_fft_supported(ps->n[a]);
    ps->tile = ps->f64 ? 4 : 8;
    if (const char *e = getenv("INSB200_FFT_TILE")) ps->tile = atoi(e) > 0 ? atoi(e) : ps->tile;
    if (ps->tile > 8) ps->tile = 8;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&ps->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    *out = ps;
    return 0;
}

// First half of a collective tear-down: unmap the peers' buffers (the exporting process must not free memory
// another process still has open, CUDA IPC); from here on the transposes / halo exchange use NCCL again.
void poisson_ipc_close(Poisson *p) {
    if (!p) return;
    if (p->comm && p->flags) comm_disable_p2p_halo(p->comm, p->flags);
    if (p->p2p) {
        cudaDeviceSynchronize();
        const int rank = comm_rank(p->comm), world = comm_world(p->comm);
        for (int s = 0; s < world; s++) {
            if (s == rank) continue;
            if (p->peerS[s]) cudaIpcCloseMemHandle(p->peerS[s]);
            if (p->peerRv[s]) cudaIpcCloseMemHandle(p->peerRv[s]);
            if (p->peerFlags[s]) cudaIpcCloseMemHandle(p->peerFlags[s]);
            p->peerS[s] = p->peerRv[s] = nullptr;
            p->peerFlags[s] = nullptr;
        }
        p->p2p = false;
    }
}

void poisson_destroy(Poisson *p) {
    if (!p) return;
    if (p->cg) cg_destroy(p->cg);
    poisson_ipc_close(p);
    for (void *q : p->owned) cudaFree(q);
    if (p->W) cudaFree(p->W);
    delete p;
}

long long poisson_bytes(const Poisson *p) { return p ? p->bytes : 0; }
int poisson_iterations(const Poisson *p) { return p && p->cg ? cg_iterations(p->cg) : 0; }

// ----------------------------------------------------------------------------------------
// energyspectrum (src/spectral.jl:75-117) with the solver's own FFT passes: per velocity
// component x R2C -> y -> z forward transforms into the half-spectrum workspace, then
// |u_hat|^2 / (2 N^2) is binned into the integer shells kappa <= |k| < kappa + 1, the modes
// whose conjugate partner is not stored (0 < kx < n0/2) counted twice (spectral.jl:36-66).
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) spectrum_bin_kernel(const Cx<T> *__restrict__ W, int m0, int m0p, int n0, int n1, int n2,
                                                           int kmax, double scale, double *ehat) {
    extern __shared__ double bins[];
    for (int i = threadIdx.x; i <= kmax; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
    const long long total = (long long)m0 * n1 * n2;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(q % m0);
        const int iy = (int)((q / m0) % n1), iz = (int)(q / ((long long)m0 * n1));
        const int ky = iy < (n1 + 1) / 2 ? iy : iy - n1;  // fftfreq_int (spectral.jl:7)
        const int kz = n2 > 1 ? (iz < (n2 + 1) / 2 ? iz : iz - n2) : 0;
        const long long k2 = (long long)kx * kx + (long long)ky * ky + (long long)kz * kz;
        int kap = (int)sqrt((double)k2);
        while ((long long)kap * kap > k2) kap--;
        while ((long long)(kap + 1) * (kap + 1) <= k2) kap++;
        if (kap > kmax) continue;
        const Cx<T> v = W[kx + (long long)iy * m0p + (long long)iz * m0p * n1];
        const double w = (kx > 0 && kx < n0 / 2) ? 2.0 : 1.0;
        atomicAdd(&bins[kap], w * scale * ((double)v.x * (double)v.x + (double)v.y * (double)v.y));
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= kmax; i += blockDim.x)
        if (bins[i] != 0.0) atomicAdd(ehat + i, bins[i]);
}

// Spectral sums of turbulence_statistics (src/spectral.jl:89-90, :121-135, :166-257): mom[j] += sum over all
// wavenumbers (conjugate modes counted) of |2 pi k / L|^(2 j) |u_hat|^2 / (2 N^2), j = 0, 1, 2; c[a] = (2 pi / L_a)^2.
template <typename T>
__global__ void __launch_bounds__(256) spectrum_moment_kernel(const Cx<T> *__restrict__ W, int m0, int m0p, int n0, int n1, int n2,
                                                              double scale, double c0, double c1, double c2, double *mom) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const long long total = (long long)m0 * n1 * n2;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(q % m0);
        const int iy = (int)((q / m0) % n1), iz = (int)(q / ((long long)m0 * n1));
        const int ky = iy < (n1 + 1) / 2 ? iy : iy - n1;  // fftfreq_int (spectral.jl:7)
        const int kz = n2 > 1 ? (iz < (n2 + 1) / 2 ? iz : iz - n2) : 0;
        const double k2 = c0 * (double)kx * (double)kx + c1 * (double)ky * (double)ky + c2 * (double)kz * (double)kz;
        const Cx<T> v = W[kx + (long long)iy * m0p + (long long)iz * m0p * n1];
        const double w = (kx > 0 && kx < n0 / 2) ? 2.0 : 1.0;
        const double e = w * scale * ((double)v.x * (double)v.x + (double)v.y * (double)v.y);
        s0 += e, s1 += k2 * e, s2 += k2 * k2 * e;
    }
    __shared__ double part[8][3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) part[warp][0] = s0, part[warp][1] = s1, part[warp][2] = s2;
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int wv = 0; wv < 8; wv++) t += part[wv][threadIdx.x];
        atomicAdd(mom + threadIdx.x, t);
    }
}

// u: ghosted velocity field (D components, stride cs); ehat: kmax + 1 device doubles (zeroed here), or -- with
// moment_c -- the three spectral moments of spectrum_moment_kernel
template <typename T>
static int spectrum_t(Poisson *ps, const T *u, long long cs, int kmax, double *ehat, Launch L, const double *moment_c = nullptr) {
    const int D = ps->D;
    long long rs[3] = {1, ps->Ng[0], (long long)ps->Ng[0] * ps->Ng[1]};
    FastGeom g;
    g.D = D;
    for (int d = 0; d < 3; d++) {
        g.n[d] = ps->n[d];
        g.ws[d] = ps->wstride[d];
        g.rs[d] = rs[d];
    }
    g.m0p = (int)ps->wstride[1];
    const T *e0 = (const T *)ps->eig[0];
    INSB_CUDA(cudaMemsetAsync(ehat, 0, (size_t)(moment_c ? 3 : kmax + 1) * sizeof(double), L.stream));
    double np = 1.0;
    for (int a = 0; a < D; a++) np *= (double)ps->n[a];
    for (int c = 0; c < D; c++) {
        T *R = const_cast<T *>(u) + c * cs;  // the R2C pass only reads R
        for (int a = 0; a < D; a++) R += ps->ioff[a] * rs[a];
        INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, g, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
        INSB_TRY(fast_pass<T>(0, ps->n[1], R, ps->W, g, 1, ps->tw[1], e0, e0, e0, T(1), 0, L));
        if (D == 3) INSB_TRY(fast_pass<T>(0, ps->n[2], R, ps->W, g, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
        if (moment_c)
            spectrum_moment_kernel<T><<<148 * 4, 256, 0, L.stream>>>((const Cx<T> *)ps->W, ps->m[0], g.m0p, ps->n[0], ps->n[1],
                                                                      D == 3 ? ps->n[2] : 1, 1.0 / (2.0 * np * np), moment_c[0],
                                                                      moment_c[1], D == 3 ? moment_c[2] : 0.0, ehat);
        else
            spectrum_bin_kernel<T><<<148 * 4, 256, (size_t)(kmax + 1) * sizeof(double), L.stream>>>(
                (const Cx<T> *)ps->W, ps->m[0], g.m0p, ps->n[0], ps->n[1], D == 3 ? ps->n[2] : 1, kmax, 1.0 / (2.0 * np * np), ehat);
        if (L.counter) ++*L.counter;
    }
    INSB_CUDA(cudaGetLastError());
    return 0;
}

int poisson_energy_spectrum(Poisson *ps, const void *u, long long cs, int kmax, double *ehat, Launch L) {
    if (!ps || !ps->fast || ps->comm) {
        set_error("energyspectrum needs a single-GPU psolver_spectral plan on a power-of-two periodic grid");
        return 1;
    }
    if (kmax < 0 || kmax > 4095) {
        set_error("energyspectrum: kmax out of range");
        return 1;
    }
    return ps->f64 ? spectrum_t<double>(ps, (const double *)u, cs, kmax, ehat, L)
                   : spectrum_t<float>(ps, (const float *)u, cs, kmax, ehat, L);
}

int poisson_spectral_moments(Poisson *ps, const void *u, long long cs, const double lengths[3], double *mom, Launch L) {
    if (!ps || !ps->fast || ps->comm) {
        set_error("turbulence_statistics needs a single-GPU psolver_spectral plan on a power-of-two periodic grid");
        return 1;
    }
    double c[3];
    for (int a = 0; a < 3; a++) {
        const double k = lengths[a] > 0.0 ? 2.0 * 3.14159265358979323846 / lengths[a] : 0.0;
        c[a] = k * k;
    }
    return ps->f64 ? spectrum_t<double>(ps, (const double *)u, cs, 0, mom, L, c) : spectrum_t<float>(ps, (const float *)u, cs, 0, mom, L, c);
}

// ----------------------------------------------------------------------------------------
// random_field (src/initializers.jl:109-183) on the device with the solver's own FFT passes:
//   white noise per component (the caller's, or Philox4x32-10 + Box-Muller here) -> forward transforms ->
//   projection onto the divergence-free space of the STAGGERED grid with the symbol of the backward difference,
//   m_a = (1 - exp(-i theta_a)) / h_a (:133-151) -> shell energies (spectral.jl:39-67, conjugate modes counted
//   twice) -> shell-wise rescaling to the target profile (:160-172) -> inverse transforms into view(u, Ip, a).
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0, c[1] = lo1, c[2] = n2, c[3] = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
}

// N(0, 1) into the interior of `ncomp` ghosted scalar fields: four values per counter, interior cell q of
// component c comes from counter (q / 4, c) -- independent of the launch shape
// (q4off: counter of the first cell of this rank's slab -- the field does not depend on the decomposition)
template <typename T>
__global__ void __launch_bounds__(256) rf_noise_kernel(T *u, long long cs, int ncomp, int n0, int n1, int n2, long long s1,
                                                       long long s2, unsigned long long seed, long long q4off) {
    const long long total4 = ((long long)n0 * n1 * n2 + 3) / 4;
    const long long q4 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (q4 >= total4 || c >= ncomp) return;
    const long long g4 = q4 + q4off;
    unsigned ctr[4] = {(unsigned)g4, (unsigned)(g4 >> 32), (unsigned)c, 0x1234567u};
    philox4x32_10(ctr, (unsigned)seed, (unsigned)(seed >> 32));
    double z[4];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const double u1 = ((double)ctr[2 * h] + 0.5) * 2.3283064365386963e-10;  // (0, 1)
        const double u2 = ((double)ctr[2 * h + 1] + 0.5) * 2.3283064365386963e-10;
        const double r = sqrt(-2.0 * log(u1));
        double sn, cs2;
        sincospi(2.0 * u2, &sn, &cs2);
        z[2 * h] = r * cs2, z[2 * h + 1] = r * sn;
    }
    const long long total = (long long)n0 * n1 * n2;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const long long q = 4 * q4 + e;
        if (q >= total) break;
        const int i = (int)(q % n0), j = (int)((q / n0) % n1), k = (int)(q / ((long long)n0 * n1));
        u[c * cs + i + j * s1 + k * s2] = (T)z[e];
    }
}

struct RfGeom {
    int D, m0, m0p, n0, n1, n2;
    double h[3];
    int kmax;
    double escale;  // 1 / (2 Ntot^2)
    // spectral array of this rank: [n2][n1l][m0p], row iy stands for ky index iy + ky0 (one GPU: n1l = n1, ky0 = 0;
    // decomposed: the ky-slab of this rank after the transpose)
    int n1l, ky0;
};

template <typename T>
__device__ __forceinline__ Cx<T> rf_symbol(int k, int n, double h) {
    double sn, cs;
    sincospi(2.0 * (double)k / (double)n, &sn, &cs);
    return Cx<T>{(T)((1.0 - cs) / h), (T)(sn / h)};  // (1 - exp(-i theta)) / h
}

__device__ __forceinline__ int rf_shell(long long k2) {
    int kap = (int)sqrt((double)k2);
    while ((long long)kap * kap > k2) kap--;
    while ((long long)(kap + 1) * (kap + 1) <= k2) kap++;
    return kap;
}

template <typename T>
__global__ void __launch_bounds__(256) rf_project_kernel(Cx<T> *U0, Cx<T> *U1, Cx<T> *U2, RfGeom g, double *eshell) {
    extern __shared__ double bins[];
    for (int i = threadIdx.x; i <= g.kmax; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
    const long long total = (long long)g.m0 * g.n1l * g.n2;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(q % g.m0);
        const int iyl = (int)((q / g.m0) % g.n1l), iz = (int)(q / ((long long)g.m0 * g.n1l));
        const int iy = iyl + g.ky0;
        const int ky = iy < (g.n1 + 1) / 2 ? iy : iy - g.n1;
        const int kz = g.D == 3 ? (iz < (g.n2 + 1) / 2 ? iz : iz - g.n2) : 0;
        const long long at = kx + (long long)iyl * g.m0p + (long long)iz * g.m0p * g.n1l;
        Cx<T> u[3] = {U0[at], U1[at], g.D == 3 ? U2[at] : Cx<T>{T(0), T(0)}};
        Cx<T> m[3] = {rf_symbol<T>(kx, g.n0, g.h[0]), rf_symbol<T>(ky, g.n1, g.h[1]),
                      g.D == 3 ? rf_symbol<T>(kz, g.n2, g.h[2]) : Cx<T>{T(0), T(0)}};
        Cx<T> mdotu = cmul(m[0], u[0]);
        mdotu = cadd(mdotu, cmul(m[1], u[1]));
        if (g.D == 3) mdotu = cadd(mdotu, cmul(m[2], u[2]));
        T m2 = (m[0].x * m[0].x + m[0].y * m[0].y) + (m[1].x * m[1].x + m[1].y * m[1].y);
        if (g.D == 3) m2 = m2 + (m[2].x * m[2].x + m[2].y * m[2].y);
        double e = 0.0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (a == 2 && g.D == 2) break;
            if (m2 > T(0)) {
                const Cx<T> t = cmul(cconj(m[a]), mdotu);
                u[a] = Cx<T>{u[a].x - t.x / m2, u[a].y - t.y / m2};
            }
            e += (double)u[a].x * (double)u[a].x + (double)u[a].y * (double)u[a].y;
        }
        U0[at] = u[0], U1[at] = u[1];
        if (g.D == 3) U2[at] = u[2];
        const int kap = rf_shell((long long)kx * kx + (long long)ky * ky + (long long)kz * kz);
        if (kap <= g.kmax) atomicAdd(&bins[kap], ((kx > 0 && kx < g.n0 / 2) ? 2.0 : 1.0) * g.escale * e);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= g.kmax; i += blockDim.x)
        if (bins[i] != 0.0) atomicAdd(eshell + i, bins[i]);
}

// factor[kappa] = sqrt(e0 / eshell) (0 for an empty shell), evaluated in the field type like :165-167, times the
// 1 / prod(Np) of the inverse transform
template <typename T>
__global__ void rf_factor_kernel(const double *e0, const double *eshell, int kmax, double invn, T *factor) {
    for (int i = threadIdx.x; i <= kmax; i += blockDim.x) {
        const T es = (T)eshell[i], et = (T)e0[i];
        const T f = es > T(0) ? (T)sqrt((double)(et / es)) : T(0);
        factor[i] = (T)((double)f * invn);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) rf_scale_kernel(Cx<T> *U0, Cx<T> *U1, Cx<T> *U2, RfGeom g, const T *__restrict__ factor) {
    const long long total = (long long)g.m0p * g.n1l * g.n2;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(q % g.m0p);
        const int iy = (int)((q / g.m0p) % g.n1l) + g.ky0, iz = (int)(q / ((long long)g.m0p * g.n1l));
        const int ky = iy < (g.n1 + 1) / 2 ? iy : iy - g.n1;
        const int kz = g.D == 3 ? (iz < (g.n2 + 1) / 2 ? iz : iz - g.n2) : 0;
        const int kap = rf_shell((long long)kx * kx + (long long)ky * ky + (long long)kz * kz);
        const T f = (kx < g.m0 && kap <= g.kmax) ? factor[kap] : T(0);
        U0[q] = Cx<T>{U0[q].x * f, U0[q].y * f}, U1[q] = Cx<T>{U1[q].x * f, U1[q].y * f};
        if (g.D == 3) U2[q] = Cx<T>{U2[q].x * f, U2[q].y * f};
    }
}

template <typename T>
static int random_field_t(Poisson *ps, T *u, long long cs, const double *e0_host, int nshell, int fill, unsigned long long seed,
                          const double xlims[3][2], Launch L) {
    const int D = ps->D;
    long long rs[3] = {1, ps->Ng[0], (long long)ps->Ng[0] * ps->Ng[1]};
    FastGeom g;
    g.D = D;
    for (int d = 0; d < 3; d++) {
        g.n[d] = ps->n[d];
        g.ws[d] = ps->wstride[d];
        g.rs[d] = rs[d];
    }
    g.m0p = (int)ps->wstride[1];
    RfGeom rg{};
    rg.D = D, rg.m0 = ps->m[0], rg.m0p = g.m0p, rg.n0 = ps->n[0], rg.n1 = ps->n[1], rg.n2 = D == 3 ? ps->n[2] : 1;
    double np = 1.0, k2max = 0.0;
    for (int a = 0; a < D; a++) {
        np *= (double)ps->n[a];
        rg.h[a] = (xlims[a][1] - xlims[a][0]) / ps->n[a];
        k2max += (double)(ps->n[a] / 2) * (ps->n[a] / 2);
    }
    rg.kmax = (int)std::floor(std::sqrt(k2max));  // kdiag: the last shell holds the largest wavenumber of the grid
    while ((double)(rg.kmax + 1) * (rg.kmax + 1) <= k2max) rg.kmax++;
    while ((double)rg.kmax * rg.kmax > k2max) rg.kmax--;
    if (nshell != rg.kmax + 1) {
        set_error("random_field: the profile table must hold kdiag + 1 = " + std::to_string(rg.kmax + 1) + " shells");
        return 1;
    }
    rg.escale = 1.0 / (2.0 * np * np);
    rg.n1l = ps->n[1], rg.ky0 = 0;
    const int world = comm_world(ps->comm), rank = comm_rank(ps->comm);
    const int nzl = ps->nzl, kyl = ps->kyl;
    if (ps->comm) rg.n1l = kyl, rg.ky0 = rank * kyl;
    // (decomposed: [n3][kyl][m0p] holds as many elements as the slab's [nzl][n1][m0p])
    const size_t wb = (size_t)g.m0p * ps->n[1] * (ps->comm ? nzl : rg.n2) * sizeof(Cx<T>);
    void *tmp[2] = {nullptr, nullptr};
    double *dsh = nullptr;
    T *dfac = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(L.stream);
        for (void *p : tmp) cudaFree(p);
        cudaFree(dsh);
        cudaFree(dfac);
    };
    if (cudaMalloc(&tmp[0], wb) != cudaSuccess || (D == 3 && cudaMalloc(&tmp[1], wb) != cudaSuccess) ||
        cudaMalloc((void **)&dsh, 2 * (size_t)nshell * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&dfac, (size_t)nshell * sizeof(T)) != cudaSuccess) {
        cleanup();
        set_error("random_field: cudaMalloc of the spectral work arrays failed");
        return 1;
    }
    void *tmp3 = nullptr;  // decomposed: W is the slab-side workspace of every transform, so three more arrays
    if (ps->comm && cudaMalloc(&tmp3, wb) != cudaSuccess) {
        cleanup();
        set_error("random_field: cudaMalloc of the spectral work arrays failed");
        return 1;
    }
    Cx<T> *U[3] = {ps->comm ? (Cx<T> *)tmp3 : (Cx<T> *)ps->W, (Cx<T> *)tmp[0], D == 3 ? (Cx<T> *)tmp[1] : (Cx<T> *)tmp[0]};
    T *R0 = u;
    for (int a = 0; a < D; a++) R0 += ps->ioff[a] * rs[a];
    const T *e0 = (const T *)ps->eig[0];
    int rc = 0;
    // ---- decomposed transforms (plain block layouts; the same transposes as the solve, poisson_solve) ----
    const long long m0p = g.m0p, blk = m0p * kyl * nzl;
    int sh = 0, shz = 0;
    while ((1 << sh) < kyl) sh++;
    while ((1 << shz) < nzl) shz++;
    FastGeom gx = g;
    gx.n[2] = nzl;
    // all ranks have passed this point of the stream (flag row `row`): buffers written by peers may be reused
    auto sync_ranks = [&](int row) -> int {
        if (ps->p2p) {
            INSB_TRY(comm_signal_peers(ps->peerFlags, row, rank, world, ps->epoch, L.stream));
            return comm_wait_flags(ps->flags, row, world, ps->epoch, ps->flags + 16, L.stream);
        }
        return 0;  // NCCL all-to-all calls are collective and stream ordered
    };
    auto forward_slab = [&](T *R, Cx<T> *dst) -> int {  // slab of one component -> dst = [n3][kyl][m0p]
        INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, gx, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
        FastGeom gy = g;
        gy.nother = nzl, gy.in = ps->W, gy.out = ps->S;
        gy.out_sd = m0p, gy.out_os = m0p * kyl, gy.out_sh = sh, gy.out_blk = blk;
        ps->epoch++;
        if (ps->p2p) {
            gy.use_outp = 1;
            for (int s2 = 0; s2 < world; s2++) gy.outp[s2] = (Cx<T> *)ps->peerRv[s2] + (long long)rank * blk;
        }
        INSB_TRY(fast_pass<T>(0, ps->n[1], R, ps->W, gy, 1, ps->tw[1], e0, e0, e0, T(1), 0, L));
        if (ps->p2p)
            INSB_TRY(sync_ranks(0));
        else
            INSB_TRY(comm_all_to_all(ps->comm, ps->S, ps->Rv, (size_t)blk * sizeof(Cx<T>), L.stream));
        FastGeom gz = g;
        gz.nother = kyl, gz.in = ps->Rv, gz.out = dst;
        gz.in_sd = gz.out_sd = m0p * kyl, gz.in_os = gz.out_os = m0p;
        INSB_TRY(fast_pass<T>(0, ps->n[2], R, ps->W, gz, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
        return sync_ranks(1);  // everybody has read its Rv: the next component may store into it
    };
    auto inverse_slab = [&](Cx<T> *src, T *R) -> int {  // src = [n3][kyl][m0p] -> slab of one component
        FastGeom gz = g;
        gz.nother = kyl, gz.in = src, gz.out = ps->Rv;
        gz.in_sd = gz.out_sd = m0p * kyl, gz.in_os = gz.out_os = m0p;
        ps->epoch++;
        if (ps->p2p) {  // z index k -> owner k / nzl, whose S holds [source rank][z local][ky local][kx]
            gz.use_outp = 1, gz.out_sh = shz, gz.out_sd = m0p * kyl, gz.out_os = m0p;
            for (int s2 = 0; s2 < world; s2++) gz.outp[s2] = (Cx<T> *)ps->peerS[s2] + (long long)rank * blk;
        }
        INSB_TRY(fast_pass<T>(1, ps->n[2], R, ps->W, gz, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
        if (ps->p2p)
            INSB_TRY(sync_ranks(1));
        else
            INSB_TRY(comm_all_to_all(ps->comm, ps->Rv, ps->S, (size_t)blk * sizeof(Cx<T>), L.stream));
        FastGeom gi = g;
        gi.nother = nzl, gi.in = ps->S, gi.out = ps->W;
        gi.in_sd = m0p, gi.in_os = m0p * kyl, gi.in_sh = sh, gi.in_blk = blk;
        INSB_TRY(fast_pass<T>(1, ps->n[1], R, ps->W, gi, 1, ps->tw[1], e0, e0, e0, T(1), 0, L));
        INSB_TRY(fast_pass<T>(4, ps->n[0], R, ps->W, gx, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
        return sync_ranks(0);  // everybody has read its S
    };
    auto run = [&]() -> int {
        if (fill) {
            const int n2l = ps->comm ? nzl : rg.n2;
            const long long total4 = ((long long)rg.n0 * rg.n1 * n2l + 3) / 4;
            dim3 gr((unsigned)((total4 + 255) / 256), D);
            rf_noise_kernel<T><<<gr, 256, 0, L.stream>>>(R0, cs, D, rg.n0, rg.n1, n2l, rs[1], rs[2], seed,
                                                         (long long)rank * nzl * rg.n0 * rg.n1 / 4);
            if (L.counter) ++*L.counter;
        }
        for (int c = 0; c < D; c++) {
            T *R = R0 + c * cs;
            if (ps->comm) {
                INSB_TRY(forward_slab(R, U[c]));
                continue;
            }
            INSB_TRY(fast_pass<T>(3, ps->n[0], R, U[c], g, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
            INSB_TRY(fast_pass<T>(0, ps->n[1], R, U[c], g, 1, ps->tw[1], e0, e0, e0, T(1), 0, L));
            if (D == 3) INSB_TRY(fast_pass<T>(0, ps->n[2], R, U[c], g, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
        }
        INSB_CUDA(cudaMemsetAsync(dsh, 0, (size_t)nshell * sizeof(double), L.stream));
        INSB_CUDA(cudaMemcpyAsync(dsh + nshell, e0_host, (size_t)nshell * sizeof(double), cudaMemcpyHostToDevice, L.stream));
        const size_t sm = (size_t)nshell * sizeof(double);
        rf_project_kernel<T><<<148 * 4, 256, sm, L.stream>>>(U[0], U[1], U[2], rg, dsh);
        if (ps->comm) INSB_TRY(comm_sum_doubles(ps->comm, dsh, nshell, L.stream));
        rf_factor_kernel<T><<<1, 256, 0, L.stream>>>(dsh + nshell, dsh, rg.kmax, 1.0 / np, dfac);
        rf_scale_kernel<T><<<148 * 8, 256, 0, L.stream>>>(U[0], U[1], U[2], rg, dfac);
        if (L.counter) *L.counter += 3;
        INSB_CUDA(cudaGetLastError());
        for (int c = 0; c < D; c++) {
            T *R = R0 + c * cs;
            if (ps->comm) {
                INSB_TRY(inverse_slab(U[c], R));
                continue;
            }
            if (D == 3) INSB_TRY(fast_pass<T>(1, ps->n[2], R, U[c], g, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
            INSB_TRY(fast_pass<T>(1, ps->n[1], R, U[c], g, 1, ps->tw[1], e0, e0, e0, T(1), 0, L));
            INSB_TRY(fast_pass<T>(4, ps->n[0], R, U[c], g, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
        }
        return 0;
    };
    rc = run();
    cleanup();
    cudaFree(tmp3);
    return rc;
}

int poisson_random_field(Poisson *ps, void *u, long long cs, const double *e0_host, int nshell, int fill,
                         unsigned long long seed, const double xlims[3][2], Launch L) {
    if (!ps || !ps->fast || (ps->comm && ps->D != 3)) {
        set_error("random_field needs a psolver_spectral plan on a power-of-two periodic grid");
        return 1;
    }
    if (nshell < 1 || nshell > 4096) {
        set_error("random_field: shell count out of range");
        return 1;
    }
    return ps->f64 ? random_field_t<double>(ps, (double *)u, cs, e0_host, nshell, fill, seed, xlims, L)
                   : random_field_t<float>(ps, (float *)u, cs, e0_host, nshell, fill, seed, xlims, L);
}

int poisson_ipc_export(Poisson *ps, void *blob) {
    if (!ps || !ps->comm) {
        set_error("ipc export: the plan is not slab-decomposed");
        return 1;
    }
    cudaIpcMemHandle_t h[3];
    INSB_CUDA(cudaIpcGetMemHandle(&h[0], ps->S));
    INSB_CUDA(cudaIpcGetMemHandle(&h[1], ps->Rv));
    INSB_CUDA(cudaIpcGetMemHandle(&h[2], ps->flags));
    static_assert(sizeof(h) == POISSON_IPC_BYTES, "blob size");
    std::memcpy(blob, h, sizeof(h));
    return 0;
}

int poisson_ipc_import(Poisson *ps, const void *blobs) {
    if (!ps || !ps->comm) {
        set_error("ipc import: the plan is not slab-decomposed");
        return 1;
    }
    const int rank = comm_rank(ps->comm), world = comm_world(ps->comm);
    if (world > 8) {
        set_error("ipc import: at most 8 ranks");
        return 1;
    }
    if ((ps->nzl & (ps->nzl - 1)) != 0) {
        set_error("ipc import: planes per rank must be a power of two");
        return 1;
    }
    for (int s = 0; s < world; s++) {
        if (s == rank) {
            ps->peerS[s] = ps->S;
            ps->peerRv[s] = ps->Rv;
            ps->peerFlags[s] = ps->flags;
            continue;
        }
        cudaIpcMemHandle_t h[3];
        std::memcpy(h, (const char *)blobs + (size_t)s * POISSON_IPC_BYTES, sizeof(h));
        INSB_CUDA(cudaIpcOpenMemHandle(&ps->peerS[s], h[0], cudaIpcMemLazyEnablePeerAccess));
        INSB_CUDA(cudaIpcOpenMemHandle(&ps->peerRv[s], h[1], cudaIpcMemLazyEnablePeerAccess));
        INSB_CUDA(cudaIpcOpenMemHandle((void **)&ps->peerFlags[s], h[2], cudaIpcMemLazyEnablePeerAccess));
    }
    ps->p2p = true;
    INSB_TRY(comm_enable_p2p_halo(ps->comm, (void *const *)ps->peerFlags, ps->halo_slot));
    return 0;
}

template <typename T>
static int solve_t(Poisson *ps, T *p, Launch L, const DivSource *div, const PoissonIO *io) {
    const int D = ps->D;
    long long rs[3] = {1, ps->Ng[0], (long long)ps->Ng[0] * ps->Ng[1]};
    T *R = p;
    for (int a = 0; a < D; a++) R += ps->ioff[a] * rs[a];
    T *Rout = R;
    if (io) {
        if (!poisson_supports_io(ps) || div) {
            set_error("poisson_solve: separate input / output arrays need the spectral fast path");
            return 1;
        }
        R = const_cast<T *>((const T *)io->rhs);  // the R2C pass only reads it
        Rout = (T *)io->sol;
        rs[1] = io->rs1;
        rs[2] = io->rs2;
    }
    if (div && !poisson_fuses_divergence(ps)) {
        set_error("poisson_solve: this solver cannot take the divergence from the velocity field");
        return 1;
    }
    // fused divergence! + scalewithvolume!: the x R2C pass computes its rows from the velocity
    auto with_div = [&](FastGeom &gx) {
        if (!div) return;
        const T *u = (const T *)div->u;
        for (int a = 0; a < D; a++) {
            u += ps->ioff[a] * rs[a];
            gx.div_r[a] = (const T *)div->rdlg[a] + ps->ioff[a];
            gx.div_d[a] = (const T *)div->dl[a] + ps->ioff[a];
        }
        gx.div_u = u;
        gx.div_cs = div->cs;
        gx.div_wrap = div->wrap;
    };
    if (ps->fast || (ps->fast_dct && ps->comm)) {
        FastGeom g;
        g.D = D;
        for (int d = 0; d < 3; d++) {
            g.n[d] = ps->n[d];
            g.ws[d] = ps->wstride[d];
            g.rs[d] = rs[d];
        }
        g.m0p = (int)ps->wstride[1];
        const T factor = (T)(ps->sign * ps->scale);
        const T *e0 = (const T *)ps->eig[0], *e1 = (const T *)ps->eig[1], *e2 = (const T *)ps->eig[2];
        const int tl = ps->tile;
        if (ps->comm) {
            // x, y on the z-slab; transpose to ky-slabs; z (forward, divide, inverse); back.
            const int world = comm_world(ps->comm), rank = comm_rank(ps->comm);
            const int kyl = ps->kyl, nzl = ps->nzl;
            const long long m0p = g.m0p;
            const long long blk = m0p * kyl * nzl;  // complex elements per peer block
            int sh = 0;
            while ((1 << sh) < kyl) sh++;
            // psolver_transform: DCT kernels (kinds 5-7) along the wall-bounded directions
            auto kind_of = [&](int base_kind, int dim) { return base_kind + (ps->per[dim] ? 0 : 5); };
            auto dct_tables = [&](FastGeom &gg, int dim) {
                gg.dw = ps->dw[dim];
                gg.dwi = ps->dwi[dim];
            };
            FastGeom gx = g;  // x passes see only the local planes
            gx.n[2] = nzl;
            {
                ProfScope sc(L, 5);
                FastGeom gd = gx;
                with_div(gd);
                INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, gd, 0, ps->tw[0], e0, e1, e2, factor, tl, L));
            }
            FastGeom gy = g;  // y forward: W (natural) -> S (blocked by destination rank)
            gy.nother = nzl;
            gy.in = ps->W;
            gy.out = ps->S;
            gy.out_sd = m0p;
            gy.out_os = m0p * kyl;
            gy.out_sh = sh;
            gy.out_blk = blk;
            // row-grouped block layouts (fft_fast.h): 256-byte runs for the NVLink stores
            const bool grp = kyl % 4 == 0 && nzl % 4 == 0 && getenv("INSB200_NO_GROUPED_TRANSPOSE") == nullptr;
            gy.out_grp = grp ? 1 : 0;  // blocks hold [z local][ky local (grouped)][kx]
            if (ps->p2p) {  // store every destination's block straight into its Rv
                ps->epoch++;
                gy.use_outp = 1;
                for (int s2 = 0; s2 < world; s2++) gy.outp[s2] = (Cx<T> *)ps->peerRv[s2] + (long long)rank * blk;
            }
            {
                ProfScope sc(L, 6);
                dct_tables(gy, 1);
                INSB_TRY(fast_pass<T>(kind_of(0, 1), ps->n[1], R, ps->W, gy, 1, ps->tw[1], e0, e1, e2, factor, tl, L));
            }
            {
                ProfScope sc(L, 7);
                if (ps->p2p) {
                    INSB_TRY(comm_signal_peers(ps->peerFlags, 0, rank, world, ps->epoch, L.stream));
                    INSB_TRY(comm_wait_flags(ps->flags, 0, world, ps->epoch, ps->flags + 16, L.stream));
                } else {
                    INSB_TRY(comm_all_to_all(ps->comm, ps->S, ps->Rv, (size_t)blk * sizeof(Cx<T>), L.stream));
                }
            }
            FastGeom gz = g;  // z: in place on Rv = [n3][kyl][m0p]
            gz.nother = kyl;
            gz.in = ps->Rv;
            gz.out = ps->Rv;
            gz.in_sd = gz.out_sd = m0p * kyl;
            gz.in_os = gz.out_os = m0p;
            gz.eig_off = rank * kyl;
            gz.in_grp = gz.out_grp = grp ? 2 : 0;
            if (ps->p2p) {  // z index k -> owner k / nzl, whose S holds [source rank][z local][ky local][kx]
                int shz = 0;
                while ((1 << shz) < nzl) shz++;
                gz.use_outp = 1;
                gz.out_sh = shz;
                if (grp) {  // ... or, grouped, [source rank][ky local][z local (grouped)][kx]
                    gz.out_grp = 1;
                    gz.out_os = m0p * nzl;
                }
                for (int s2 = 0; s2 < world; s2++) gz.outp[s2] = (Cx<T> *)ps->peerS[s2] + (long long)rank * blk;
            }
            {
                ProfScope sc(L, 8);
                dct_tables(gz, 2);
                INSB_TRY(fast_pass<T>(kind_of(2, 2), ps->n[2], R, ps->W, gz, 2, ps->tw[2], e0, e1, e2, factor, tl, L));
            }
            {
                ProfScope sc(L, 9);
                if (ps->p2p) {
                    INSB_TRY(comm_signal_peers(ps->peerFlags, 1, rank, world, ps->epoch, L.stream));
                    INSB_TRY(comm_wait_flags(ps->flags, 1, world, ps->epoch, ps->flags + 16, L.stream));
                } else {
                    INSB_TRY(comm_all_to_all(ps->comm, ps->Rv, ps->S, (size_t)blk * sizeof(Cx<T>), L.stream));
                }
            }
            FastGeom gi = g;  // y inverse: S (blocked by source rank) -> W (natural)
            gi.nother = nzl;
            gi.in = ps->S;
            gi.out = ps->W;
            gi.in_sd = m0p;
            gi.in_os = m0p * kyl;
            gi.in_sh = sh;
            gi.in_blk = blk;
            if (grp) {
                if (ps->p2p) {  // [source rank][ky local][z local (grouped)][kx]
                    gi.in_grp = 2;
                    gi.in_sd = m0p * nzl;
                } else {  // the all-to-all moved [z local][ky local (grouped)][kx] blocks verbatim
                    gi.in_grp = 1;
                }
            }
            {
                ProfScope sc(L, 10);
                dct_tables(gi, 1);
                INSB_TRY(fast_pass<T>(kind_of(1, 1), ps->n[1], R, ps->W, gi, 1, ps->tw[1], e0, e1, e2, factor, tl, L));
            }
            {
                ProfScope sc(L, 11);
                FastGeom go = gx;
                go.img = io ? (io->img & 1) : 0;
                INSB_TRY(fast_pass<T>(4, ps->n[0], Rout, ps->W, go, 0, ps->tw[0], e0, e1, e2, factor, tl, L));
            }
            return 0;
        }
        {
            ProfScope sc(L, 5);
            FastGeom gd = g;
            with_div(gd);
            INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, gd, 0, ps->tw[0], e0, e1, e2, factor, tl, L));
        }
        if (D == 3) {
            {
                ProfScope sc(L, 6);
                INSB_TRY(fast_pass<T>(0, ps->n[1], R, ps->W, g, 1, ps->tw[1], e0, e1, e2, factor, tl, L));
            }
            {
                ProfScope sc(L, 8);
                INSB_TRY(fast_pass<T>(2, ps->n[2], R, ps->W, g, 2, ps->tw[2], e0, e1, e2, factor, tl, L));
            }
            {
                ProfScope sc(L, 10);
                INSB_TRY(fast_pass<T>(1, ps->n[1], R, ps->W, g, 1, ps->tw[1], e0, e1, e2, factor, tl, L));
            }
        } else {
            ProfScope sc(L, 8);
            INSB_TRY(fast_pass<T>(2, ps->n[1], R, ps->W, g, 1, ps->tw[1], e0, e1, e2, factor, tl, L));
        }
        {
            ProfScope sc(L, 11);
            FastGeom go = g;
            go.img = io ? io->img : 0;
            INSB_TRY(fast_pass<T>(4, ps->n[0], Rout, ps->W, go, 0, ps->tw[0], e0, e1, e2, factor, tl, L));
        }
        return 0;
    }
    if (ps->fast_dct) {
        // x R2C -> y (FFT or DCT-II) -> z (forward, divide, inverse; 2-D: y is the middle pass)
        // -> y inverse -> x C2R; transforms along different directions commute, so this equals
        // the reference order (DCTs first, pressure.jl:262-300)
        FastGeom g;
        g.D = D;
        for (int d = 0; d < 3; d++) {
            g.n[d] = ps->n[d];
            g.ws[d] = ps->wstride[d];
            g.rs[d] = rs[d];
        }
        g.m0p = (int)ps->wstride[1];
        const T factor = (T)(ps->sign * ps->scale);
        const T *e0 = (const T *)ps->eig[0], *e1 = (const T *)ps->eig[1], *e2 = (const T *)ps->eig[2];
        auto pass = [&](int grp, int base_kind, int dim) -> int {
            FastGeom gd = g;
            gd.dw = ps->dw[dim];
            gd.dwi = ps->dwi[dim];
            const int kind = base_kind + (ps->per[dim] ? 0 : 5);
            ProfScope sc(L, grp);
            return fast_pass<T>(kind, ps->n[dim], R, ps->W, gd, dim, ps->tw[dim], e0, e1, e2, factor, 0, L);
        };
        {
            ProfScope sc(L, 5);
            FastGeom gd = g;
            with_div(gd);
            INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, gd, 0, ps->tw[0], e0, e1, e2, factor, 0, L));
        }
        if (D == 3) {
            INSB_TRY(pass(6, 0, 1));
            INSB_TRY(pass(8, 2, 2));
            INSB_TRY(pass(10, 1, 1));
        } else {
            INSB_TRY(pass(8, 2, 1));
        }
        {
            ProfScope sc(L, 11);
            INSB_TRY(fast_pass<T>(4, ps->n[0], R, ps->W, g, 0, ps->tw[0], e0, e1, e2, factor, 0, L));
        }
        return 0;
    }
    if (ps->fast_tri) {
        // x R2C (with the 1/Ω of pressure.jl:463-465 on load) -> FFT along the other periodic
        // direction -> Thomas along the wall-normal one -> inverse FFT -> x C2R
        FastGeom g;
        g.D = D;
        for (int d = 0; d < 3; d++) {
            g.n[d] = ps->n[d];
            g.ws[d] = ps->wstride[d];
            g.rs[d] = rs[d];
        }
        g.m0p = (int)ps->wstride[1];
        const int w = ps->wdir, o = w == 1 ? 2 : 1;  // wall-normal and the other periodic dimension
        const T *e0 = (const T *)ps->eig[0], *eo = (const T *)ps->eig[o];
        FastGeom gx = g;
        gx.pre = ps->invO;
        gx.pre_dim = w;
        with_div(gx);
        if (ps->comm) {
            if (!ps->p2p) {
                set_error("decomposed psolver_tridiagonal needs peer access (insb200_plan_ipc_import)");
                return 1;
            }
            const int world = comm_world(ps->comm), rank = comm_rank(ps->comm);
            const int nzl = ps->nzl, kxl = ps->kxl, n1 = ps->n[1], n3 = ps->n[2];
            const long long pitch = g.m0p;
            ps->epoch++;
            gx.n[2] = nzl;
            gx.x_kxl = kxl;  // x R2C: blocks of kxl modes go straight into their owners' Rv = [n3][n1][kxl]
            gx.x_zoff = rank * nzl;
            for (int s2 = 0; s2 < world; s2++) gx.outp[s2] = ps->peerRv[s2];
            {
                ProfScope sc(L, 5);
                INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, gx, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
            }
            {
                ProfScope sc(L, 7);
                INSB_TRY(comm_signal_peers(ps->peerFlags, 0, rank, world, ps->epoch, L.stream));
                INSB_TRY(comm_wait_flags(ps->flags, 0, world, ps->epoch, ps->flags + 16, L.stream));
            }
            FastGeom gz = g;  // z forward, in place on Rv
            gz.m0p = kxl;
            gz.nother = n1;
            gz.in = ps->Rv;
            gz.out = ps->Rv;
            gz.in_sd = gz.out_sd = (long long)n1 * kxl;
            gz.in_os = gz.out_os = kxl;
            {
                ProfScope sc(L, 6);
                INSB_TRY(fast_pass<T>(0, n3, R, ps->W, gz, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
            }
            {
                ProfScope sc(L, 8);
                dim3 b(128), gr((kxl + 127) / 128, n3);
                thomas_march_kernel<T><<<gr, b, 0, L.stream>>>((Cx<T> *)ps->Rv, (T *)ps->thomas_w, kxl, n1, kxl,
                                                               (long long)n1 * kxl, (const T *)ps->ta, (const T *)ps->tb,
                                                               (const T *)ps->tc, e0, eo, (T)ps->scale, rank * kxl,
                                                               ps->m[0]);
                if (L.counter) ++*L.counter;
                INSB_CUDA(cudaGetLastError());
            }
            FastGeom gi = gz;  // z inverse, plane k -> its owner's S = [nzl][n1][pitch] (natural slab layout)
            int shz = 0;
            while ((1 << shz) < nzl) shz++;
            gi.use_outp = 1;
            gi.out_sh = shz;
            gi.out_sd = (long long)n1 * pitch;
            gi.out_os = pitch;
            for (int s2 = 0; s2 < world; s2++) gi.outp[s2] = (Cx<T> *)ps->peerS[s2] + (long long)rank * kxl;
            {
                ProfScope sc(L, 10);
                INSB_TRY(fast_pass<T>(1, n3, R, ps->W, gi, 2, ps->tw[2], e0, e0, e0, T(1), 0, L));
            }
            {
                ProfScope sc(L, 9);
                INSB_TRY(comm_signal_peers(ps->peerFlags, 1, rank, world, ps->epoch, L.stream));
                INSB_TRY(comm_wait_flags(ps->flags, 1, world, ps->epoch, ps->flags + 16, L.stream));
            }
            FastGeom gc = g;
            gc.n[2] = nzl;
            {
                ProfScope sc(L, 11);
                INSB_TRY(fast_pass<T>(4, ps->n[0], R, ps->S, gc, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
            }
            return 0;
        }
        {
            ProfScope sc(L, 5);
            INSB_TRY(fast_pass<T>(3, ps->n[0], R, ps->W, gx, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
        }
        {
            ProfScope sc(L, 6);
            INSB_TRY(fast_pass<T>(0, ps->n[o], R, ps->W, g, o, ps->tw[o], e0, e0, e0, T(1), 0, L));
        }
        {
            ProfScope sc(L, 8);
            dim3 b(128), gr((ps->m[0] + 127) / 128, ps->n[o]);
            thomas_march_kernel<T><<<gr, b, 0, L.stream>>>((Cx<T> *)ps->W, (T *)ps->thomas_w, ps->m[0], ps->n[w],
                                                           ps->wstride[w], ps->wstride[o], (const T *)ps->ta,
                                                           (const T *)ps->tb, (const T *)ps->tc, e0, eo, (T)ps->scale,
                                                           0, ps->m[0]);
            if (L.counter) ++*L.counter;
            INSB_CUDA(cudaGetLastError());
        }
        {
            ProfScope sc(L, 10);
            INSB_TRY(fast_pass<T>(1, ps->n[o], R, ps->W, g, o, ps->tw[o], e0, e0, e0, T(1), 0, L));
        }
        {
            ProfScope sc(L, 11);
            INSB_TRY(fast_pass<T>(4, ps->n[0], R, ps->W, g, 0, ps->tw[0], e0, e0, e0, T(1), 0, L));
        }
        return 0;
    }
    static bool attr_set[2] = {false, false};
    for (const auto &pd : ps->passes) {
        PassArgs<T> a{};
        a.op = pd.op;
        a.dim = pd.dim;
        a.n = ps->n[pd.dim];
        a.D = D;
        a.src_real = pd.src_real;
        a.dst_real = pd.dst_real;
        a.half_out = pd.half_out;
        a.half_in = pd.half_in;
        a.W = (Cx<T> *)ps->W;
        a.R = R;
        for (int d = 0; d < 3; d++) {
            a.ws[d] = ps->wstride[d];
            a.rs[d] = rs[d];
            a.m[d] = ps->m[d];
            a.eig[d] = (const T *)ps->eig[d];
        }
        // while data is still real (first pass) / already real (last pass) the companion
        // extents are the real ones
        if (pd.src_real || pd.dst_real)
            for (int d = 0; d < 3; d++) a.m[d] = ps->n[d];
        if (pd.op == P_THOMAS && !pd.src_real) a.m[pd.dim] = ps->n[pd.dim];
        a.ld = a.n + 1;
        size_t per_line = (size_t)a.ld * sizeof(Cx<T>) * 2;
        int tile = 16;
        while (tile > 1 && tile * per_line > 200 * 1024) tile >>= 1;
        if (tile * per_line > (size_t)ps->max_smem) {
            set_error("transform length too large for the shared-memory line pass");
            return 1;
        }
        a.tile = tile;
        a.nrad = (int)ps->radix[pd.dim].size();
        for (int i = 0; i < a.nrad; i++) a.radix[i] = ps->radix[pd.dim][i];
        a.tw = (const Cx<T> *)ps->tw[pd.dim];
        a.dw = (const Cx<T> *)ps->dw[pd.dim];
        a.dwi = (const Cx<T> *)ps->dwi[pd.dim];
        a.sign = (T)ps->sign;
        a.scale = (T)ps->scale;
        a.prescale = nullptr;
        if (ps->kind == INSB200_PSOLVER_TRIDIAGONAL && pd.src_real) {
            a.prescale = (const T *)ps->invO;
            a.pdim = ps->wdir;
        }
        a.ta = (const T *)ps->ta;
        a.tb = (const T *)ps->tb;
        a.tc = (const T *)ps->tc;
        int da = pd.dim == 0 ? 1 : 0, db = pd.dim == 2 ? 1 : 2;
        long long nlines = (long long)a.m[da] * (D == 3 ? a.m[db] : 1);
        unsigned blocks = (unsigned)((nlines + tile - 1) / tile);
        size_t smem = tile * per_line;
        int ti = sizeof(T) == 8;
        if (!attr_set[ti]) {
            INSB_CUDA(cudaFuncSetAttribute(line_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ps->max_smem));
            attr_set[ti] = true;
        }
        line_pass_kernel<T><<<blocks, 256, smem, L.stream>>>(a);
        if (L.counter) ++*L.counter;
        INSB_CUDA(cudaGetLastError());
    }
    return 0;
}

// Opt-in (INSB200_FUSED_DIVERGENCE=1: per-thread loads, =2: shared-memory staged rows).  Measured
// on B200 at 512^3 Float32 the fused pass (0.65-0.72 ms, latency-bound behind the FFT's barriers)
// is no faster than divergence_kernel (0.37 ms, 5.7 TB/s) + the plain R2C pass (0.24 ms), so the
// separate kernels stay the default; profiles/r1d_ncu_full_summary.md.
bool poisson_fuses_divergence(const Poisson *ps) {
    const char *e = getenv("INSB200_FUSED_DIVERGENCE");
    return ps && (ps->fast || ps->fast_dct || ps->fast_tri) && e != nullptr && atoi(e) > 0;
}

bool poisson_supports_io(const Poisson *ps) { return ps && ps->fast && ps->D == 3; }

int poisson_solve(Poisson *ps, void *p, Launch L, const DivSource *div, const PoissonIO *io) {
    if (ps->cg) return cg_solve(ps->cg, p, L);
    return ps->f64 ? solve_t<double>(ps, (double *)p, L, div, io) : solve_t<float>(ps, (float *)p, L, div, io);
}

}  // namespace insb
