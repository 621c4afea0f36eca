// diag.cu -- on-device diagnostics feeding the processors: vorticity,
// interpolation of velocity / vorticity to the pressure points, Q criterion, kinetic-energy
// field and its volume integral.  One thread per cell, every output written once, inputs read
// through the read-only cache; all of them are a single pass over the fields (HBM-bound), so a
// plotted frame or a logged energy no longer needs a device -> host copy of the whole state.
//
// Reference (1-based): src/operators.jl
//   vorticity!         :853-881     interpolate_u_p!   :921-940    interpolate_ω_p! :950-984
//   qcrit!             :989-1000 (∇_coll :883-910)     kinetic_energy!  :1027-1055
//   total_kinetic_energy :1061-1066
#include "stencil_math.cuh"

namespace insb {

template <typename T, int D>
struct GLoadD {  // p points at component 0 of the current cell
    const T *p;
    long long s1, s2, cs;
    __device__ __forceinline__ T operator()(int c, int dx, int dy, int dz) const {
        return __ldg(p + c * cs + dx + dy * s1 + (D == 3 ? dz * s2 : 0));
    }
};

enum DiagOp { DG_VORT = 0, DG_INTERP_U = 1, DG_INTERP_W = 2, DG_QCRIT = 3, DG_KE_LAST = 4, DG_KE_FIRST = 5 };

struct DiagRange {
    int lo[3], hi[3];  // pressure DOFs Ip (0-based, half open)
};

// grid: x = cells along dim 0, y = dim 1, z = dim 2; every op masks its own index range
template <typename T, int D, int OP>
__global__ void __launch_bounds__(256) diag_kernel(Grid<T> g, DiagRange ip, const T *__restrict__ in, T *__restrict__ out) {
    int I[3];
    I[0] = blockIdx.x * blockDim.x + threadIdx.x;
    I[1] = blockIdx.y * blockDim.y + threadIdx.y;
    I[2] = D == 3 ? (int)blockIdx.z : 0;
    if (I[0] >= g.N[0] || I[1] >= g.N[1]) return;
    const long long idx = I[0] + I[1] * g.s[1] + (D == 3 ? I[2] * g.s[2] : 0);
    GLoadD<T, D> U{in + idx, g.s[1], g.s[2], g.cs};
    bool inp = true;  // inside Ip
#pragma unroll
    for (int a = 0; a < D; a++) inp = inp && I[a] >= ip.lo[a] && I[a] < ip.hi[a];
    if (OP == DG_VORT) {  // I in 1:N-1 (1-based) of every dimension
        bool ok = true;
#pragma unroll
        for (int a = 0; a < D; a++) ok = ok && I[a] < g.N[a] - 1;
        if (!ok) return;
        if (D == 2) {
            out[idx] = (U(1, 1, 0, 0) - U(1, 0, 0, 0)) * __ldg(g.rdu[0] + I[0]) -
                       (U(0, 0, 1, 0) - U(0, 0, 0, 0)) * __ldg(g.rdu[1] + I[1]);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int j = (i + 1) % 3, k = (i + 2) % 3;
                const T a = (at2<D>(U, k, j, 1, j, 0) - U(k, 0, 0, 0)) * __ldg(g.rdu[j] + I[j]);
                const T b = (at2<D>(U, j, k, 1, k, 0) - U(j, 0, 0, 0)) * __ldg(g.rdu[k] + I[k]);
                out[idx + i * g.cs] = a - b;
            }
        }
    } else if (OP == DG_INTERP_U) {  // component i: I_i in 2:N_i, the other indices unrestricted
#pragma unroll
        for (int i = 0; i < D; i++)
            if (I[i] >= 1) out[idx + i * g.cs] = (at2<D>(U, i, i, -1, i, 0) + U(i, 0, 0, 0)) * T(0.5);
    } else if (OP == DG_INTERP_W) {  // `in` is the vorticity; over Ip
        if (!inp) return;
        if (D == 2) {
            out[idx] = (U(0, -1, -1, 0) + U(0, 0, 0, 0)) * T(0.5);
        } else {
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const int j = (i + 1) % 3, k = (i + 2) % 3;
                out[idx + i * g.cs] = (at2<D>(U, i, j, -1, k, -1) + U(i, 0, 0, 0)) * T(0.5);
            }
        }
    } else if (OP == DG_QCRIT) {  // I = index + 1 over Np
        bool ok = true;
#pragma unroll
        for (int a = 0; a < D; a++) ok = ok && I[a] >= 1 && I[a] < 1 + (ip.hi[a] - ip.lo[a]);
        if (!ok) return;
        T G[D][D];
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
            for (int j = 0; j < D; j++) {
                if (i == j)
                    G[i][j] = delta_g<T, D>(U, g, I, i, j, 0, 0);
                else
                    G[i][j] = (delta_g<T, D>(U, g, I, i, j, 0, 0) + delta_g<T, D>(U, g, I, i, j, -1, 0) +
                               delta_g<T, D>(U, g, I, i, j, 0, -1) + delta_g<T, D>(U, g, I, i, j, -1, -1)) * T(0.25);
            }
        T tr = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) {
            T d = T(0);
#pragma unroll
            for (int k = 0; k < D; k++) d = k == 0 ? G[i][k] * G[k][i] : d + G[i][k] * G[k][i];
            tr = i == 0 ? d : tr + d;
        }
        out[idx] = -tr * T(0.5);
    } else {  // kinetic energy over Ip
        if (!inp) return;
        T k = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) {
            const T a = U(i, 0, 0, 0), b = at2<D>(U, i, i, -1, i, 0);
            if (OP == DG_KE_FIRST)
                k += (a + b) * (a + b);
            else
                k += a * a + b * b;
        }
        out[idx] = OP == DG_KE_FIRST ? k * T(0.125) : k * T(0.25);
    }
}

template <typename T, int D, int OP>
static int launch_diag_d(const Grid<T> &g, const DiagRange &ip, const T *in, T *out, Launch L) {
    dim3 b(64, 4, 1);
    dim3 gr((g.N[0] + b.x - 1) / b.x, (g.N[1] + b.y - 1) / b.y, D == 3 ? g.N[2] : 1);
    diag_kernel<T, D, OP><<<gr, b, 0, L.stream>>>(g, ip, in, out);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int launch_diag(const Grid<T> &g, const int Ip[3][2], int op, const T *in, T *out, Launch L) {
    DiagRange ip;
    for (int a = 0; a < 3; a++) ip.lo[a] = Ip[a][0], ip.hi[a] = Ip[a][1];
#define CASE(OP)                                                                   \
    case OP:                                                                       \
        return g.D == 3 ? launch_diag_d<T, 3, OP>(g, ip, in, out, L) : launch_diag_d<T, 2, OP>(g, ip, in, out, L);
    switch (op) {
        CASE(DG_VORT)
        CASE(DG_INTERP_U)
        CASE(DG_INTERP_W)
        CASE(DG_QCRIT)
        CASE(DG_KE_LAST)
        CASE(DG_KE_FIRST)
    }
#undef CASE
    set_error("unknown diagnostic");
    return 1;
}
template int launch_diag<float>(const Grid<float> &, const int[3][2], int, const float *, float *, Launch);
template int launch_diag<double>(const Grid<double> &, const int[3][2], int, const double *, double *, Launch);

// total_kinetic_energy: sum over Ip of ke * (Δx Δy Δz) without materialising ke.  Deterministic:
// fixed assignment of cells to threads, block partials in double, one block adds the partials
// with a fixed-order shared-memory tree (same result for the same grid on every run).
template <typename T, int D, bool FIRST>
__global__ void __launch_bounds__(256) total_ke_kernel(Grid<T> g, DiagRange ip, const T *__restrict__ u, double *partial) {
    const int nx = ip.hi[0] - ip.lo[0], ny = ip.hi[1] - ip.lo[1], nz = D == 3 ? ip.hi[2] - ip.lo[2] : 1;
    const long long total = (long long)nx * ny * nz;
    double acc = 0.0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        int I[3];
        I[0] = ip.lo[0] + (int)(q % nx);
        I[1] = ip.lo[1] + (int)((q / nx) % ny);
        I[2] = D == 3 ? ip.lo[2] + (int)(q / ((long long)nx * ny)) : 0;
        const long long idx = I[0] + I[1] * g.s[1] + (D == 3 ? I[2] * g.s[2] : 0);
        GLoadD<T, D> U{u + idx, g.s[1], g.s[2], g.cs};
        T k = T(0);
#pragma unroll
        for (int i = 0; i < D; i++) {
            const T a = U(i, 0, 0, 0), b = at2<D>(U, i, i, -1, i, 0);
            k += FIRST ? (a + b) * (a + b) : a * a + b * b;
        }
        k = FIRST ? k * T(0.125) : k * T(0.25);
        T om = __ldg(g.dl[0] + I[0]) * __ldg(g.dl[1] + I[1]);
        if (D == 3) om = om * __ldg(g.dl[2] + I[2]);
        acc += (double)(k * om);
    }
    __shared__ double sm[256];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double *partial, int n, double *out) {
    __shared__ double sm[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];  // fixed assignment, fixed tree: deterministic
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sm[threadIdx.x] += sm[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

template <typename T>
int launch_total_ke(const Grid<T> &g, const int Ip[3][2], const T *u, bool first, double *partial, int npartial,
                    double *out, Launch L) {
    DiagRange ip;
    for (int a = 0; a < 3; a++) ip.lo[a] = Ip[a][0], ip.hi[a] = Ip[a][1];
    if (g.D == 3) {
        if (first) total_ke_kernel<T, 3, true><<<npartial, 256, 0, L.stream>>>(g, ip, u, partial);
        else total_ke_kernel<T, 3, false><<<npartial, 256, 0, L.stream>>>(g, ip, u, partial);
    } else {
        if (first) total_ke_kernel<T, 2, true><<<npartial, 256, 0, L.stream>>>(g, ip, u, partial);
        else total_ke_kernel<T, 2, false><<<npartial, 256, 0, L.stream>>>(g, ip, u, partial);
    }
    sum_partials_kernel<<<1, 256, 0, L.stream>>>(partial, npartial, out);
    if (L.counter) *L.counter += 2;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_total_ke<float>(const Grid<float> &, const int[3][2], const float *, bool, double *, int, double *, Launch);
template int launch_total_ke<double>(const Grid<double> &, const int[3][2], const double *, bool, double *, int, double *, Launch);

}  // namespace insb
