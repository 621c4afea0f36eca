// cg.cu -- psolver_cg (src/pressure.jl:135-203): diagonally preconditioned conjugate gradients on
// the pressure Laplacian (laplacian!, src/operators.jl:320-359), the Poisson solver for every
// boundary-condition / grid combination the FFT-type solvers do not cover.
//
// Per iteration the reference runs ~10 broadcasts / kernels and three host-synchronising
// reductions.  Here an iteration is three passes over the fields --
//   q = z + beta q            (z = -r/d, the Jacobi preconditioner, recomputed on the fly)
//   Lq = laplacian(q), <q, Lq>  (ghost values by index logic: no apply_bc_p! pass)
//   p += alpha q, r -= alpha Lq, <r, r>, <z, r>
// -- with alpha, beta, rho kept in device memory (one-block kernels finish the deterministic
// two-stage reductions with a fixed-order shared-memory tree), and the host only reads the
// residual every few iterations to decide when to stop.
#include <cmath>

#include "internal.h"

namespace insb {

namespace {

constexpr int NPART = 148 * 4;
// device scalars
enum { S_RHO = 0, S_RHO_PREV = 1, S_QL = 2, S_RES2 = 3, S_ALPHA = 4, S_BETA = 5, S_N = 8 };

template <typename T>
struct CgDev {
    int D;
    int N[3], np[3], ioff[3];
    int bc[3][2];
    long long s[3];
    const T *dl[3];   // Δ   (ghosted index)
    const T *du[3];   // Δu
};

template <typename T>
__device__ __forceinline__ bool cell_of(const CgDev<T> &g, long long c, int I[3], long long &idx) {
    const long long total = (long long)g.np[0] * g.np[1] * g.np[2];
    if (c >= total) return false;
    I[0] = g.ioff[0] + (int)(c % g.np[0]);
    I[1] = g.ioff[1] + (int)((c / g.np[0]) % g.np[1]);
    I[2] = g.ioff[2] + (int)(c / ((long long)g.np[0] * g.np[1]));
    idx = I[0] + I[1] * g.s[1] + I[2] * g.s[2];
    return true;
}

// d of create_laplace_diag (pressure.jl:136-153); z = -r / d
template <typename T>
__device__ __forceinline__ T diag_of(const CgDev<T> &g, const int I[3]) {
    T om = g.dl[0][I[0]];
    for (int a = 1; a < g.D; a++) om = om * g.dl[a][I[a]];
    T d = T(0);
    for (int a = 0; a < g.D; a++) d -= om / g.dl[a][I[a]] * (T(1) / g.du[a][I[a]] + T(1) / g.du[a][I[a] - 1]);
    return d;
}

__device__ __forceinline__ void block_sum2(double a, double b, double *pa, double *pb) {
    __shared__ double sa[256], sb[256];
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sa[threadIdx.x] += sa[threadIdx.x + s];
            sb[threadIdx.x] += sb[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pa[blockIdx.x] = sa[0];
        pb[blockIdx.x] = sb[0];
    }
}

// r = rhs on Ip; partial <r, r> and <z, r>
template <typename T>
__global__ void __launch_bounds__(256) cg_init_kernel(CgDev<T> g, const T *__restrict__ p, T *__restrict__ r,
                                                      double *part_a, double *part_b) {
    double a = 0.0, b = 0.0;
    const long long total = (long long)g.np[0] * g.np[1] * g.np[2];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        int I[3];
        long long idx;
        cell_of(g, c, I, idx);
        const T rv = p[idx];
        r[idx] = rv;
        const T z = -rv / diag_of(g, I);
        a += (double)rv * (double)rv;
        b += (double)z * (double)rv;
    }
    block_sum2(a, b, part_a, part_b);
}

// q = z + beta q on Ip
template <typename T>
__global__ void __launch_bounds__(256) cg_direction_kernel(CgDev<T> g, const T *__restrict__ r, T *__restrict__ q,
                                                           const double *__restrict__ S) {
    const T beta = (T)S[S_BETA];
    const long long total = (long long)g.np[0] * g.np[1] * g.np[2];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        int I[3];
        long long idx;
        cell_of(g, c, I, idx);
        const T z = -r[idx] / diag_of(g, I);
        q[idx] = z + beta * q[idx];
    }
}

// Lq = laplacian!(q) on Ip (operators.jl:320-359) with the ghost values of apply_bc_p! taken by
// index: periodic -> opposite end, symmetric -> the cell itself, pressure -> 0; a Dirichlet side
// drops its one-sided term, which is the symmetric rule as well.  partial <q, Lq>
template <typename T>
__global__ void __launch_bounds__(256) cg_laplacian_kernel(CgDev<T> g, const T *__restrict__ q, T *__restrict__ Lq,
                                                           double *part_a, double *part_b) {
    double acc = 0.0;
    const long long total = (long long)g.np[0] * g.np[1] * g.np[2];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        int I[3];
        long long idx;
        cell_of(g, c, I, idx);
        const T qc = q[idx];
        T om = g.dl[0][I[0]];
        for (int a = 1; a < g.D; a++) om = om * g.dl[a][I[a]];
        T lap = T(0);
        for (int a = 0; a < g.D; a++) {
            const int ci = I[a] - g.ioff[a];
            T qm, qp;
            if (ci > 0)
                qm = q[idx - g.s[a]];
            else if (g.bc[a][0] == INSB200_BC_PERIODIC)
                qm = q[idx + (long long)(g.np[a] - 1) * g.s[a]];
            else if (g.bc[a][0] == INSB200_BC_PRESSURE)
                qm = T(0);
            else
                qm = qc;
            if (ci < g.np[a] - 1)
                qp = q[idx + g.s[a]];
            else if (g.bc[a][1] == INSB200_BC_PERIODIC)
                qp = q[idx - (long long)(g.np[a] - 1) * g.s[a]];
            else if (g.bc[a][1] == INSB200_BC_PRESSURE)
                qp = T(0);
            else
                qp = qc;
            lap += om / g.dl[a][I[a]] * ((qp - qc) / g.du[a][I[a]] - (qc - qm) / g.du[a][I[a] - 1]);
        }
        Lq[idx] = lap;
        acc += (double)qc * (double)lap;
    }
    block_sum2(acc, 0.0, part_a, part_b);
}

// p += alpha q, r -= alpha Lq on Ip; partial <r, r> and <z, r>
template <typename T>
__global__ void __launch_bounds__(256) cg_update_kernel(CgDev<T> g, T *__restrict__ p, T *__restrict__ r,
                                                        const T *__restrict__ q, const T *__restrict__ Lq,
                                                        const double *__restrict__ S, double *part_a, double *part_b) {
    const T alpha = (T)S[S_ALPHA];
    double a = 0.0, b = 0.0;
    const long long total = (long long)g.np[0] * g.np[1] * g.np[2];
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        int I[3];
        long long idx;
        cell_of(g, c, I, idx);
        p[idx] += alpha * q[idx];
        const T rv = r[idx] - alpha * Lq[idx];
        r[idx] = rv;
        const T z = -rv / diag_of(g, I);
        a += (double)rv * (double)rv;
        b += (double)z * (double)rv;
    }
    block_sum2(a, b, part_a, part_b);
}

// one block: add the block partials (fixed assignment + fixed tree order: deterministic) and derive
// the scalars.  (A single thread walking the 2 x 592 partials took 40 us per call -- more than the
// field passes of a small problem; profiles/r1e_ncu_secondary_summary.md.)
//   mode 0 (after init / update): res2 = sum a, rho_prev = rho, rho = sum b, beta = rho / rho_prev
//   mode 1 (after laplacian):     qL = sum a, alpha = rho / qL
__global__ void __launch_bounds__(256) cg_finish_kernel(const double *part_a, const double *part_b, int n, double *S,
                                                        int mode, int first) {
    __shared__ double sa[256], sb[256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += part_a[i];
        b += part_b[i];
    }
    sa[threadIdx.x] = a;
    sb[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sa[threadIdx.x] += sa[threadIdx.x + s];
            sb[threadIdx.x] += sb[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    a = sa[0];
    b = sb[0];
    if (mode == 0) {
        S[S_RES2] = a;
        const double prev = first ? 1.0 : S[S_RHO];  // rho_prev = one(T) before the first iteration
        S[S_RHO_PREV] = prev;
        S[S_RHO] = b;
        S[S_BETA] = prev != 0.0 ? b / prev : 0.0;
    } else {
        S[S_QL] = a;
        S[S_ALPHA] = a != 0.0 ? S[S_RHO] / a : 0.0;
    }
}

}  // namespace

struct CgSolver {
    bool f64 = false;
    int D = 0;
    int N[3] = {1, 1, 1}, np[3] = {1, 1, 1}, ioff[3] = {0, 0, 0};
    int bc[3][2] = {};
    void *dl[3] = {}, *du[3] = {};
    void *r = nullptr, *q = nullptr, *Lq = nullptr;
    double *S = nullptr, *part = nullptr;
    double *host = nullptr;  // pinned: residual read-back
    double reltol = 0, abstol = 0;
    bool stagnation_stop = false;  // tight default of default_psolver: stop when the residual no longer falls
    long long maxiter = 0;
    long long bytes = 0;
    int iterations = 0;
    double residual = 0;
};

void cg_destroy(CgSolver *c) {
    if (!c) return;
    for (int a = 0; a < 3; a++) {
        if (c->dl[a]) cudaFree(c->dl[a]);
        if (c->du[a]) cudaFree(c->du[a]);
    }
    for (void *p : {c->r, c->q, c->Lq, (void *)c->S, (void *)c->part})
        if (p) cudaFree(p);
    if (c->host) cudaFreeHost(c->host);
    delete c;
}

long long cg_bytes(const CgSolver *c) { return c ? c->bytes : 0; }
int cg_iterations(const CgSolver *c) { return c ? c->iterations : 0; }

int cg_create(const insb200_setup_t &s, const int Ip[3][2], double reltol, double abstol, long long maxiter,
              CgSolver **out) {
    *out = nullptr;
    CgSolver *c = new CgSolver();
    c->f64 = s.dtype == INSB200_F64;
    c->D = s.D;
    const size_t esz = c->f64 ? 8 : 4;
    long long ncell = 1, ndof = 1;
    for (int a = 0; a < s.D; a++) {
        c->N[a] = s.N[a];
        c->np[a] = Ip[a][1] - Ip[a][0];
        c->ioff[a] = Ip[a][0];
        c->bc[a][0] = s.bc_u[a][0];
        c->bc[a][1] = s.bc_u[a][1];
        ncell *= s.N[a];
        ndof *= c->np[a];
    }
    auto fail = [&](const char *m) {
        set_error(m);
        cg_destroy(c);
        return 1;
    };
    for (int a = 0; a < s.D; a++) {
        const size_t b = (size_t)s.N[a] * esz;
        if (cudaMalloc(&c->dl[a], b) != cudaSuccess || cudaMalloc(&c->du[a], b) != cudaSuccess) return fail("cudaMalloc failed (cg tables)");
        cudaMemcpy(c->dl[a], s.Delta[a], b, cudaMemcpyHostToDevice);
        cudaMemcpy(c->du[a], s.DeltaU[a], b, cudaMemcpyHostToDevice);
    }
    const size_t fb = (size_t)ncell * esz;
    if (cudaMalloc(&c->r, fb) != cudaSuccess || cudaMalloc(&c->q, fb) != cudaSuccess || cudaMalloc(&c->Lq, fb) != cudaSuccess)
        return fail("cudaMalloc failed (cg work fields)");
    cudaMemset(c->r, 0, fb);
    cudaMemset(c->q, 0, fb);
    cudaMemset(c->Lq, 0, fb);
    if (cudaMalloc((void **)&c->S, S_N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&c->part, 2 * NPART * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void **)&c->host, 2 * sizeof(double)) != cudaSuccess)
        return fail("cudaMalloc failed (cg scalars)");
    c->bytes = 3 * (long long)fb;
    // defaults of psolver_cg (pressure.jl:158-161)
    c->stagnation_stop = reltol < 0;
    if (reltol < 0) reltol = -reltol;
    c->reltol = reltol > 0 ? reltol : std::sqrt(c->f64 ? 2.220446049250313e-16 : 1.1920929e-07);
    c->abstol = abstol > 0 ? abstol : 0.0;
    c->maxiter = maxiter > 0 ? maxiter : ndof;
    *out = c;
    return 0;
}

template <typename T>
static int cg_solve_t(CgSolver *c, T *p, Launch L) {
    CgDev<T> g{};
    g.D = c->D;
    long long ncell = 1;
    for (int a = 0; a < 3; a++) {
        g.N[a] = c->N[a];
        g.np[a] = c->np[a];
        g.ioff[a] = c->ioff[a];
        g.bc[a][0] = c->bc[a][0];
        g.bc[a][1] = c->bc[a][1];
        g.dl[a] = (const T *)c->dl[a];
        g.du[a] = (const T *)c->du[a];
        ncell *= c->N[a];
    }
    g.s[0] = 1;
    g.s[1] = c->N[0];
    g.s[2] = (long long)c->N[0] * c->N[1];
    cudaStream_t st = L.stream;
    T *r = (T *)c->r, *q = (T *)c->q, *Lq = (T *)c->Lq;
    double *pa = c->part, *pb = c->part + NPART;
    auto count = [&](int n) {
        if (L.counter) *L.counter += n;
    };
    // q .= 0; r .= p; p .= 0 (pressure.jl:170-178)
    INSB_CUDA(cudaMemsetAsync(q, 0, (size_t)ncell * sizeof(T), st));
    cg_init_kernel<T><<<NPART, 256, 0, st>>>(g, p, r, pa, pb);
    INSB_CUDA(cudaMemsetAsync(p, 0, (size_t)ncell * sizeof(T), st));
    cg_finish_kernel<<<1, 256, 0, st>>>(pa, pb, NPART, c->S, 0, 1);
    count(2);
    INSB_CUDA(cudaMemcpyAsync(c->host, c->S + S_RES2, sizeof(double), cudaMemcpyDeviceToHost, st));
    INSB_CUDA(cudaStreamSynchronize(st));
    double residual = std::sqrt(c->host[0]);
    const double tolerance = std::fmax(c->reltol * residual, c->abstol);
    long long it = 0;
    constexpr int CHECK = 4;  // iterations between residual read-backs (the count reported is a multiple of it)
    double best = residual;
    long long best_it = 0;
    while (it < c->maxiter && residual > tolerance) {
        for (int k = 0; k < CHECK && it < c->maxiter; k++, it++) {
            cg_direction_kernel<T><<<NPART, 256, 0, st>>>(g, r, q, c->S);
            cg_laplacian_kernel<T><<<NPART, 256, 0, st>>>(g, q, Lq, pa, pb);
            cg_finish_kernel<<<1, 256, 0, st>>>(pa, pb, NPART, c->S, 1, 0);
            cg_update_kernel<T><<<NPART, 256, 0, st>>>(g, p, r, q, Lq, c->S, pa, pb);
            cg_finish_kernel<<<1, 256, 0, st>>>(pa, pb, NPART, c->S, 0, 0);
            count(5);
        }
        INSB_CUDA(cudaMemcpyAsync(c->host, c->S + S_RES2, sizeof(double), cudaMemcpyDeviceToHost, st));
        INSB_CUDA(cudaStreamSynchronize(st));
        residual = std::sqrt(c->host[0]);
        if (!(residual == residual)) {
            set_error("psolver_cg: the residual became NaN");
            return 1;
        }
        if (residual < 0.9 * best) best = residual, best_it = it;
        if (c->stagnation_stop && it - best_it >= 200) break;  // round-off floor reached above the tight tolerance
    }
    c->iterations = (int)it;
    c->residual = residual;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

int cg_solve(CgSolver *c, void *p, Launch L) {
    return c->f64 ? cg_solve_t<double>(c, (double *)p, L) : cg_solve_t<float>(c, (float *)p, L);
}

}  // namespace insb
