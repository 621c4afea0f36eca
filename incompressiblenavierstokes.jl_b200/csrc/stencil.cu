// stencil.cu -- stage kernel (right-hand side + Runge-Kutta stage combine), divergence,
// pressure gradient / projection update, CFL and max-abs reductions.
//
// Replaces, per stage: fill!(k,0) + convectiondiffusion! (src/solver.jl:2-6,
// src/operators.jl:616-628) [+ applygravity!, convection_diffusion_temp!, dissipation!,
// src/solver.jl:29-34] + copyto!(state, statestart) + the `@. u += Δt*A[i,j]*k` broadcasts
// (src/time_steppers/step_explicit_runge_kutta.jl:24-32) with ONE pass over the fields.
#include <cstdlib>

#include "stencil_math.cuh"

namespace insb {

// ------------------------------------------------------------------------------------
// loaders
// ------------------------------------------------------------------------------------
template <typename T, int D>
struct GLoad {  // p points at component 0 of the current cell
    const T *p;
    long long s1, s2, cs;
    __device__ __forceinline__ T operator()(int c, int dx, int dy, int dz) const {
        return __ldg(p + c * cs + dx + dy * s1 + (D == 3 ? dz * s2 : 0));
    }
};

// ------------------------------------------------------------------------------------
// stage kernel, one thread per cell, neighbours through L1/L2 ("direct" variant)
// ------------------------------------------------------------------------------------
template <typename T, typename CT, int D, bool BOUSS, bool FULL>
__device__ __forceinline__ void stage_cell(const Grid<T> &g, const StageArgs<T> &a, const int I[3],
                                           long long idx, const T k[D], T kt) {
#pragma unroll
    for (int c = 0; c < D; c++) {
        long long q = idx + c * g.cs;
        T kc = k[c];
        if (a.k_out) a.k_out[q] = a.add_to_k ? a.k_out[q] + kc : kc;
        if (a.v) {
            T x = a.vbase[q];
            for (int j = 0; j < a.nprev; j++) x = axpy_rn<T, CT>(x, a.cprev[j], a.kprev[j][q]);
            a.v[q] = axpy_rn<T, CT>(x, a.cv, kc);
        }
        if (a.acc_out) a.acc_out[q] = axpy_rn<T, CT>(a.acc_in[q], a.ca, kc);
    }
    if (BOUSS) {
        if (a.tk_out) a.tk_out[idx] = a.add_to_k ? a.tk_out[idx] + kt : kt;
        if (a.tv) {
            T x = a.tvbase[idx];
            for (int j = 0; j < a.nprev; j++) x = axpy_rn<T, CT>(x, a.cprev[j], a.tkprev[j][idx]);
            a.tv[idx] = axpy_rn<T, CT>(x, a.cv, kt);
        }
        if (a.tacc_out) a.tacc_out[idx] = axpy_rn<T, CT>(a.tacc_in[idx], a.ca, kt);
    }
}

template <typename T, typename CT, int D, bool BOUSS, bool FULL>
__global__ void __launch_bounds__(256) stage_direct_kernel(Grid<T> g, StageArgs<T> a) {
    int I[3];
    I[0] = blockIdx.x * blockDim.x + threadIdx.x;
    I[1] = blockIdx.y * blockDim.y + threadIdx.y;
    I[2] = (D == 3) ? (int)blockIdx.z : 0;
    if (I[0] >= g.N[0] || I[1] >= g.N[1]) return;
    long long idx = I[0] + I[1] * g.s[1] + (D == 3 ? I[2] * g.s[2] : 0);
    bool interior = I[0] >= 1 && I[0] < g.N[0] - 1 && I[1] >= 1 && I[1] < g.N[1] - 1 &&
                    (D == 2 || (I[2] >= 1 && I[2] < g.N[2] - 1));
    if (!interior) {
        if (a.zero_outside && !a.add_to_k) {
            if (a.k_out)
                for (int c = 0; c < D; c++) a.k_out[idx + c * g.cs] = T(0);
            if (BOUSS && a.tk_out) a.tk_out[idx] = T(0);
        }
        return;
    }
    GLoad<T, D> U{a.u + idx, g.s[1], g.s[2], g.cs};
    T k[D];
    momentum_rhs<T, D, FULL>(U, g, I, (T)a.nu, (T)a.conv, k);
#pragma unroll
    for (int c = 0; c < D; c++) k[c] += (T)a.bodyforce[c];
    T kt = T(0);
    if (BOUSS) {
        GLoad<T, D> Tm{a.temp + idx, g.s[1], g.s[2], 0};
#pragma unroll
        for (int c = 0; c < D; c++)
            if (c == a.gdir) k[c] += (T)a.grav * avg_face<D>(Tm, g, I, c, 0);
        kt = temperature_rhs<T, D>(U, Tm, g, I, (T)a.kappa, (T)a.nu, a.dodiss != 0);
    }
    stage_cell<T, CT, D, BOUSS, FULL>(g, a, I, idx, k, kt);
}


// ------------------------------------------------------------------------------------
// stage kernel, 2.5-D marching variant (3-D, momentum only): a block owns a TX x TY column of
// cells and marches along z.  The planes z-1, z, z+1 of the three velocity components (with
// a one-cell xy halo) sit in a 4-slot shared-memory ring; plane z+2 is prefetched into
// registers while plane z is being computed and dropped into the free slot afterwards, so
// there is a single barrier per plane.  Every velocity value is read from HBM/L2 once per
// block column instead of ~45 times through L1 in the direct variant.
// ------------------------------------------------------------------------------------
template <typename T, int PX, int PLANE>
struct TileLoad {
    const T *pl[3];  // component 0 at this thread's cell in planes z-1, z, z+1
    __device__ __forceinline__ T operator()(int c, int dx, int dy, int dz) const {
        return pl[dz + 1][c * PLANE + dy * PX + dx];
    }
};

template <typename T, typename CT, int TX, int TY>
__global__ void __launch_bounds__(TX *TY) stage_tile_kernel(Grid<T> g, StageArgs<T> a, int zchunk) {
    constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY, NT = TX * TY, NE = 3 * PLANE;
    constexpr int NL = (NE + NT - 1) / NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s = reinterpret_cast<T *>(smem_raw);  // [4 slots][3 comps][PY][PX]
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = 1 + blockIdx.x * TX, y0 = 1 + blockIdx.y * TY;
    const int zbeg = 1 + blockIdx.z * zchunk;
    const int zend = min(zbeg + zchunk, g.N[2] - 1);
    if (zbeg >= zend) return;
    // this thread's share of a haloed plane: element e = (comp, py, px)
    int off2d[NL], comp[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        int e = tid + i * NT;
        int c = e / PLANE, r = e - c * PLANE;
        int py = r / PX, px = r - py * PX;
        int gx = min(x0 - 1 + px, g.N[0] - 1), gy = min(y0 - 1 + py, g.N[1] - 1);
        off2d[i] = gx + gy * (int)g.s[1];
        comp[i] = e < NE ? c : -1;
    }
    T reg[NL];
    auto load_plane = [&](int z) {
#pragma unroll
        for (int i = 0; i < NL; i++)
            if (comp[i] >= 0) reg[i] = __ldg(a.u + comp[i] * g.cs + z * g.s[2] + off2d[i]);
    };
    auto store_plane = [&](int z) {
        T *dst = s + (z & 3) * NE;
#pragma unroll
        for (int i = 0; i < NL; i++)
            if (comp[i] >= 0) dst[tid + i * NT] = reg[i];
    };
    for (int p = -1; p <= 1; p++) {
        load_plane(zbeg + p);
        store_plane(zbeg + p);
    }
    __syncthreads();
    int I[3] = {x0 + tx, y0 + ty, 0};
    const bool active = I[0] < g.N[0] - 1 && I[1] < g.N[1] - 1;
    const int lidx = (ty + 1) * PX + tx + 1;
    const long long idx2d = I[0] + I[1] * g.s[1];
    for (int z = zbeg; z < zend; z++) {
        const bool more = z + 1 < zend;
        if (more) load_plane(z + 2);
        if (active) {
            I[2] = z;
            const long long idx = idx2d + z * g.s[2];
            // pointwise operands first, so that their latency overlaps the stencil arithmetic
            T vb[3], ai[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                vb[c] = a.v ? a.vbase[idx + c * g.cs] : T(0);
                ai[c] = a.acc_out ? a.acc_in[idx + c * g.cs] : T(0);
            }
            TileLoad<T, PX, PLANE> U;
            U.pl[0] = s + ((z - 1) & 3) * NE + lidx;
            U.pl[1] = s + (z & 3) * NE + lidx;
            U.pl[2] = s + ((z + 1) & 3) * NE + lidx;
            T k[3];
            momentum_rhs<T, 3, true>(U, g, I, (T)a.nu, T(1), k);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                T kc = k[c] + (T)a.bodyforce[c];
                long long q = idx + c * g.cs;
                if (a.k_out) a.k_out[q] = kc;
                if (a.v) {
                    T x = vb[c];
                    for (int j = 0; j < a.nprev; j++) x = axpy_rn<T, CT>(x, a.cprev[j], a.kprev[j][q]);
                    a.v[q] = axpy_rn<T, CT>(x, a.cv, kc);
                }
                if (a.acc_out) a.acc_out[q] = axpy_rn<T, CT>(ai[c], a.ca, kc);
            }
        }
        if (more) store_plane(z + 2);
        __syncthreads();
    }
}

template <typename T, typename CT>
static int launch_stage_tile(const Grid<T> &g, const StageArgs<T> &a, Launch L) {
    constexpr int TX = 32, TY = 16;
    constexpr size_t smem = (size_t)4 * 3 * (TX + 2) * (TY + 2) * sizeof(T);
    static bool attr = false;
    if (!attr) {
        INSB_CUDA(cudaFuncSetAttribute(stage_tile_kernel<T, CT, TX, TY>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    int nz = g.N[2] - 2;
    int zchunk = 128;
    if (const char *e = getenv("INSB200_ZCHUNK")) zchunk = atoi(e) > 0 ? atoi(e) : zchunk;
    dim3 b(TX, TY, 1);
    dim3 gr((g.N[0] - 2 + TX - 1) / TX, (g.N[1] - 2 + TY - 1) / TY, (nz + zchunk - 1) / zchunk);
    stage_tile_kernel<T, CT, TX, TY><<<gr, b, smem, L.stream>>>(g, a, zchunk);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, typename CT, int D>
static int launch_stage_t(const Grid<T> &g, const StageArgs<T> &a, Launch L) {
    dim3 b(64, 4, 1);
    dim3 gr((g.N[0] + b.x - 1) / b.x, (g.N[1] + b.y - 1) / b.y, D == 3 ? g.N[2] : 1);
    bool full = a.conv && a.diff;
    StageArgs<T> aa = a;
    if (!a.diff) aa.nu = 0.0;
    if (a.bouss) {
        if (full)
            stage_direct_kernel<T, CT, D, true, true><<<gr, b, 0, L.stream>>>(g, aa);
        else
            stage_direct_kernel<T, CT, D, true, false><<<gr, b, 0, L.stream>>>(g, aa);
    } else {
        if (full)
            stage_direct_kernel<T, CT, D, false, true><<<gr, b, 0, L.stream>>>(g, aa);
        else
            stage_direct_kernel<T, CT, D, false, false><<<gr, b, 0, L.stream>>>(g, aa);
    }
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
struct WideOf {
    typedef double type;
};

template <typename T>
int launch_stage(const Grid<T> &g, const StageArgs<T> &a, bool ct_double, Launch L) {
    static const bool force_direct = getenv("INSB200_DIRECT_STAGE") != nullptr;
    if (g.D == 3 && !a.bouss && a.conv && a.diff && !a.zero_outside && !a.add_to_k && !force_direct) {
        if (sizeof(T) == 8 || ct_double) return launch_stage_tile<T, double>(g, a, L);
        return launch_stage_tile<T, T>(g, a, L);
    }
    if (sizeof(T) == 8 || ct_double) {
        return g.D == 3 ? launch_stage_t<T, double, 3>(g, a, L) : launch_stage_t<T, double, 2>(g, a, L);
    }
    return g.D == 3 ? launch_stage_t<T, T, 3>(g, a, L) : launch_stage_t<T, T, 2>(g, a, L);
}
template int launch_stage<float>(const Grid<float> &, const StageArgs<float> &, bool, Launch);
template int launch_stage<double>(const Grid<double> &, const StageArgs<double> &, bool, Launch);

// ------------------------------------------------------------------------------------
// divergence (+ volume scaling)
// ------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256) divergence_kernel(Grid<T> g, const T *u, T *div, int scale) {
    int I[3];
    I[0] = blockIdx.x * blockDim.x + threadIdx.x + 1;
    I[1] = blockIdx.y * blockDim.y + threadIdx.y + 1;
    I[2] = (D == 3) ? (int)blockIdx.z + 1 : 0;
    if (I[0] >= g.N[0] - 1 || I[1] >= g.N[1] - 1) return;
    long long idx = I[0] + I[1] * g.s[1] + (D == 3 ? I[2] * g.s[2] : 0);
    GLoad<T, D> U{u + idx, g.s[1], g.s[2], g.cs};
    div[idx] = divergence_at<T, D>(U, g, I, scale != 0);
}

template <typename T>
int launch_divergence(const Grid<T> &g, const T *u, T *div, bool scale, Launch L) {
    dim3 b(64, 4, 1);
    dim3 gr((g.N[0] - 2 + b.x - 1) / b.x, (g.N[1] - 2 + b.y - 1) / b.y, g.D == 3 ? g.N[2] - 2 : 1);
    if (g.D == 3)
        divergence_kernel<T, 3><<<gr, b, 0, L.stream>>>(g, u, div, scale);
    else
        divergence_kernel<T, 2><<<gr, b, 0, L.stream>>>(g, u, div, scale);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_divergence<float>(const Grid<float> &, const float *, float *, bool, Launch);
template int launch_divergence<double>(const Grid<double> &, const double *, double *, bool, Launch);

template <typename T>
__global__ void __launch_bounds__(256) scalewithvolume_kernel(Grid<T> g, T *p) {
    int ix = blockIdx.x * blockDim.x + threadIdx.x;
    int iy = blockIdx.y * blockDim.y + threadIdx.y;
    int iz = blockIdx.z;
    if (ix >= g.N[0] || iy >= g.N[1]) return;
    long long idx = ix + iy * g.s[1] + iz * g.s[2];
    T om = g.dl[0][ix] * g.dl[1][iy];
    if (g.D == 3) om = om * g.dl[2][iz];
    p[idx] *= om;
}

template <typename T>
int launch_scalewithvolume(const Grid<T> &g, T *p, Launch L) {
    dim3 b(64, 4, 1);
    dim3 gr((g.N[0] + b.x - 1) / b.x, (g.N[1] + b.y - 1) / b.y, g.D == 3 ? g.N[2] : 1);
    scalewithvolume_kernel<T><<<gr, b, 0, L.stream>>>(g, p);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_scalewithvolume<float>(const Grid<float> &, float *, Launch);
template int launch_scalewithvolume<double>(const Grid<double> &, double *, Launch);

// ------------------------------------------------------------------------------------
// pressure gradient: G = δp (pressuregradient!) or u += -δp (applypressure!)
// ------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256) pgrad_kernel(Grid<T> g, const T *p, T *G, T *usub) {
    int I[3];
    I[0] = blockIdx.x * blockDim.x + threadIdx.x + 1;
    I[1] = blockIdx.y * blockDim.y + threadIdx.y + 1;
    I[2] = (D == 3) ? (int)blockIdx.z + 1 : 0;
    if (I[0] >= g.N[0] - 1 || I[1] >= g.N[1] - 1) return;
    long long idx = I[0] + I[1] * g.s[1] + (D == 3 ? I[2] * g.s[2] : 0);
    T pc = __ldg(p + idx);
#pragma unroll
    for (int i = 0; i < D; i++) {
        T d = (__ldg(p + idx + g.s[i]) - pc) * __ldg(g.rdu[i] + I[i]);
        if (G) G[idx + i * g.cs] = d;
        if (usub) usub[idx + i * g.cs] += -d;
    }
}

template <typename T>
int launch_pressuregradient(const Grid<T> &g, const T *p, T *G, T *u_sub, Launch L) {
    dim3 b(64, 4, 1);
    dim3 gr((g.N[0] - 2 + b.x - 1) / b.x, (g.N[1] - 2 + b.y - 1) / b.y, g.D == 3 ? g.N[2] - 2 : 1);
    if (g.D == 3)
        pgrad_kernel<T, 3><<<gr, b, 0, L.stream>>>(g, p, G, u_sub);
    else
        pgrad_kernel<T, 2><<<gr, b, 0, L.stream>>>(g, p, G, u_sub);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_pressuregradient<float>(const Grid<float> &, const float *, float *, float *, Launch);
template int launch_pressuregradient<double>(const Grid<double> &, const double *, double *, double *,
                                             Launch);

// ------------------------------------------------------------------------------------
// reductions: convective CFL limit (src/solver.jl:171-186) and maximum(abs, u)
// scratch[0]: min as ordered bits of a non-negative double, scratch[1]: NaN flag / max bits
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename T>
struct IuRange {
    int lo[3][3], hi[3][3];  // [comp][dim], half-open
};

template <typename T, int D>
__global__ void __launch_bounds__(256) cfl_kernel(Grid<T> g, IuRange<T> r, const T *u,
                                                  unsigned long long *scratch) {
    long long ncell = g.cs;
    double m = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    int nanflag = 0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < ncell;
         q += (long long)gridDim.x * blockDim.x) {
        int I[3];
        I[0] = (int)(q % g.N[0]);
        long long r1 = q / g.N[0];
        I[1] = (int)(r1 % g.N[1]);
        I[2] = (int)(r1 / g.N[1]);
#pragma unroll
        for (int c = 0; c < D; c++) {
            bool in = true;
#pragma unroll
            for (int d = 0; d < D; d++) in = in && I[d] >= r.lo[c][d] && I[d] < r.hi[c][d];
            if (in) {
                T uv = u[q + c * g.cs];
                T v = g.du[c][I[c]] / fabs(uv);
                if (v != v) nanflag = 1;
                m = fmin(m, (double)v);
            }
        }
    }
    m = warp_min(m);
    nanflag = __any_sync(0xffffffffu, nanflag);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(scratch, (unsigned long long)__double_as_longlong(m));
        if (nanflag) atomicOr(scratch + 1, 1ULL);
    }
}

template <typename T>
int launch_cfl(const Grid<T> &g, const int Iu[3][3][2], const T *u, double *scratch, Launch L) {
    IuRange<T> r;
    for (int c = 0; c < 3; c++)
        for (int d = 0; d < 3; d++) {
            r.lo[c][d] = Iu[c][d][0];
            r.hi[c][d] = Iu[c][d][1];
        }
    unsigned long long init[2] = {0x7ff0000000000000ULL, 0ULL};
    INSB_CUDA(cudaMemcpyAsync(scratch, init, sizeof(init), cudaMemcpyHostToDevice, L.stream));
    int blocks = 148 * 8;
    if (g.D == 3)
        cfl_kernel<T, 3><<<blocks, 256, 0, L.stream>>>(g, r, u, (unsigned long long *)scratch);
    else
        cfl_kernel<T, 2><<<blocks, 256, 0, L.stream>>>(g, r, u, (unsigned long long *)scratch);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_cfl<float>(const Grid<float> &, const int[3][3][2], const float *, double *, Launch);
template int launch_cfl<double>(const Grid<double> &, const int[3][3][2], const double *, double *,
                                Launch);

template <typename T>
__global__ void __launch_bounds__(256) maxabs_kernel(const T *u, long long n,
                                                     unsigned long long *scratch) {
    double m = 0.0;
    int nanflag = 0;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n;
         q += (long long)gridDim.x * blockDim.x) {
        T v = fabs(u[q]);
        if (v != v) nanflag = 1;
        m = fmax(m, (double)v);
    }
    m = warp_max(m);
    nanflag = __any_sync(0xffffffffu, nanflag);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(scratch, (unsigned long long)__double_as_longlong(m));
        if (nanflag) atomicOr(scratch + 1, 1ULL);
    }
}

template <typename T>
int launch_maxabs(const T *u, long long n, double *scratch, Launch L) {
    INSB_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(double), L.stream));
    maxabs_kernel<T><<<148 * 8, 256, 0, L.stream>>>(u, n, (unsigned long long *)scratch);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_maxabs<float>(const float *, long long, double *, Launch);
template int launch_maxabs<double>(const double *, long long, double *, Launch);

}  // namespace insb
