// les.cu -- eddy-viscosity closures: eddy_viscosity_closure! with the
// Smagorinsky / WALE / Vreman / QR models, 3-D (src/eddyviscosity.jl:57-74, 235-378).
//
// Same sequence as the reference -- gradient tensor, wall zeroing + ghost fills, viscosity at the
// cell centres, ghost fills, symmetrise + scale with the interpolated viscosity, ghost fills,
// divergence of the stress -- with symmetrize! folded into the scaling pass (the ghost entries it
// would touch are overwritten by the fills that follow) and the divergence pass applying the
// result directly to the stage outputs (k, v = v + cv f, acc = acc + ca f) instead of writing f.
// The nine tensor components are one allocation, component (a, b) = du_a/dx_b at index a + 3 b
// (the reference's tensorfield order xx, yx, zx, xy, yy, zy, xz, yz, zz).
#include "stencil_math.cuh"

namespace insb {

#define IX(a, b) ((a) + 3 * (b))

template <typename T>
__global__ void __launch_bounds__(256) les_gradient_kernel(Grid<T> g, const T *__restrict__ u, T *__restrict__ G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i >= g.N[0] - 1 || j >= g.N[1] - 1) return;
    const int I[3] = {i, j, k};
    const long long s[3] = {1, g.s[1], g.s[2]};
    const long long idx = i + j * g.s[1] + k * g.s[2];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const T *ua = u + a * g.cs + idx;
        const T c = __ldg(ua);
#pragma unroll
        for (int b = 0; b < 3; b++) {
            T v;
            if (a == b)
                v = (c - __ldg(ua - s[a])) / __ldg(g.dl[a] + I[a]);
            else
                v = (__ldg(ua + s[b]) - c) / __ldg(g.du[b] + I[b]);
            G[IX(a, b) * g.cs + idx] = v;
        }
    }
}

template <typename T>
__device__ __forceinline__ T eps_of();
template <>
__device__ __forceinline__ float eps_of<float>() { return 1.1920929e-07f; }
template <>
__device__ __forceinline__ double eps_of<double>() { return 2.220446049250313e-16; }

// MODEL: 1 Smagorinsky, 2 WALE, 3 Vreman, 4 QR
template <typename T, int MODEL>
__global__ void __launch_bounds__(256) les_viscosity_kernel(Grid<T> g, const T *__restrict__ G, T *__restrict__ visc, T C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i >= g.N[0] - 1 || j >= g.N[1] - 1) return;
    const long long s[3] = {1, g.s[1], g.s[2]};
    const long long idx = i + j * g.s[1] + k * g.s[2];
    // collocate_tensor (:235-250)
    T M[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const T *p = G + IX(a, b) * g.cs + idx;
            if (a == b) {
                M[a][b] = __ldg(p);
            } else {
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                M[a][b] = (__ldg(p) + __ldg(p - s[lo]) + __ldg(p - s[hi]) + __ldg(p - s[lo] - s[hi])) * T(0.25);
            }
        }
    const T dx = __ldg(g.dl[0] + i), dy = __ldg(g.dl[1] + j), dz = __ldg(g.dl[2] + k);
    T out;
    if (MODEL == 3) {  // Vreman (:307-340)
        const T d1 = dx * dx, d2 = dy * dy, d3 = dz * dz;
        const T a11 = M[0][0], a21 = M[0][1], a31 = M[0][2];
        const T a12 = M[1][0], a22 = M[1][1], a32 = M[1][2];
        const T a13 = M[2][0], a23 = M[2][1], a33 = M[2][2];
        const T b11 = d1 * a11 * a11 + d2 * a21 * a21 + d3 * a31 * a31;
        const T b12 = d1 * a11 * a12 + d2 * a21 * a22 + d3 * a31 * a32;
        const T b13 = d1 * a11 * a13 + d2 * a21 * a23 + d3 * a31 * a33;
        const T b22 = d1 * a12 * a12 + d2 * a22 * a22 + d3 * a32 * a32;
        const T b23 = d1 * a12 * a13 + d2 * a22 * a23 + d3 * a32 * a33;
        const T b33 = d1 * a13 * a13 + d2 * a23 * a23 + d3 * a33 * a33;
        const T abeta = a11 * a11 + a12 * a12 + a13 * a13 + a21 * a21 + a22 * a22 + a23 * a23 + a31 * a31 + a32 * a32 + a33 * a33;
        const T bbeta = b11 * b22 - b12 * b12 + b11 * b33 - b13 * b13 + b22 * b33 - b23 * b23;
        const T value = C * C * sqrt(bbeta / abeta);
        const T tol = sizeof(T) == 8 ? T(1e-12) : T(1e-6);
        out = abeta > tol ? value : T(0);
    } else {
        const T d = cbrt(dx * dy * dz);  // gridsize_vol (operators.jl:914-915)
        T S[3][3];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) S[a][b] = (M[a][b] + M[b][a]) * T(0.5);
        T ss = T(0);
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int a = 0; a < 3; a++) ss += S[a][b] * S[a][b];
        const T cd2 = (C * d) * (C * d);
        if (MODEL == 1) {  // Smagorinsky (:265-278)
            out = cd2 * sqrt(T(2) * ss);
        } else if (MODEL == 2) {  // WALE (:280-292)
            T G2[3][3];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) G2[a][b] = M[a][0] * M[0][b] + M[a][1] * M[1][b] + M[a][2] * M[2][b];
            const T tr3 = (G2[0][0] + G2[1][1] + G2[2][2]) / T(3);
            T sdsd = T(0);
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const T sd = (G2[a][b] + G2[b][a]) * T(0.5) - (a == b ? tr3 : T(0));
                    sdsd += sd * sd;
                }
            out = cd2 * pow(sdsd, T(1.5)) / (pow(ss, T(2.5)) + pow(sdsd, T(1.25)) + eps_of<T>());
        } else {  // QR (:294-305)
            T SS[3][3];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) SS[a][b] = S[a][0] * S[0][b] + S[a][1] * S[1][b] + S[a][2] * S[2][b];
            const T QS = (SS[0][0] + SS[1][1] + SS[2][2]) * T(0.5);
            T tr = T(0);
#pragma unroll
            for (int a = 0; a < 3; a++) tr += SS[a][0] * S[0][a] + SS[a][1] * S[1][a] + SS[a][2] * S[2][a];
            out = cd2 * fabs(tr / T(3)) / (QS + eps_of<T>());
        }
    }
    visc[idx] = out;
}

// symmetrize! + apply_eddy_viscosity_kernel! (:81-86, :144-181): sigma = -2 nu_t S in the slots
// xx, yy, zz, xy, xz, yz
template <typename T>
__global__ void __launch_bounds__(256) les_stress_kernel(Grid<T> g, const T *__restrict__ visc, T *G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i >= g.N[0] - 1 || j >= g.N[1] - 1) return;
    const long long s1 = g.s[1], s2 = g.s[2];
    const long long idx = i + j * s1 + k * s2;
    const T dxa = __ldg(g.dl[0] + i), dxb = __ldg(g.dl[0] + i + 1);
    const T dya = __ldg(g.dl[1] + j), dyb = __ldg(g.dl[1] + j + 1);
    const T dza = __ldg(g.dl[2] + k), dzb = __ldg(g.dl[2] + k + 1);
    const T ax = dxb / (dxa + dxb), bx = dxa / (dxa + dxb);
    const T ay = dyb / (dya + dyb), by = dya / (dya + dyb);
    const T az = dzb / (dza + dzb), bz = dza / (dza + dzb);
    const T *v = visc + idx;
    const T v0 = __ldg(v), vx = __ldg(v + 1), vy = __ldg(v + s1), vz = __ldg(v + s2);
    const T vxy = ax * ay * v0 + bx * ay * vx + ax * by * vy + bx * by * __ldg(v + 1 + s1);
    const T vxz = ax * az * v0 + bx * az * vx + ax * bz * vz + bx * bz * __ldg(v + 1 + s2);
    const T vyz = ay * az * v0 + by * az * vy + ay * bz * vz + by * bz * __ldg(v + s1 + s2);
    T *p = G + idx;
    const long long cs = g.cs;
    p[IX(0, 0) * cs] = T(-2) * v0 * p[IX(0, 0) * cs];
    p[IX(1, 1) * cs] = T(-2) * v0 * p[IX(1, 1) * cs];
    p[IX(2, 2) * cs] = T(-2) * v0 * p[IX(2, 2) * cs];
    p[IX(0, 1) * cs] = T(-2) * vxy * ((p[IX(0, 1) * cs] + p[IX(1, 0) * cs]) * T(0.5));
    p[IX(0, 2) * cs] = T(-2) * vxz * ((p[IX(0, 2) * cs] + p[IX(2, 0) * cs]) * T(0.5));
    p[IX(1, 2) * cs] = T(-2) * vyz * ((p[IX(1, 2) * cs] + p[IX(2, 1) * cs]) * T(0.5));
}

// divoftensor_kernel! (:200-219), applied to the stage outputs: with fl = -(div sigma),
//   k += fl,  v = v + cv fl,  acc = acc + ca fl     (any of the three may be absent)
template <typename T, typename CT>
__global__ void __launch_bounds__(256) les_divergence_kernel(Grid<T> g, const T *__restrict__ G, T *k_out, T *v, double cv,
                                                             T *acc, double ca) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1, k = blockIdx.z + 1;
    if (i >= g.N[0] - 1 || j >= g.N[1] - 1) return;
    const long long s1 = g.s[1], s2 = g.s[2], cs = g.cs;
    const long long idx = i + j * s1 + k * s2;
    const T dpx = __ldg(g.dl[0] + i), dpy = __ldg(g.dl[1] + j), dpz = __ldg(g.dl[2] + k);
    const T dux = __ldg(g.du[0] + i), duy = __ldg(g.du[1] + j), duz = __ldg(g.du[2] + k);
    auto S = [&](int a, int b, long long off) { return __ldg(G + IX(a, b) * cs + idx + off); };
    const T sxy = S(0, 1, 0), sxz = S(0, 2, 0), syz = S(1, 2, 0);
    T f[3];
    f[0] = (S(0, 0, 1) - S(0, 0, 0)) / dux + (sxy - S(0, 1, -s1)) / dpy + (sxz - S(0, 2, -s2)) / dpz;
    f[1] = (sxy - S(0, 1, -1)) / dpx + (S(1, 1, s1) - S(1, 1, 0)) / duy + (syz - S(1, 2, -s2)) / dpz;
    f[2] = (sxz - S(0, 2, -1)) / dpx + (syz - S(1, 2, -s1)) / dpy + (S(2, 2, s2) - S(2, 2, 0)) / duz;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const T fl = -f[c];
        const long long q = idx + c * cs;
        if (k_out) k_out[q] += fl;
        if (v) v[q] = axpy_rn<T, CT>(v[q], cv, fl);
        if (acc) acc[q] = axpy_rn<T, CT>(acc[q], ca, fl);
    }
}

template <typename T>
int launch_les_closure(const Grid<T> &g, const BcDesc &b, int model, double C, const T *u, T *G, T *visc, T *k_out, T *v,
                       double cv, T *acc, double ca, bool ct_double, Launch L) {
    if (g.D != 3) {
        set_error("eddy_viscosity_closure!: 3-D only (like the reference's model kernels)");
        return 1;
    }
    if (model < 1 || model > 4) {
        set_error("eddy viscosity model must be 1 (Smagorinsky), 2 (WALE), 3 (Vreman) or 4 (QR)");
        return 1;
    }
    dim3 blk(64, 4, 1);
    dim3 gr((g.N[0] - 2 + blk.x - 1) / blk.x, (g.N[1] - 2 + blk.y - 1) / blk.y, g.N[2] - 2);
    // fill!(visc, 0); fill!(g, 0) (:358-361): every interior entry is written below and every ghost
    // plane by zero_out_wall! / apply_bc_p!, except the second ghost layer of a PressureBC side
    bool pressure_bc = false;
    for (int a = 0; a < 3; a++) pressure_bc = pressure_bc || b.bc_u[a][0] == INSB200_BC_PRESSURE || b.bc_u[a][1] == INSB200_BC_PRESSURE;
    if (pressure_bc) {
        INSB_CUDA(cudaMemsetAsync(G, 0, (size_t)9 * g.cs * sizeof(T), L.stream));
        INSB_CUDA(cudaMemsetAsync(visc, 0, (size_t)g.cs * sizeof(T), L.stream));
    }
    les_gradient_kernel<T><<<gr, blk, 0, L.stream>>>(g, u, G);
    if (L.counter) ++*L.counter;
    INSB_TRY(launch_zero_out_wall<T>(g, b, G, 9, L));
    INSB_TRY(launch_bc_p_fields<T>(g, b, G, 9, L));
    switch (model) {
        case 1: les_viscosity_kernel<T, 1><<<gr, blk, 0, L.stream>>>(g, G, visc, (T)C); break;
        case 2: les_viscosity_kernel<T, 2><<<gr, blk, 0, L.stream>>>(g, G, visc, (T)C); break;
        case 3: les_viscosity_kernel<T, 3><<<gr, blk, 0, L.stream>>>(g, G, visc, (T)C); break;
        default: les_viscosity_kernel<T, 4><<<gr, blk, 0, L.stream>>>(g, G, visc, (T)C); break;
    }
    if (L.counter) ++*L.counter;
    INSB_TRY(launch_zero_out_wall<T>(g, b, visc, 1, L));
    INSB_TRY(launch_bc_p_fields<T>(g, b, visc, 1, L));
    les_stress_kernel<T><<<gr, blk, 0, L.stream>>>(g, visc, G);
    if (L.counter) ++*L.counter;
    INSB_TRY(launch_zero_out_wall<T>(g, b, G, 9, L));
    INSB_TRY(launch_bc_p_fields<T>(g, b, G, 9, L));
    if (sizeof(T) == 8 || ct_double)
        les_divergence_kernel<T, double><<<gr, blk, 0, L.stream>>>(g, G, k_out, v, cv, acc, ca);
    else
        les_divergence_kernel<T, T><<<gr, blk, 0, L.stream>>>(g, G, k_out, v, cv, acc, ca);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_les_closure<float>(const Grid<float> &, const BcDesc &, int, double, const float *, float *, float *,
                                       float *, float *, double, float *, double, bool, Launch);
template int launch_les_closure<double>(const Grid<double> &, const BcDesc &, int, double, const double *, double *,
                                        double *, double *, double *, double, double *, double, bool, Launch);

}  // namespace insb
