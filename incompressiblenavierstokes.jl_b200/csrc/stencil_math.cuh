// stencil_math.cuh -- point functions of the staggered-grid operators, written once and
// shared by every kernel variant (global-memory loader, shared-memory tile loader).
//
// Arithmetic spec (reference, 1-based): src/operators.jl
//   δ            :179-188   guarded velocity derivative
//   avg          :197-200   linear interpolation of a scalar to a face
//   convstress   :378-393,  diffstress :556,  tensordivergence :395-401
//   convdiff_temp:675-681,  δ_coll/∇_coll :883-910, dissipation :772-781
//   applygravity :825-829
// Divisions by the 1-D metric tables are done as multiplications with reciprocal tables
// (Grid::rdl, rdu, rdlg, rdug) built once at plan creation; "/ 2" is "* 0.5" (exact).
#pragma once
#include "internal.h"

namespace insb {

// offset helper: value of component c at I + sa*e_da + sb*e_db
template <int D, typename LD>
__device__ __forceinline__ auto at2(const LD &U, int c, int da, int sa, int db, int sb)
    -> decltype(U(0, 0, 0, 0)) {
    int dx = (da == 0 ? sa : 0) + (db == 0 ? sb : 0);
    int dy = (da == 1 ? sa : 0) + (db == 1 ? sb : 0);
    int dz = (da == 2 ? sa : 0) + (db == 2 ? sb : 0);
    return U(c, dx, dy, dz);
}

// Metric providers: what momentum_rhs needs from the 1-D tables at cell I.
//   rdlg(i, s) = guarded 1/Δ_i[I_i + s]   (s = 0, 1)      rdug(j, s) = guarded 1/Δu_j[I_j + s] (s = -1, 0)
//   rdu(i) = 1/Δu_i[I_i]   rdl(j) = 1/Δ_j[I_j]   a2(i) = A2_i[I_i]   a1p(i) = A1_i[I_i + 1]
template <typename T>
struct GridMetrics {  // straight from the device tables (read-only cache)
    const Grid<T> &g;
    const int *I;
    __device__ __forceinline__ T rdlg(int i, int s) const { return __ldg(g.rdlg[i] + I[i] + s); }
    __device__ __forceinline__ T rdug(int j, int s) const { return __ldg(g.rdug[j] + I[j] + s); }
    __device__ __forceinline__ T rdu(int i) const { return __ldg(g.rdu[i] + I[i]); }
    __device__ __forceinline__ T rdl(int j) const { return __ldg(g.rdl[j] + I[j]); }
    __device__ __forceinline__ T a2(int i) const { return __ldg(g.a2[i] + I[i]); }
    __device__ __forceinline__ T a1p(int i) const { return __ldg(g.a1[i] + I[i] + 1); }
};

template <typename T>
struct RegMetrics {  // preloaded into registers (marching kernel: x/y entries are loop invariant)
    T v_rdlg[3][2], v_rdug[3][2], v_rdu[3], v_rdl[3], v_a2[3], v_a1p[3];
    __device__ __forceinline__ T rdlg(int i, int s) const { return v_rdlg[i][s]; }
    __device__ __forceinline__ T rdug(int j, int s) const { return v_rdug[j][s + 1]; }
    __device__ __forceinline__ T rdu(int i) const { return v_rdu[i]; }
    __device__ __forceinline__ T rdl(int j) const { return v_rdl[j]; }
    __device__ __forceinline__ T a2(int i) const { return v_a2[i]; }
    __device__ __forceinline__ T a1p(int i) const { return v_a1p[i]; }
};

// Momentum right-hand side k_i = (T_i1 + T_i2) + T_i3 at cell I.
// cfac scales the convective stress (1 or 0), nu the diffusive one (0 disables it).
template <typename T, int D, bool FULL, typename LD, typename MET>
__device__ __forceinline__ void momentum_rhs_m(const LD &U, const MET &M, T nu, T cfac, T k[D]) {
    const T half = T(0.5);
#pragma unroll
    for (int i = 0; i < D; i++) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < D; j++) {
            T t;
            if (i == j) {
                T um = at2<D>(U, i, i, -1, i, 0);
                T u0 = U(i, 0, 0, 0);
                T up = at2<D>(U, i, i, +1, i, 0);
                T m0 = (um + u0) * half;
                T m1 = (u0 + up) * half;
                T c0 = m0 * m0, c1 = m1 * m1;
                if (!FULL) {
                    c0 *= cfac;
                    c1 *= cfac;
                }
                T s0 = c0 - nu * ((u0 - um) * M.rdlg(i, 0));
                T s1 = c1 - nu * ((up - u0) * M.rdlg(i, 1));
                t = -(s1 - s0) * M.rdu(i);
            } else {
                T w2 = M.a2(i);
                T w1 = M.a1p(i);
                T ui_m = at2<D>(U, i, j, -1, j, 0);
                T ui_0 = U(i, 0, 0, 0);
                T ui_p = at2<D>(U, i, j, +1, j, 0);
                T uj_0 = U(j, 0, 0, 0);
                T uj_pi = at2<D>(U, j, i, +1, i, 0);
                T uj_m = at2<D>(U, j, j, -1, j, 0);
                T uj_mpi = at2<D>(U, j, j, -1, i, +1);
                T cb = ((ui_0 + ui_p) * half) * (w2 * uj_0 + w1 * uj_pi);
                T ca = ((ui_m + ui_0) * half) * (w2 * uj_m + w1 * uj_mpi);
                if (!FULL) {
                    cb *= cfac;
                    ca *= cfac;
                }
                T sb = cb - nu * ((ui_p - ui_0) * M.rdug(j, 0));
                T sa = ca - nu * ((ui_0 - ui_m) * M.rdug(j, -1));
                t = -(sb - sa) * M.rdl(j);
            }
            acc = (j == 0) ? t : acc + t;
        }
        k[i] = acc;
    }
}

template <typename T, int D, bool FULL, typename LD>
__device__ __forceinline__ void momentum_rhs(const LD &U, const Grid<T> &g, const int I[3], T nu,
                                             T cfac, T k[D]) {
    GridMetrics<T> M{g, I};
    momentum_rhs_m<T, D, FULL>(U, M, nu, cfac, k);
}

// Metric tables of the temperature / collocated-gradient functions at offset o (-1, 0) from I[i]:
// straight from global memory (read-only cache), or from the slices a tile kernel staged in shared memory.
template <typename T>
struct GridTab {
    const Grid<T> &g;
    const int *I;
    __device__ __forceinline__ T rdu(int i, int o) const { return __ldg(g.rdu[i] + I[i] + o); }
    __device__ __forceinline__ T rdl(int i, int o) const { return __ldg(g.rdl[i] + I[i] + o); }
    __device__ __forceinline__ T rdlg(int i, int o) const { return __ldg(g.rdlg[i] + I[i] + o); }
    __device__ __forceinline__ T rdug(int i, int o) const { return __ldg(g.rdug[i] + I[i] + o); }
    __device__ __forceinline__ T wa(int i, int o) const { return __ldg(g.wa[i] + I[i] + o); }
    __device__ __forceinline__ T wb(int i, int o) const { return __ldg(g.wb[i] + I[i] + o); }
};
template <typename T, int LEN>
struct SmemTab {  // p[i] points at the entry of I[i] in table 0 (rdu); tables LEN entries apart: rdu, rdl, rdlg, rdug, wa, wb
    const T *p[3];
    __device__ __forceinline__ T rdu(int i, int o) const { return p[i][0 * LEN + o]; }
    __device__ __forceinline__ T rdl(int i, int o) const { return p[i][1 * LEN + o]; }
    __device__ __forceinline__ T rdlg(int i, int o) const { return p[i][2 * LEN + o]; }
    __device__ __forceinline__ T rdug(int i, int o) const { return p[i][3 * LEN + o]; }
    __device__ __forceinline__ T wa(int i, int o) const { return p[i][4 * LEN + o]; }
    __device__ __forceinline__ T wb(int i, int o) const { return p[i][5 * LEN + o]; }
};

// avg(temp, Δ, I + s*e_i, i)
template <int D, typename T, typename LT, typename MET>
__device__ __forceinline__ T avg_face_m(const LT &Tm, const MET &M, int i, int s) {
    T t0 = at2<D>(Tm, 0, i, s, i, 0);
    T t1 = at2<D>(Tm, 0, i, s + 1, i, 0);
    return M.wa(i, s) * t0 + M.wb(i, s) * t1;
}
template <int D, typename T, typename LT>
__device__ __forceinline__ T avg_face(const LT &Tm, const Grid<T> &g, const int I[3], int i, int s) {
    return avg_face_m<D, T>(Tm, GridTab<T>{g, I}, i, s);
}

// guarded δ(u, i, j, I + off) for the collocated gradient
template <typename T, int D, typename LD, typename MET>
__device__ __forceinline__ T delta_g_m(const LD &U, const MET &M, int i, int j, int oi, int oj) {
    // offset = oi*e_i + oj*e_j (oi, oj in {0,-1}); i == j: oi is the only offset
    if (i == j) {
        T a = at2<D>(U, i, i, oi - 1, i, 0);
        T b = at2<D>(U, i, i, oi, i, 0);
        return (b - a) * M.rdlg(i, oi);
    }
    T a = at2<D>(U, i, i, oi, j, oj);
    T b = at2<D>(U, i, i, oi, j, oj + 1);
    return (b - a) * M.rdug(j, oj);
}
template <typename T, int D, typename LD>
__device__ __forceinline__ T delta_g(const LD &U, const Grid<T> &g, const int I[3], int i, int j, int oi, int oj) {
    return delta_g_m<T, D>(U, GridTab<T>{g, I}, i, j, oi, oj);
}

// Temperature right-hand side (convection_diffusion_temp! + optional dissipation!).
template <typename T, int D, typename LD, typename LT, typename MET>
__device__ __forceinline__ T temperature_rhs_m(const LD &U, const LT &Tm, const MET &M, T kappa, T nu, bool dodiss) {
    T acc = T(0);
    T tc = Tm(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < D; i++) {
        T tm = at2<D>(Tm, 0, i, -1, i, 0);
        T tp = at2<D>(Tm, 0, i, +1, i, 0);
        T dT1 = (tc - tm) * M.rdu(i, -1);
        T dT2 = (tp - tc) * M.rdu(i, 0);
        T uT1 = at2<D>(U, i, i, -1, i, 0) * avg_face_m<D, T>(Tm, M, i, -1);
        T uT2 = U(i, 0, 0, 0) * avg_face_m<D, T>(Tm, M, i, 0);
        T t = (-(uT2 - uT1) + kappa * (dT2 - dT1)) * M.rdl(i, 0);
        acc = (i == 0) ? t : acc + t;
    }
    if (dodiss) {
        T ss = T(0);
#pragma unroll
        for (int j = 0; j < D; j++) {
#pragma unroll
            for (int i = 0; i < D; i++) {
                T gij;
                if (i == j) {
                    gij = delta_g_m<T, D>(U, M, i, j, 0, 0);
                } else {
                    gij = (delta_g_m<T, D>(U, M, i, j, 0, 0) + delta_g_m<T, D>(U, M, i, j, -1, 0) +
                           delta_g_m<T, D>(U, M, i, j, 0, -1) + delta_g_m<T, D>(U, M, i, j, -1, -1)) *
                          T(0.25);
                }
                ss += gij * gij;
            }
        }
        acc += nu * ss;
    }
    return acc;
}
template <typename T, int D, typename LD, typename LT>
__device__ __forceinline__ T temperature_rhs(const LD &U, const LT &Tm, const Grid<T> &g, const int I[3], T kappa, T nu,
                                             bool dodiss) {
    return temperature_rhs_m<T, D>(U, Tm, GridTab<T>{g, I}, kappa, nu, dodiss);
}

// Divergence (divergence!, src/operators.jl:242-247) optionally scaled with the cell volume
// (scalewithvolume!, :218-231).
template <typename T, int D, typename LD>
__device__ __forceinline__ T divergence_at(const LD &U, const Grid<T> &g, const int I[3],
                                           bool scale) {
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < D; i++) {
        T d = (U(i, 0, 0, 0) - at2<D>(U, i, i, -1, i, 0)) * __ldg(g.rdlg[i] + I[i]);
        acc = (i == 0) ? d : acc + d;
    }
    if (scale) {
        T om = __ldg(g.dl[0] + I[0]) * __ldg(g.dl[1] + I[1]);
        if (D == 3) om = om * __ldg(g.dl[2] + I[2]);
        acc *= om;
    }
    return acc;
}

// rounding-exact helpers for the stage combine (no FMA contraction, like the reference
// broadcast `u + Δt*A[i,j]*k`, step_explicit_runge_kutta.jl:28-32)
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T, typename CT>
__device__ __forceinline__ T axpy_rn(T x, double c, T k) {
    return (T)add_rn((CT)x, mul_rn((CT)c, (CT)k));
}

}  // namespace insb
