// les_math.cuh -- point functions of the eddy-viscosity closure (src/eddyviscosity.jl), shared by the
// four-pass kernels and the two fused kernels of les.cu.  Everything here is __host__ __device__ so
// that tests/host_les.cu can run the very same cell functions on the CPU against the oracle.
//
// The fused kernels never store the gradient tensor or the stress: a value of such a cell field at a
// neighbour J of cell I is recomputed from the velocity.  When J lies on a ghost plane the reference
// would read what zero_out_wall! (:344-353) followed by apply_bc_p! (boundary_conditions.jl:200-208)
// left there.  Both act plane by plane over the full extent of the other dimensions, dimensions in
// order, so the ghost entry equals -- dimension by dimension, independently --
//   periodic  : the entry of the opposite interior plane      (N-2 for plane 0, 1 for plane N-1)
//   symmetric : the entry of the adjacent interior plane      (1 for plane 0, N-2 for plane N-1)
//   Dirichlet : zero (zeroed by zero_out_wall!, apply_bc_p! is a no-op there, and a later copy along
//               another dimension copies a zero)
// LesEdge carries the BC kind per dimension and side.  PressureBC sides (two ghost layers) and
// decomposed directions are left to the four-pass path.
#pragma once
#include "internal.h"

namespace insb {

#if defined(__CUDA_ARCH__)
#define LES_LD(p) __ldg(p)
#else
#define LES_LD(p) (*(p))
#endif
#define LES_HD __host__ __device__ __forceinline__

struct LesEdge {
    int lo[3], hi[3];  // INSB200_BC_PERIODIC / _SYMMETRIC / _DIRICHLET per dimension and side
};

template <typename T>
LES_HD T les_eps();
template <>
LES_HD float les_eps<float>() { return 1.1920929e-07f; }
template <>
LES_HD double les_eps<double>() { return 2.220446049250313e-16; }

// eddy_viscosity_kernel! bodies (:265-340) on the collocated gradient M[a][b] = du_a/dx_b.
// MODEL: 1 Smagorinsky, 2 WALE, 3 Vreman, 4 QR
template <typename T, int MODEL>
LES_HD T eddy_visc_value(const T (&M)[3][3], T dx, T dy, T dz, T C) {
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
            // x^(3/2), x^(5/2), x^(5/4) by square roots (three pow calls cost more than the rest of the kernel)
            const T rs = sqrt(sdsd), rq = sqrt(ss);
            out = cd2 * (sdsd * rs) / (ss * ss * rq + sdsd * sqrt(rs) + les_eps<T>());
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
            out = cd2 * fabs(tr / T(3)) / (QS + les_eps<T>());
        }
    }
    return out;
}

// Field access relative to the current cell: U(c, dx, dy, dz) = component c at I + (dx, dy, dz), offsets in
// [-1, 1].  LesGLoad reads global memory (p = component 0 at the cell); the tile kernels read shared planes.
template <typename T>
struct LesGLoad {
    const T *p;
    long long s1, s2, cs;
    LES_HD T operator()(int c, int dx, int dy, int dz) const { return LES_LD(p + c * cs + dx + dy * s1 + dz * s2); }
};

// Metric tables by absolute index t of dimension a: straight from the Grid (read-only cache), or from the slices a
// tile kernel staged in shared memory (p[a] points at the entry of index 0 of table 0; tables LEN entries apart).
template <typename T>
struct LesGridTab {
    const Grid<T> &g;
    LES_HD T rdl(int a, int t) const { return LES_LD(g.rdl[a] + t); }
    LES_HD T rdu(int a, int t) const { return LES_LD(g.rdu[a] + t); }
    LES_HD T wa(int a, int t) const { return LES_LD(g.wa[a] + t); }
    LES_HD T wb(int a, int t) const { return LES_LD(g.wb[a] + t); }
};
template <typename T, int LEN>
struct LesSmemTab {
    const T *p[3];
    LES_HD T rdl(int a, int t) const { return p[a][0 * LEN + t]; }
    LES_HD T rdu(int a, int t) const { return p[a][1 * LEN + t]; }
    LES_HD T wa(int a, int t) const { return p[a][2 * LEN + t]; }
    LES_HD T wb(int a, int t) const { return p[a][3 * LEN + t]; }
};

// Where the values of a cell field at the neighbours I - e_d / I + e_d come from (EDGE: cell I touches a
// ghost plane).  Periodic: the velocity / viscosity ghost planes hold the periodic images, so the stencil is
// evaluated at the ghost position itself, only the metric tables are taken at the image (N-2 resp. 1: the
// ghost entries of the tables are not periodic images).  Symmetric: the cell itself.  Dirichlet: zero.
template <bool EDGE>
struct LesNb {
    int om[3], op[3];   // field offset of the source of I - e_d / I + e_d
    int tm[3], tp[3];   // metric-table index of that source
    bool zm[3], zp[3];  // that neighbour reads as zero
    template <typename T>
    LES_HD void init(const Grid<T> &g, const LesEdge &e, const int I[3]) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            om[d] = -1, op[d] = 1;
            tm[d] = I[d] - 1, tp[d] = I[d] + 1;
            zm[d] = zp[d] = false;
            if (EDGE) {
                if (I[d] == 1) {
                    if (e.lo[d] == INSB200_BC_PERIODIC) {
                        tm[d] = g.N[d] - 2;
                    } else {
                        om[d] = 0, tm[d] = 1;
                        zm[d] = e.lo[d] != INSB200_BC_SYMMETRIC;
                    }
                }
                if (I[d] == g.N[d] - 2) {
                    if (e.hi[d] == INSB200_BC_PERIODIC) {
                        tp[d] = 1;
                    } else {
                        op[d] = 0, tp[d] = I[d];
                        zp[d] = e.hi[d] != INSB200_BC_SYMMETRIC;
                    }
                }
            }
        }
    }
};

// collocate_tensor (:235-250) straight from the velocity: M[a][b] at the centre of cell I, the staggered
// gradients gradient_tensor_kernel! (:57-74) would have stored recomputed at I, I - e_lo, I - e_hi,
// I - e_lo - e_hi (the reference's summation order)
template <typename T, bool EDGE, typename LU, typename MET>
LES_HD void les_collocated_gradient(const Grid<T> &g, const MET &mt, const LesEdge &e, const LU &U, const int I[3], T (&M)[3][3]) {
    LesNb<EDGE> nb;
    nb.init(g, e, I);
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int b = 0; b < 3; b++) {
            if (a == b) {
                M[a][a] = (U(a, 0, 0, 0) - U(a, a == 0 ? -1 : 0, a == 1 ? -1 : 0, a == 2 ? -1 : 0)) * mt.rdl(a, I[a]);
            } else {
                const int lo = a < b ? a : b, hi = a < b ? b : a;
                T sum = T(0);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    int o[3] = {0, 0, 0};
                    int tb = I[b];
                    bool z = false;
                    if (c & 1) {
                        o[lo] = nb.om[lo], z = z || nb.zm[lo];
                        if (lo == b) tb = nb.tm[lo];
                    }
                    if (c & 2) {
                        o[hi] = nb.om[hi], z = z || nb.zm[hi];
                        if (hi == b) tb = nb.tm[hi];
                    }
                    int o2[3] = {o[0], o[1], o[2]};
                    o2[b] += 1;
                    T v = (U(a, o2[0], o2[1], o2[2]) - U(a, o[0], o[1], o[2])) * mt.rdu(b, tb);
                    if (EDGE && z) v = T(0);
                    sum = c == 0 ? v : sum + v;
                }
                M[a][b] = sum * T(0.25);
            }
        }
    }
}

// sigma_aa = -2 nu_t du_a/dx_a at the cell I + o (apply_eddy_viscosity_kernel!, :144-181); ta = index of the
// metric tables of dimension a there
template <typename T, typename LU, typename LV, typename MET>
LES_HD T les_sigma_diag(const MET &mt, const LU &U, const LV &V, int a, const int o[3], int ta) {
    int o1[3] = {o[0], o[1], o[2]};
    o1[a] -= 1;
    const T G = (U(a, o[0], o[1], o[2]) - U(a, o1[0], o1[1], o1[2])) * mt.rdl(a, ta);
    return T(-2) * V(0, o[0], o[1], o[2]) * G;
}

// sigma_ab = -2 nu_ab (du_a/dx_b + du_b/dx_a)/2 at the edge point of the cell I + o, a < b; nu_ab is the
// viscosity interpolated from the four surrounding centres (symmetrize! :81-86 + :144-181); ta, tb = indices
// of the metric tables of dimensions a and b there
template <typename T, typename LU, typename LV, typename MET>
LES_HD T les_sigma_off(const MET &mt, const LU &U, const LV &V, int a, int b, const int o[3], int ta, int tb) {
    int oa[3] = {o[0], o[1], o[2]}, ob[3] = {o[0], o[1], o[2]}, oab[3] = {o[0], o[1], o[2]};
    oa[a] += 1, ob[b] += 1, oab[a] += 1, oab[b] += 1;
    const T Gab = (U(a, ob[0], ob[1], ob[2]) - U(a, o[0], o[1], o[2])) * mt.rdu(b, tb);
    const T Gba = (U(b, oa[0], oa[1], oa[2]) - U(b, o[0], o[1], o[2])) * mt.rdu(a, ta);
    const T wa_a = mt.wa(a, ta), wb_a = mt.wb(a, ta);
    const T wa_b = mt.wa(b, tb), wb_b = mt.wb(b, tb);
    const T nu = wa_a * wa_b * V(0, o[0], o[1], o[2]) + wb_a * wa_b * V(0, oa[0], oa[1], oa[2]) +
                 wa_a * wb_b * V(0, ob[0], ob[1], ob[2]) + wb_a * wb_b * V(0, oab[0], oab[1], oab[2]);
    return T(-2) * nu * ((Gab + Gba) * T(0.5));
}

// divoftensor_kernel! (:200-219) with the stress recomputed on the fly: f[a] = sum_b d(sigma_ab)/dx_b at cell I
template <typename T, bool EDGE, typename LU, typename LV, typename MET>
LES_HD void les_stress_divergence(const Grid<T> &g, const MET &mt, const LesEdge &e, const LU &U, const LV &V, const int I[3],
                                  T (&f)[3]) {
    LesNb<EDGE> nb;
    nb.init(g, e, I);
    // diagonal stresses at I and I + e_a
    T dterm[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        int o[3] = {0, 0, 0};
        const T s0 = les_sigma_diag<T>(mt, U, V, a, o, I[a]);
        o[a] = nb.op[a];
        T s1 = les_sigma_diag<T>(mt, U, V, a, o, nb.tp[a]);
        if (EDGE && nb.zp[a]) s1 = T(0);
        dterm[a] = (s1 - s0) * mt.rdu(a, I[a]);
    }
    // off-diagonal stresses at I, I - e_a, I - e_b:  oterm[a][b] = d(sigma_ab)/dx_b
    T oterm[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = a + 1; b < 3; b++) {
            int o[3] = {0, 0, 0};
            const T s0 = les_sigma_off<T>(mt, U, V, a, b, o, I[a], I[b]);
            o[b] = nb.om[b];
            T sb = les_sigma_off<T>(mt, U, V, a, b, o, I[a], nb.tm[b]);
            if (EDGE && nb.zm[b]) sb = T(0);
            o[b] = 0;
            o[a] = nb.om[a];
            T sa = les_sigma_off<T>(mt, U, V, a, b, o, nb.tm[a], I[b]);
            if (EDGE && nb.zm[a]) sa = T(0);
            oterm[a][b] = (s0 - sb) * mt.rdl(b, I[b]);
            oterm[b][a] = (s0 - sa) * mt.rdl(a, I[a]);
        }
    f[0] = dterm[0] + oterm[0][1] + oterm[0][2];
    f[1] = oterm[1][0] + dterm[1] + oterm[1][2];
    f[2] = oterm[2][0] + oterm[2][1] + dterm[2];
}

}  // namespace insb
