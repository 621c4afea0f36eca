/* kernels_body.h -- included twice by ins_ref_kernels.c with REAL = float / double.
 *
 * Oracle (test infrastructure, never shipped): cell-by-cell restatement of the reference's
 * KernelAbstractions CPU kernels, one OpenMP thread team over the outer indices like the
 * reference's threaded CPU backend.  Index convention: 0-based, x fastest, component slowest
 * (the reference's column-major (N..., D) arrays).  Evaluation order follows the Julia source
 * (compiled with -ffp-contract=off: no FMA, like Julia's CPU code). */

#define IDX(i, j, k, c) ((size_t)(i) + (size_t)N0 * ((size_t)(j) + (size_t)N1 * ((size_t)(k) + (size_t)N2 * (size_t)(c))))

/* delta(u, i, j, J): guarded velocity derivative, src/operators.jl:179-188 */
static inline __attribute__((always_inline)) REAL FN(delta_u)(const REAL *u, int N0, int N1, int N2, const REAL *const *Delta,
                               const REAL *const *DeltaU, int i, int j, const int J[3]) {
    int Jm[3] = {J[0], J[1], J[2]}, Jp[3] = {J[0], J[1], J[2]};
    REAL d, du;
    if (i == j) {
        Jm[j] -= 1;
        d = Delta[j][J[j]];
        du = (u[IDX(J[0], J[1], J[2], i)] - u[IDX(Jm[0], Jm[1], Jm[2], i)]) / d;
    } else {
        Jp[j] += 1;
        d = DeltaU[j][J[j]];
        du = (u[IDX(Jp[0], Jp[1], Jp[2], i)] - u[IDX(J[0], J[1], J[2], i)]) / d;
    }
    return d > 2 * EPS ? du : (REAL)0;
}

/* convstress + diffstress at J, src/operators.jl:378-393, 556, 627-628 */
static inline __attribute__((always_inline)) REAL FN(stress)(const REAL *u, int N0, int N1, int N2, const REAL *const *Delta,
                              const REAL *const *DeltaU, const REAL *const *A1, const REAL *const *A2, REAL visc,
                              int conv, int diff, int i, int j, const int J[3]) {
    REAL s = 0;
    if (conv) {
        REAL uij, uji;
        if (i == j) {
            int Jm[3] = {J[0], J[1], J[2]};
            Jm[j] -= 1;
            uij = (u[IDX(Jm[0], Jm[1], Jm[2], i)] + u[IDX(J[0], J[1], J[2], i)]) / (REAL)2;
            uji = uij;
        } else {
            int Jj[3] = {J[0], J[1], J[2]}, Ji[3] = {J[0], J[1], J[2]};
            Jj[j] += 1;
            Ji[i] += 1;
            uij = (u[IDX(J[0], J[1], J[2], i)] + u[IDX(Jj[0], Jj[1], Jj[2], i)]) / (REAL)2;
            uji = A2[i][J[i]] * u[IDX(J[0], J[1], J[2], j)] + A1[i][J[i] + 1] * u[IDX(Ji[0], Ji[1], Ji[2], j)];
        }
        s = uij * uji;
    }
    if (diff) {
        REAL dstress = (-visc) * FN(delta_u)(u, N0, N1, N2, Delta, DeltaU, i, j, J);
        s = conv ? s + dstress : dstress;
    }
    return s;
}

/* one component c (compile-time constant after inlining) of
 * convectiondiffusion! / convection! / diffusion! (src/operators.jl:407-416, 545-554, 616-628):
 * f[I, c] += (T_c1 + T_c2) + T_c3 over I in 2:N-1 (tensordivergence :395-401) */
static inline __attribute__((always_inline)) void FN(convdiff_comp)(REAL *f, const REAL *u, const int *N, const int D,
                                                                    const REAL *const *Delta, const REAL *const *DeltaU,
                                                                    const REAL *const *A1, const REAL *const *A2,
                                                                    REAL visc, const int conv, const int diff,
                                                                    const int c) {
    const int N0 = N[0], N1 = N[1], N2 = D == 3 ? N[2] : 1;
    const int k0 = D == 3 ? 1 : 0, k1 = D == 3 ? N2 - 1 : 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = k0; k < k1; k++)
        for (int jj = 1; jj < N1 - 1; jj++)
            for (int ii = 1; ii < N0 - 1; ii++) {
                const int I[3] = {ii, jj, k};
                REAL t[3] = {0, 0, 0};
#pragma GCC unroll 3
                for (int j = 0; j < D; j++) {
                    REAL dd, sa, sb;
                    int Ia[3] = {ii, jj, k}, Ib[3] = {ii, jj, k};
                    if (c == j) {
                        Ib[j] += 1;
                        dd = DeltaU[j][I[j]];
                    } else {
                        Ia[j] -= 1;
                        dd = Delta[j][I[j]];
                    }
                    sa = FN(stress)(u, N0, N1, N2, Delta, DeltaU, A1, A2, visc, conv, diff, c, j, Ia);
                    sb = FN(stress)(u, N0, N1, N2, Delta, DeltaU, A1, A2, visc, conv, diff, c, j, Ib);
                    t[j] = -(sb - sa) / dd;
                }
                REAL s = t[0] + t[1];
                if (D == 3) s = s + t[2];
                f[IDX(ii, jj, k, c)] += s;
            }
}

#define INSREF_COMP(DD, CV, DF)                                                                      \
    do {                                                                                             \
        FN(convdiff_comp)(f, u, N, DD, Delta, DeltaU, A1, A2, visc, CV, DF, 0);                      \
        FN(convdiff_comp)(f, u, N, DD, Delta, DeltaU, A1, A2, visc, CV, DF, 1);                      \
        if (DD == 3) FN(convdiff_comp)(f, u, N, DD, Delta, DeltaU, A1, A2, visc, CV, DF, 2);         \
    } while (0)

void FN(convdiff_add)(REAL *f, const REAL *u, const int *N, int D, const REAL *const *Delta,
                      const REAL *const *DeltaU, const REAL *const *A1, const REAL *const *A2, REAL visc,
                      int conv, int diff) {
    /* specialise on (D, conv, diff) like the reference's Val-dispatched kernels */
    if (D == 3) {
        if (conv && diff) INSREF_COMP(3, 1, 1);
        else if (conv) INSREF_COMP(3, 1, 0);
        else INSREF_COMP(3, 0, 1);
    } else {
        if (conv && diff) INSREF_COMP(2, 1, 1);
        else if (conv) INSREF_COMP(2, 1, 0);
        else INSREF_COMP(2, 0, 1);
    }
}
#undef INSREF_COMP

/* divergence! (src/operators.jl:242-247) */
void FN(divergence)(REAL *div, const REAL *u, const int *N, int D, const REAL *const *Delta,
                    const REAL *const *DeltaU) {
    const int N0 = N[0], N1 = N[1], N2 = D == 3 ? N[2] : 1;
    const int k0 = D == 3 ? 1 : 0, k1 = D == 3 ? N2 - 1 : 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = k0; k < k1; k++)
        for (int jj = 1; jj < N1 - 1; jj++)
            for (int ii = 1; ii < N0 - 1; ii++) {
                const int I[3] = {ii, jj, k};
                REAL s = FN(delta_u)(u, N0, N1, N2, Delta, DeltaU, 0, 0, I) +
                         FN(delta_u)(u, N0, N1, N2, Delta, DeltaU, 1, 1, I);
                if (D == 3) s = s + FN(delta_u)(u, N0, N1, N2, Delta, DeltaU, 2, 2, I);
                div[IDX(ii, jj, k, 0)] = s;
            }
}

/* scalewithvolume! (src/operators.jl:218-231): the WHOLE array */
void FN(scalewithvolume)(REAL *p, const int *N, int D, const REAL *const *Delta) {
    const int N0 = N[0], N1 = N[1], N2 = D == 3 ? N[2] : 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < N2; k++)
        for (int jj = 0; jj < N1; jj++)
            for (int ii = 0; ii < N0; ii++) {
                REAL w = Delta[0][ii] * Delta[1][jj];
                if (D == 3) w = w * Delta[2][k];
                p[IDX(ii, jj, k, 0)] *= w;
            }
}

/* applypressure! (src/operators.jl:306-311, 192): u[I, i] += -((p[I+e_i] - p[I]) / Δu_i) */
void FN(applypressure)(REAL *u, const REAL *p, const int *N, int D, const REAL *const *DeltaU) {
    const int N0 = N[0], N1 = N[1], N2 = D == 3 ? N[2] : 1;
    const int k0 = D == 3 ? 1 : 0, k1 = D == 3 ? N2 - 1 : 1;
#pragma omp parallel for collapse(3) schedule(static)
    for (int c = 0; c < D; c++)
        for (int k = k0; k < k1; k++)
            for (int jj = 1; jj < N1 - 1; jj++)
                for (int ii = 1; ii < N0 - 1; ii++) {
                    int Ip[3] = {ii, jj, k};
                    Ip[c] += 1;
                    const int I[3] = {ii, jj, k};
                    REAL g = (p[IDX(Ip[0], Ip[1], Ip[2], 0)] - p[IDX(ii, jj, k, 0)]) / DeltaU[c][I[c]];
                    u[IDX(ii, jj, k, c)] += -g;
                }
}

#undef IDX
