// host_les.cu -- test infrastructure: runs the cell functions of csrc/les_math.cuh (the ones the fused
// eddy-viscosity kernels are made of) on the CPU, cell by cell, so that tests/test_host_les.py can compare
// them with oracle/eddyviscosity.py without a GPU.  Not part of the product library.
#include "../../incompressiblenavierstokes.jl_b200/csrc/les_math.cuh"

using namespace insb;

template <typename T>
static Grid<T> make_grid(const int N[3], const void *const tab[3][5]) {
    Grid<T> g{};
    g.D = 3;
    for (int a = 0; a < 3; a++) g.N[a] = N[a];
    g.s[0] = 1, g.s[1] = N[0], g.s[2] = (long long)N[0] * N[1];
    g.cs = (long long)N[0] * N[1] * N[2];
    for (int a = 0; a < 3; a++) {
        g.dl[a] = (const T *)tab[a][0];
        g.rdl[a] = (const T *)tab[a][1];
        g.rdu[a] = (const T *)tab[a][2];
        g.wa[a] = (const T *)tab[a][3];
        g.wb[a] = (const T *)tab[a][4];
    }
    return g;
}

template <typename T>
static void run(int phase, bool edge_all, const int N[3], const void *const tab[3][5], const int lo[3], const int hi[3], int model, double C,
                const T *u, T *visc, T *f) {
    Grid<T> g = make_grid<T>(N, tab);
    LesEdge e;
    for (int d = 0; d < 3; d++) e.lo[d] = lo[d], e.hi[d] = hi[d];
    for (int k = 1; k < N[2] - 1; k++)
        for (int j = 1; j < N[1] - 1; j++)
            for (int i = 1; i < N[0] - 1; i++) {
                const int I[3] = {i, j, k};
                const long long idx = i + j * g.s[1] + k * g.s[2];
                const LesGLoad<T> U{u + idx, g.s[1], g.s[2], g.cs}, V{visc + idx, g.s[1], g.s[2], 0};
                // like the kernels: the ghost rules only where a cell touches a ghost plane (edge_all: everywhere,
                // which must give the same result)
                const bool edge = edge_all || i == 1 || j == 1 || k == 1 || i == N[0] - 2 || j == N[1] - 2 || k == N[2] - 2;
                const LesGridTab<T> mt{g};
                if (phase == 1) {
                    T M[3][3];
                    if (edge)
                        les_collocated_gradient<T, true>(g, mt, e, U, I, M);
                    else
                        les_collocated_gradient<T, false>(g, mt, e, U, I, M);
                    const T dx = g.dl[0][i], dy = g.dl[1][j], dz = g.dl[2][k];
                    T out;
                    switch (model) {
                        case 1: out = eddy_visc_value<T, 1>(M, dx, dy, dz, (T)C); break;
                        case 2: out = eddy_visc_value<T, 2>(M, dx, dy, dz, (T)C); break;
                        case 3: out = eddy_visc_value<T, 3>(M, dx, dy, dz, (T)C); break;
                        default: out = eddy_visc_value<T, 4>(M, dx, dy, dz, (T)C); break;
                    }
                    visc[idx] = out;
                } else {
                    T fl[3];
                    if (edge)
                        les_stress_divergence<T, true>(g, mt, e, U, V, I, fl);
                    else
                        les_stress_divergence<T, false>(g, mt, e, U, V, I, fl);
                    for (int c = 0; c < 3; c++) f[idx + c * g.cs] += -fl[c];
                }
            }
}

extern "C" int host_les(int f64, int phase, int edge_all, const int *N, const void *const *tab15, const int *lo, const int *hi, int model,
                        double C, const void *u, void *visc, void *f) {
    const void *tab[3][5];
    for (int a = 0; a < 3; a++)
        for (int t = 0; t < 5; t++) tab[a][t] = tab15[a * 5 + t];
    if (f64)
        run<double>(phase, edge_all != 0, N, tab, lo, hi, model, C, (const double *)u, (double *)visc, (double *)f);
    else
        run<float>(phase, edge_all != 0, N, tab, lo, hi, model, C, (const float *)u, (float *)visc, (float *)f);
    return 0;
}
