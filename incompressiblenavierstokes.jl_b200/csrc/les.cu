// les.cu -- eddy-viscosity closures: eddy_viscosity_closure! with the
// Smagorinsky / WALE / Vreman / QR models, 3-D (src/eddyviscosity.jl:57-74, 235-378).
//
// Default: two fused kernels (below, cell functions in les_math.cuh) that store nothing but nu_t -- the
// staggered gradients and the stresses a cell needs are recomputed from the velocity, ghost entries by the
// rule zero_out_wall! + apply_bc_p! amount to.  20 scalar fields of traffic per closure instead of 62.
// Four-pass path (PressureBC sides, decomposed grids, INSB200_LES_UNFUSED): same sequence as the reference
// -- gradient tensor, wall zeroing + ghost fills, viscosity at the
// cell centres, ghost fills, symmetrise + scale with the interpolated viscosity, ghost fills,
// divergence of the stress -- with symmetrize! folded into the scaling pass (the ghost entries it
// would touch are overwritten by the fills that follow) and the divergence pass applying the
// result directly to the stage outputs (k, v = v + cv f, acc = acc + ca f) instead of writing f.
// The nine tensor components are one allocation, component (a, b) = du_a/dx_b at index a + 3 b
// (the reference's tensorfield order xx, yx, zx, xy, yy, zy, xz, yz, zz).
#include "stencil_math.cuh"
#include "les_math.cuh"

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

// MODEL: 1 Smagorinsky, 2 WALE, 3 Vreman, 4 QR (bodies in les_math.cuh)
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
    visc[idx] = eddy_visc_value<T, MODEL>(M, dx, dy, dz, C);
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

// ---- fused closure: two kernels, neither the gradient tensor nor the stress is ever stored --------------
// (1) nu_t at the cell centres straight from the velocity; (2) after the ghost fills of nu_t, the
// divergence of the stress with the stress recomputed at the six / nine points each cell needs, applied to
// the stage outputs.  Traffic per closure: 4 + (4 + 2 per output field) scalar fields instead of 62.
// Default: tile kernels -- a block marches a TX x TY column along z with the planes z-1, z, z+1 of
// (u_x, u_y, u_z[, nu_t]) and a one-cell halo in a 4-slot shared-memory ring (the next plane prefetched
// into registers during the arithmetic), so the ~55 / ~105 neighbour reads per cell are shared-memory
// reads.  INSB200_LES_TILE=0: one thread per cell, neighbours through L1 / L2.
struct LesBox {  // half-open index box the one-thread-per-cell kernels cover
    int lo[3], hi[3];
};

template <typename T, int MODEL>
__global__ void __launch_bounds__(256) les_visc_fused_kernel(Grid<T> g, LesEdge e, LesBox bx, const T *__restrict__ u,
                                                            T *__restrict__ visc, T C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + bx.lo[0], j = blockIdx.y * blockDim.y + threadIdx.y + bx.lo[1];
    const int k = blockIdx.z + bx.lo[2];
    if (i >= bx.hi[0] || j >= bx.hi[1]) return;
    const int I[3] = {i, j, k};
    const long long idx = i + j * g.s[1] + k * g.s[2];
    const LesGLoad<T> U{u + idx, g.s[1], g.s[2], g.cs};
    const LesGridTab<T> mt{g};
    T M[3][3];
    if (i == 1 || j == 1 || k == 1)
        les_collocated_gradient<T, true>(g, mt, e, U, I, M);
    else
        les_collocated_gradient<T, false>(g, mt, e, U, I, M);
    const T dx = __ldg(g.dl[0] + i), dy = __ldg(g.dl[1] + j), dz = __ldg(g.dl[2] + k);
    visc[idx] = eddy_visc_value<T, MODEL>(M, dx, dy, dz, C);
}

template <typename T, typename CT>
__device__ __forceinline__ void les_apply_cell(const Grid<T> &g, long long idx, const T (&f)[3], T *__restrict__ k_out,
                                               T *__restrict__ v, double cv, T *__restrict__ acc, double ca) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const T fl = -f[c];
        const long long q = idx + c * g.cs;
        if (k_out) k_out[q] += fl;
        if (v) v[q] = axpy_rn<T, CT>(v[q], cv, fl);
        if (acc) acc[q] = axpy_rn<T, CT>(acc[q], ca, fl);
    }
}

// The same in two halves for the tile kernel: the old values are requested before the stress arithmetic, so that
// their DRAM latency runs under it (the warps of a block march in lock-step behind one barrier per plane).
template <typename T>
struct LesOld {
    T k[3], v[3], a[3];
};
template <typename T>
__device__ __forceinline__ void les_apply_load(const Grid<T> &g, long long idx, const T *__restrict__ k_out,
                                               const T *__restrict__ v, const T *__restrict__ acc, LesOld<T> &o) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const long long q = idx + c * g.cs;
        if (k_out) o.k[c] = k_out[q];
        if (v) o.v[c] = v[q];
        if (acc) o.a[c] = acc[q];
    }
}
template <typename T, typename CT>
__device__ __forceinline__ void les_apply_store(const Grid<T> &g, long long idx, const T (&f)[3], const LesOld<T> &o,
                                                T *__restrict__ k_out, T *__restrict__ v, double cv, T *__restrict__ acc,
                                                double ca) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const T fl = -f[c];
        const long long q = idx + c * g.cs;
        if (k_out) k_out[q] = o.k[c] + fl;
        if (v) v[q] = axpy_rn<T, CT>(o.v[c], cv, fl);
        if (acc) acc[q] = axpy_rn<T, CT>(o.a[c], ca, fl);
    }
}

template <typename T, typename CT>
__global__ void __launch_bounds__(256) les_apply_fused_kernel(Grid<T> g, LesEdge e, LesBox bx, const T *__restrict__ u,
                                                             const T *__restrict__ visc, T *__restrict__ k_out,
                                                             T *__restrict__ v, double cv, T *__restrict__ acc, double ca) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + bx.lo[0], j = blockIdx.y * blockDim.y + threadIdx.y + bx.lo[1];
    const int k = blockIdx.z + bx.lo[2];
    if (i >= bx.hi[0] || j >= bx.hi[1]) return;
    const int I[3] = {i, j, k};
    const long long idx = i + j * g.s[1] + k * g.s[2];
    const LesGLoad<T> U{u + idx, g.s[1], g.s[2], g.cs}, V{visc + idx, g.s[1], g.s[2], 0};
    const LesGridTab<T> mt{g};
    T f[3];
    if (i == 1 || j == 1 || k == 1 || i == g.N[0] - 2 || j == g.N[1] - 2 || k == g.N[2] - 2)
        les_stress_divergence<T, true>(g, mt, e, U, V, I, f);
    else
        les_stress_divergence<T, false>(g, mt, e, U, V, I, f);
    les_apply_cell<T, CT>(g, idx, f, k_out, v, cv, acc, ca);
}

// shared planes of the tile kernels: pl[dz + 1] points at field 0 of this thread's cell in plane z + dz
template <typename T, int PX, int PLANE>
struct LesTileLoad {
    const T *pl[3];
    int f0;  // first field (the viscosity is field 3 of the apply kernel's ring)
    __device__ __forceinline__ T operator()(int c, int dx, int dy, int dz) const {
        return pl[dz + 1][(f0 + c) * PLANE + dy * PX + dx];
    }
};

// MODEL > 0: viscosity kernel (fields u_x, u_y, u_z; writes nu_t); MODEL == 0: apply kernel (fields u_x, u_y,
// u_z, nu_t; updates k_out / v / acc).  The cells next to a ghost plane are left to the one-thread-per-cell
// kernels (launch_les_shell): the ghost rules cost 46 registers here, i.e. one of two resident blocks.
// PREF (apply kernel): request the old output values before the stress arithmetic (les_apply_load); costs
// registers (Float64: capped at 128 by the launch bounds, a few spills) -- INSB200_LES_PREFETCH=0 switches it off.
template <typename T, typename CT, int MODEL, int TX, int TY, int ZC, bool PREF>
__global__ void __launch_bounds__(TX *TY, (MODEL == 0 && sizeof(T) == 8) ? 2 : (MODEL == 0 ? 3 : 4)) les_tile_kernel(Grid<T> g, LesEdge e, const T *__restrict__ u,
                                                         const T *__restrict__ visc_in, T *__restrict__ visc_out, T C,
                                                         T *__restrict__ k_out, T *__restrict__ v, double cv,
                                                         T *__restrict__ acc, double ca) {
    constexpr int NF = MODEL == 0 ? 4 : 3;
    constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY, NT = TX * TY, NE = NF * PLANE;
    constexpr int NL = (NE + NT - 1) / NT;
    constexpr int LX = TX + 2, LY = TY + 2, LZ = ZC + 2, LEN = LX + LY + LZ;  // metric-table slices: x, y, z back to back
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s = reinterpret_cast<T *>(smem_raw);  // [4 slots][NF fields][PY][PX]
    T *mtab = s + 4 * NE;                    // [rdl, rdu, wa, wb][LEN]
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = 1 + blockIdx.x * TX, y0 = 1 + blockIdx.y * TY;
    const int zbeg = 1 + blockIdx.z * ZC;
    const int zend = min(zbeg + ZC, g.N[2] - 1);
    if (zbeg >= zend) return;
    // entry r of a slice <-> index base - 1 + r of the dimension (clamped; clamped entries are never used)
    for (int el = tid; el < 4 * LEN; el += NT) {
        const int tb = el / LEN, r = el - tb * LEN;
        const int d = r < LX ? 0 : r < LX + LY ? 1 : 2;
        const int rl = d == 0 ? r : d == 1 ? r - LX : r - LX - LY;
        const int gi = min((d == 0 ? x0 : d == 1 ? y0 : zbeg) - 1 + rl, g.N[d] - 1);
        const T *tp = tb == 0 ? g.rdl[d] : tb == 1 ? g.rdu[d] : tb == 2 ? g.wa[d] : g.wb[d];
        mtab[el] = __ldg(tp + gi);
    }
    LesSmemTab<T, LEN> mt;  // addressed by the absolute index: entry of index 0 = slice start - (base - 1)
    mt.p[0] = mtab - (x0 - 1), mt.p[1] = mtab + LX - (y0 - 1), mt.p[2] = mtab + LX + LY - (zbeg - 1);
    const T *lp[NL];
    bool lv[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const int el = min(tid + i * NT, NE - 1);
        const int c = el / PLANE, r = el - c * PLANE;
        const int py = r / PX, px = r - py * PX;
        const int gx = min(x0 - 1 + px, g.N[0] - 1), gy = min(y0 - 1 + py, g.N[1] - 1);
        lv[i] = tid + i * NT < NE;
        const T *base = c < 3 ? u + c * g.cs : visc_in;
        lp[i] = base + gx + (long long)gy * g.s[1] + (long long)(zbeg - 1) * g.s[2];
    }
    T reg[NL];
    const long long zs = g.s[2];
    auto load_plane = [&]() {
#pragma unroll
        for (int i = 0; i < NL; i++) {
            if (lv[i]) reg[i] = __ldg(lp[i]);
            lp[i] += zs;
        }
    };
    auto store_plane = [&](int z) {
        T *dst = s + (z & 3) * NE + tid;
#pragma unroll
        for (int i = 0; i < NL; i++)
            if (lv[i]) dst[i * NT] = reg[i];
    };
    for (int p = -1; p <= 1; p++) {
        load_plane();
        store_plane(zbeg + p);
    }
    const int I0 = x0 + tx, I1 = y0 + ty;
    const bool active = I0 < g.N[0] - 1 && I1 < g.N[1] - 1;
    const int Ic0 = min(I0, g.N[0] - 2), Ic1 = min(I1, g.N[1] - 2);
    const bool edge_xy = Ic0 == 1 || Ic1 == 1 || Ic0 == g.N[0] - 2 || Ic1 == g.N[1] - 2;
    __syncthreads();
    const int lidx = (ty + 1) * PX + tx + 1;
    long long idx = Ic0 + (long long)Ic1 * g.s[1] + (long long)zbeg * g.s[2];
    for (int z = zbeg; z < zend; z++, idx += zs) {
        const bool more = z + 1 < zend;
        if (more) load_plane();
        if (active) {
            const int I[3] = {Ic0, Ic1, z};
            LesTileLoad<T, PX, PLANE> U, V;
#pragma unroll
            for (int k = 0; k < 3; k++) U.pl[k] = V.pl[k] = s + ((z - 1 + k) & 3) * NE + lidx;
            U.f0 = 0, V.f0 = 3;
            const bool edge = edge_xy || z == 1 || z == g.N[2] - 2;
            if (edge) {
            } else if (MODEL > 0) {
                T M[3][3];
                les_collocated_gradient<T, false>(g, mt, e, U, I, M);
                const T dx = __ldg(g.dl[0] + Ic0), dy = __ldg(g.dl[1] + Ic1), dz = __ldg(g.dl[2] + z);
                visc_out[idx] = eddy_visc_value<T, (MODEL > 0 ? MODEL : 1)>(M, dx, dy, dz, C);
            } else {
                LesOld<T> old;
                if (PREF) les_apply_load<T>(g, idx, k_out, v, acc, old);
                T f[3];
                les_stress_divergence<T, false>(g, mt, e, U, V, I, f);
                if (PREF)
                    les_apply_store<T, CT>(g, idx, f, old, k_out, v, cv, acc, ca);
                else
                    les_apply_cell<T, CT>(g, idx, f, k_out, v, cv, acc, ca);
            }
        }
        if (more) store_plane(z + 2);
        __syncthreads();
    }
}

template <typename T, typename CT, int MODEL>
static int launch_les_tile(const Grid<T> &g, const LesEdge &e, const T *u, const T *visc_in, T *visc_out, T C, T *k_out, T *v,
                           double cv, T *acc, double ca, Launch L) {
    constexpr int TX = 32, TY = 8, ZC = 64, NF = MODEL == 0 ? 4 : 3;
    constexpr size_t smem = (size_t)4 * NF * (TX + 2) * (TY + 2) * sizeof(T) + (size_t)4 * (TX + 2 + TY + 2 + ZC + 2) * sizeof(T);
    dim3 b(TX, TY, 1);
    dim3 gr((g.N[0] - 2 + TX - 1) / TX, (g.N[1] - 2 + TY - 1) / TY, (g.N[2] - 2 + ZC - 1) / ZC);
    const char *pe = getenv("INSB200_LES_PREFETCH");
    auto kern = les_tile_kernel<T, CT, MODEL, TX, TY, ZC, false>;
    if (MODEL == 0 && (!pe || atoi(pe) != 0)) kern = les_tile_kernel<T, CT, MODEL, TX, TY, ZC, MODEL == 0>;
    if (smem > 48 * 1024) INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<gr, b, smem, L.stream>>>(g, e, u, visc_in, visc_out, C, k_out, v, cv, acc, ca);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

// The cells with a ghost-plane neighbour as at most six disjoint boxes (z faces, y faces without the z faces, x
// faces without both); fn(box) launches a one-thread-per-cell kernel on each.
template <typename T, typename F>
static int les_for_shell_boxes(const Grid<T> &g, F fn) {
    int face[3][2], nface[3], in_lo[3], in_hi[3];
    for (int d = 0; d < 3; d++) {
        face[d][0] = 1, face[d][1] = g.N[d] - 2;
        nface[d] = g.N[d] - 2 > 1 ? 2 : 1;
        in_lo[d] = 2, in_hi[d] = g.N[d] - 2;  // interior of the interior (may be empty)
    }
    for (int d = 2; d >= 0; d--)
        for (int f = 0; f < nface[d]; f++) {
            LesBox bx;
            bool empty = false;
            for (int c = 0; c < 3; c++) {
                if (c == d)
                    bx.lo[c] = face[d][f], bx.hi[c] = face[d][f] + 1;
                else if (c > d)
                    bx.lo[c] = in_lo[c], bx.hi[c] = in_hi[c];
                else
                    bx.lo[c] = 1, bx.hi[c] = g.N[c] - 1;
                empty = empty || bx.hi[c] <= bx.lo[c];
            }
            if (!empty) INSB_TRY(fn(bx));
        }
    return 0;
}

static void les_box_launch(const LesBox &bx, dim3 &gr, dim3 &blk) {
    const int ex = bx.hi[0] - bx.lo[0], ey = bx.hi[1] - bx.lo[1];
    blk = ex >= 64 ? dim3(64, 4, 1) : ex >= 8 ? dim3(8, 32, 1) : dim3(1, 256, 1);
    gr = dim3((ex + blk.x - 1) / blk.x, (ey + blk.y - 1) / blk.y, bx.hi[2] - bx.lo[2]);
}

template <typename T>
static int launch_les_visc_box(const Grid<T> &g, const LesEdge &e, const LesBox &bx, int model, const T *u, T *visc, T C,
                               Launch L) {
    dim3 gr, blk;
    les_box_launch(bx, gr, blk);
    switch (model) {
        case 1: les_visc_fused_kernel<T, 1><<<gr, blk, 0, L.stream>>>(g, e, bx, u, visc, C); break;
        case 2: les_visc_fused_kernel<T, 2><<<gr, blk, 0, L.stream>>>(g, e, bx, u, visc, C); break;
        case 3: les_visc_fused_kernel<T, 3><<<gr, blk, 0, L.stream>>>(g, e, bx, u, visc, C); break;
        default: les_visc_fused_kernel<T, 4><<<gr, blk, 0, L.stream>>>(g, e, bx, u, visc, C); break;
    }
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int launch_les_apply_box(const Grid<T> &g, const LesEdge &e, const LesBox &bx, const T *u, const T *visc, T *k_out,
                                T *v, double cv, T *acc, double ca, bool wide, Launch L) {
    dim3 gr, blk;
    les_box_launch(bx, gr, blk);
    if (wide)
        les_apply_fused_kernel<T, double><<<gr, blk, 0, L.stream>>>(g, e, bx, u, visc, k_out, v, cv, acc, ca);
    else
        les_apply_fused_kernel<T, T><<<gr, blk, 0, L.stream>>>(g, e, bx, u, visc, k_out, v, cv, acc, ca);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

// BC kinds of the ghost planes of a cell field after zero_out_wall! + apply_bc_p! (les_math.cuh); false when a
// side needs the four-pass path.
static bool les_edge(const BcDesc &b, LesEdge &e) {
    if (b.comm) return false;  // a ghost plane of the gradient would need two velocity planes of the neighbour
    for (int d = 0; d < 3; d++)
        for (int side = 0; side < 2; side++) {
            const int bc = b.bc_u[d][side];
            if (bc != INSB200_BC_PERIODIC && bc != INSB200_BC_SYMMETRIC && bc != INSB200_BC_DIRICHLET)
                return false;  // PressureBC: two ghost layers
            (side ? e.hi : e.lo)[d] = bc;
        }
    return true;
}

bool les_needs_tensor(int D, const int N[3], const BcDesc &b, const void *u, const void *k_out, const void *v,
                      const void *acc) {
    LesEdge e;
    if (D != 3 || getenv("INSB200_LES_UNFUSED") != nullptr) return true;  // (read per call: the tests toggle it)
    if (u == k_out || u == v || u == acc) return true;  // the fused apply kernel reads u around the cells it updates
    (void)N;
    return !les_edge(b, e);
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
    if (!les_needs_tensor(g.D, g.N, b, u, k_out, v, acc)) {
        LesEdge e;
        les_edge(b, e);
        const char *te = getenv("INSB200_LES_TILE");  // (read per call: the tests toggle it)
        const bool tile = !te || atoi(te) != 0;
        const bool wide = sizeof(T) == 8 || ct_double;
        LesBox all;
        for (int d = 0; d < 3; d++) all.lo[d] = 1, all.hi[d] = g.N[d] - 1;
        if (tile) {
            switch (model) {
                case 1: INSB_TRY((launch_les_tile<T, T, 1>(g, e, u, nullptr, visc, (T)C, nullptr, nullptr, 0, nullptr, 0, L))); break;
                case 2: INSB_TRY((launch_les_tile<T, T, 2>(g, e, u, nullptr, visc, (T)C, nullptr, nullptr, 0, nullptr, 0, L))); break;
                case 3: INSB_TRY((launch_les_tile<T, T, 3>(g, e, u, nullptr, visc, (T)C, nullptr, nullptr, 0, nullptr, 0, L))); break;
                default: INSB_TRY((launch_les_tile<T, T, 4>(g, e, u, nullptr, visc, (T)C, nullptr, nullptr, 0, nullptr, 0, L))); break;
            }
            INSB_TRY(les_for_shell_boxes<T>(g, [&](const LesBox &bx) { return launch_les_visc_box<T>(g, e, bx, model, u, visc, (T)C, L); }));
        } else {
            INSB_TRY(launch_les_visc_box<T>(g, e, all, model, u, visc, (T)C, L));
        }
        INSB_TRY(launch_zero_out_wall<T>(g, b, visc, 1, L));
        INSB_TRY(launch_bc_p_fields<T>(g, b, visc, 1, L));
        if (tile) {
            if (wide)
                INSB_TRY((launch_les_tile<T, double, 0>(g, e, u, visc, nullptr, (T)0, k_out, v, cv, acc, ca, L)));
            else
                INSB_TRY((launch_les_tile<T, T, 0>(g, e, u, visc, nullptr, (T)0, k_out, v, cv, acc, ca, L)));
            INSB_TRY(les_for_shell_boxes<T>(
                g, [&](const LesBox &bx) { return launch_les_apply_box<T>(g, e, bx, u, visc, k_out, v, cv, acc, ca, wide, L); }));
        } else {
            INSB_TRY(launch_les_apply_box<T>(g, e, all, u, visc, k_out, v, cv, acc, ca, wide, L));
        }
        return 0;
    }
    if (!G) {
        set_error("eddy_viscosity_closure!: the four-pass path needs the tensor scratch (les_needs_tensor)");
        return 1;
    }
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
