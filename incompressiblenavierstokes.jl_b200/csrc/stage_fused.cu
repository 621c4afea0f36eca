// stage_fused.cu -- the fused stage kernel of the periodic time-step path (see fused.h).
//
// One launch per Runge-Kutta stage replaces, for 3-D periodic boxes,
//   applypressure! + apply_bc_u! of the PREVIOUS stage (src/operators.jl:306-311, boundary_conditions.jl:273-285)
//   fill!(k, 0) + convectiondiffusion!                  (src/solver.jl:2-6, src/operators.jl:616-628)
//   copyto! + the stage-combine broadcasts              (step_explicit_runge_kutta.jl:24-32)
//   apply_bc_u! + divergence! + scalewithvolume!        (src/pressure.jl:44-46, src/operators.jl:218-247)
// with one pass over the fields.  A block of 9 warps owns a 64 x 8 column of cells and marches along z:
//   * warps 0..7: two x-adjacent cells per thread, packed arithmetic, z-face fluxes carried in registers
//     from plane to plane (the scheme of stage_pair_kernel, stencil.cu);
//   * warp 8 recomputes the stage output on the tile's low halo (k_y on row y0-1, k_x on column x0-1), so
//     that the block can difference ITS OWN output: d = div(v_new) * volume needs v_new at I and I - e;
//     the z neighbour is the thread's own value of the previous plane, x / y neighbours travel through a
//     lane shuffle and a small double-buffered shared-memory exchange, and d is emitted one plane late.
//   * input planes: the first stage of a step reads the caller's array (reference layout, 4-byte
//     cp.async: its Float32 rows are not 16-byte aligned); later stages read the unprojected v and the
//     pressure p of the previous stage from plan-owned padded arrays as TMA boxes
//     (cp.async.bulk.tensor, one instruction per plane and field, mbarrier completion) and form
//     u = v - G p into the plane ring on the fly -- the projected velocity is never stored.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "fused.h"
#include "pair_math.cuh"

namespace insb {

namespace {

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, void *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, void *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
template <typename T>
__device__ __forceinline__ void cpa_elem(T *smem_dst, const T *gsrc) {
    unsigned d = smem_u32(smem_dst);
    if (sizeof(T) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// separately rounded helpers (no contraction: the halo recomputation must reproduce the owner's bits)
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// periodic wrap of a 1-based index into [1, n]
__device__ __forceinline__ int wrap1(int i, int n) { return i < 1 ? i + n : (i > n ? i - n : i); }

// ---- shapes ---------------------------------------------------------------------------------
// IN: 0 = velocity from the reference-layout array (first stage, cp.async), 1 = lazy (v, p) by TMA
// VB: 0 = vbase (and acc_in) are the input velocity itself, 1 = vbase from the reference-layout array
//     (cp.async), 2 = vbase from an internal array (TMA)
template <typename T, int IN, int VB, bool ACC>
struct FShape {
    static constexpr int TXC = 64, TY = 7, NW = TY + 1, NT = 32 * NW, NTM = 32 * TY;  // 7 main warps + the halo warp:
    // 8 warps spread evenly over the 4 register-file partitions of the SM (two blocks per SM at <= 128 registers)
    // plane ring: column c <-> cell i0 - 4 + c (TMA boxes must start on a 16-byte boundary of the innermost
    // dimension: a Float32 box starting at cell i0 - 2 faults), row r <-> cell j0 - 2 + r
    static constexpr int PX = 72, PYU = TY + 3, PYP = TY + 4, PLANE = PX * PYU, NE = 3 * PLANE, NEP = PX * PYP;
    static constexpr int D = 2;  // planes (groups) in flight ahead of their use
    static constexpr int ZC = 128, ZT = ZC + 4;  // z tables: planes zbeg - 2 .. zend + 1
    static constexpr size_t al(size_t b) { return (b + 127) / 128 * 128; }
    static constexpr int NVRAW = IN ? D + 1 : 1, NPRAW = IN ? D + 2 : 1, NU = IN ? 3 : D + 2, NPW = D + 1;
    static constexpr int NVRAW_SM = IN ? NVRAW : 0, NPRAW_SM = IN ? NPRAW : 0;  // slots that exist in shared memory
    static constexpr size_t VRAW_SLOT = al(NE * sizeof(T)), PRAW_SLOT = al(NEP * sizeof(T)), U_SLOT = al(NE * sizeof(T));
    static constexpr int PWV_ELEMS = VB == 1 ? 3 * 2 * NTM + 64 + 8 : VB == 2 ? 3 * (TY + 1) * PX : 0;
    static constexpr size_t PWV_SLOT = al(PWV_ELEMS * sizeof(T));
    static constexpr bool PWA = ACC && VB != 0;
    static constexpr size_t PWA_SLOT = PWA ? al(3 * TY * 64 * sizeof(T)) : 0;
    static constexpr size_t OFF_VRAW = 0;
    static constexpr size_t OFF_PRAW = OFF_VRAW + NVRAW_SM * VRAW_SLOT;
    static constexpr size_t OFF_U = OFF_PRAW + NPRAW_SM * PRAW_SLOT;
    static constexpr size_t OFF_PWV = OFF_U + NU * U_SLOT;
    static constexpr size_t OFF_PWA = OFF_PWV + (VB ? NPW * PWV_SLOT : 0);
    static constexpr size_t OFF_EXY = OFF_PWA + (PWA ? NPW * PWA_SLOT : 0);
    static constexpr size_t OFF_EXX = OFF_EXY + al(2 * (TY + 1) * 64 * sizeof(T));
    static constexpr size_t OFF_ZT = OFF_EXX + al(2 * 8 * sizeof(T));
    // lazy projection (halo warp): 1/dx pairs and 1/dy per PAIR of the plane ring (flat pair index = row * 36 + pair)
    static constexpr int NPAIR = PLANE / 2;
    static constexpr size_t OFF_PRJ = OFF_ZT + al(8 * ZT * sizeof(T));
    static constexpr size_t OFF_BAR = OFF_PRJ + (IN ? al(3 * NPAIR * sizeof(T)) : 0);
    static constexpr size_t total = OFF_BAR + 128;
};

template <typename T, typename CT>
struct FusedDev {
    // reference-layout inputs
    const T *u_ref, *u_below, *vb_ref;
    int N0, N1;            // reference extents (ghosted) of u_ref / vb_ref
    long long rs2, rcs;    // plane and component stride of the reference layout
    // internal outputs (pointers to component 0, element (0, 0, 0) of the padded array)
    T *v_out, *acc_out, *d_out, *u0_out;
    long long s1, s2, cs;  // internal strides
    int n0, n1, nz;        // interior extents (local planes)
    int zwrap, zoff, n2g;
    CT cv, ca;
    T nu;
    T bf[3];
};

// ---------------------------------------------------------------------------------------------
template <typename T, typename CT, int IN, int VB, bool ACC>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 2 : 1)
    stage_fused_kernel(const Grid<T> g, const FusedDev<T, CT> a, const __grid_constant__ CUtensorMap tm_v,
                       const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_vb,
                       const __grid_constant__ CUtensorMap tm_acc) {
    typedef typename PairOf<T>::type P;
    typedef FShape<T, IN, VB, ACC> S;
    constexpr int PX = S::PX, PLANE = S::PLANE, NE = S::NE, D = S::D, ZT = S::ZT, NT = S::NT, NTM = S::NTM, ZC = S::ZC, TY = S::TY;
    constexpr int OX = FusedLayout::OX, OY = FusedLayout::OY, OZ = FusedLayout::OZ;
    extern __shared__ __align__(1024) unsigned char smem[];
    T *const EXY = reinterpret_cast<T *>(smem + S::OFF_EXY);
    T *const EXX = reinterpret_cast<T *>(smem + S::OFF_EXX);
    T *const zt = reinterpret_cast<T *>(smem + S::OFF_ZT);
    unsigned long long *const bars = reinterpret_cast<unsigned long long *>(smem + S::OFF_BAR);
    // ring slots are tracked as small rotating indices (no modulo arithmetic in the march)
    auto uslot = [&](int slot) { return reinterpret_cast<T *>(smem + S::OFF_U + (size_t)slot * S::U_SLOT); };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool halo = warp == TY;
    const int n0 = a.n0, n1 = a.n1, nz = a.nz;
    const int i0 = 1 + blockIdx.x * S::TXC, j0 = 1 + blockIdx.y * S::TY;  // first cell of the tile (1-based)
    const int zbeg = 1 + blockIdx.z * ZC;
    const int zend = min(zbeg + ZC, nz + 1);  // planes zbeg .. zend - 1
    if (zbeg >= zend) return;
    const int zs = zbeg - 1;  // first computed plane: its z component feeds the divergence of plane zbeg
    const T nu = a.nu;

    // ---- z tables (scalars), row r <-> plane zbeg - 2 + r:
    //  0: -nu*rdlg[z+1]  1: -nu*rdug[z]  2: rdu[z]  3: rdl[z]  4: a2[z]/2  5: a1[z+1]/2  6: rdlg[z]  7: dl[z]
    for (int e = tid; e < 8 * ZT; e += NT) {
        const int tb = e / ZT, r = e - tb * ZT;
        const int kw = wrap1(a.zoff + zbeg - 2 + r, a.n2g) - a.zoff;  // the owner's entry of this plane
        T v = tb == 0 ? -nu * g.rdlg[2][kw + 1] : tb == 1 ? -nu * g.rdug[2][kw] : tb == 2 ? g.rdu[2][kw]
            : tb == 3 ? g.rdl[2][kw] : tb == 4 ? T(0.5) * g.a2[2][kw] : tb == 5 ? T(0.5) * g.a1[2][kw + 1]
            : tb == 6 ? g.rdlg[2][kw] : g.dl[2][kw];
        zt[e] = v;
    }
    if (IN == 1 && tid == 0) {
#pragma unroll
        for (int b = 0; b < D + 1; b++) mbar_init(bars + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- this thread's cells ----
    // main warps: cells (Xa, Xb = Xa + 1) of row Y; warp 8: the same columns on the halo row j0 - 1
    const int ty = warp;
    const int Xa = min(i0 + 2 * lane, n0 - 1), Xb = Xa + 1;  // n0 is even: pairs never straddle the edge
    const int Yraw = halo ? j0 - 1 : j0 + ty;
    const int Y = wrap1(min(Yraw, n1), n1);
    const bool act = !halo && i0 + 2 * lane < n0 && Yraw <= n1;
    const int rowr = halo ? 1 : ty + 2;
    const int lidx = rowr * PX + 2 * lane + 4;

    // loop-invariant x (pairs over the two cells) and y (broadcast) metrics, like stage_pair_kernel
    const P quarter = pbc(T(0.25));
    const P x_lg0 = pmk(-nu * g.rdlg[0][Xa], -nu * g.rdlg[0][Xb]);
    const P x_lg1 = pmk(-nu * g.rdlg[0][Xa + 1], -nu * g.rdlg[0][Xb + 1]);
    const P x_ugm = pmk(-nu * g.rdug[0][Xa - 1], -nu * g.rdug[0][Xb - 1]);
    const P x_ug0 = pmk(-nu * g.rdug[0][Xa], -nu * g.rdug[0][Xb]);
    const P x_rdu = pmk(g.rdu[0][Xa], g.rdu[0][Xb]), x_rdl = pmk(g.rdl[0][Xa], g.rdl[0][Xb]);
    const P x_hw2 = pmk(T(0.5) * g.a2[0][Xa], T(0.5) * g.a2[0][Xb]);
    const P x_hw1 = pmk(T(0.5) * g.a1[0][Xa + 1], T(0.5) * g.a1[0][Xb + 1]);
    const P y_lg0 = pbc(-nu * g.rdlg[1][Y]), y_lg1 = pbc(-nu * g.rdlg[1][Y + 1]);
    const P y_ugm = pbc(-nu * g.rdug[1][Y - 1]), y_ug0 = pbc(-nu * g.rdug[1][Y]);
    const P y_rdu = pbc(g.rdu[1][Y]), y_rdl = pbc(g.rdl[1][Y]);
    const P y_hw2 = pbc(T(0.5) * g.a2[1][Y]), y_hw1 = pbc(T(0.5) * g.a1[1][Y + 1]);
    // divergence metrics (divergence! + scalewithvolume!, src/operators.jl:218-247)
    const P dv_rx = pmk(g.rdlg[0][Xa], g.rdlg[0][Xb]), dv_dx = pmk(g.dl[0][Xa], g.dl[0][Xb]);
    const T dv_ry = g.rdlg[1][Y], dv_dy = g.dl[1][Y];

    // ---- cp.async plan of the reference-layout velocity planes (IN == 0) ----
    constexpr int NLD = 3 * S::PYU * 67, NL = (NLD + NT - 1) / NT;
    unsigned goff[IN == 0 ? NL : 1];
    int soff[IN == 0 ? NL : 1];
    if (IN == 0) {
#pragma unroll
        for (int q = 0; q < NL; q++) {
            const int e = min(tid + q * NT, NLD - 1);
            const int c = e / (S::PYU * 67), r = e - c * (S::PYU * 67);
            const int py = r / 67, px = r - py * 67;
            const int gi = wrap1(min(i0 - 2 + px, n0 + 2), n0), gj = wrap1(min(j0 - 2 + py, n1 + 2), n1);
            goff[q] = (unsigned)(c * a.rcs + gi + (long long)gj * a.N0);
            soff[q] = c * PLANE + py * PX + px + 2;
        }
    }
    int zu = zs - 1;  // next reference-layout velocity plane to fetch ...
    int fu = 0;       // ... into this slot (plane q sits in slot (q - (zs - 1)) mod NU)
    auto fetch_uref = [&]() {
        if (IN == 0) {
            if (zu <= zend) {
                T *dst = uslot(fu);
                if (zu < 0 && !a.zwrap) {  // plane -1 of the slab: sent by the rank below, [3][N1][N0]
#pragma unroll
                    for (int q = 0; q < NL; q++)
                        if ((q + 1) * NT <= NLD || tid + q * NT < NLD) {
                            const int c = soff[q] / PLANE;
                            cpa_elem<T>(dst + soff[q], a.u_below + (goff[q] - c * a.rcs) + (long long)c * a.N0 * a.N1);
                        }
                } else {
                    const T *pl = a.u_ref + (long long)(a.zwrap ? wrap1(zu, nz) : zu) * a.rs2;
#pragma unroll
                    for (int q = 0; q < NL; q++)
                        if ((q + 1) * NT <= NLD || tid + q * NT < NLD) cpa_elem<T>(dst + soff[q], pl + goff[q]);
                }
            }
            zu++;
            fu = fu + 1 == S::NU ? 0 : fu + 1;
        }
    };
    // ---- pointwise operands of plane k ----
    // VB == 1: vbase from the reference layout, 4-byte copies: main threads [c][thread][2], warp 8: the halo
    // row (y component) and the halo column (x component)
    const long long pw_off = (long long)min(i0 + 2 * lane, n0 - 1) + (long long)wrap1(min(Yraw, n1), n1) * a.N0;
    auto fetch_pw_ref = [&](int k, int slot) {
        if (VB == 1) {
            if (k >= zs && k < zend) {
                T *dst = reinterpret_cast<T *>(smem + S::OFF_PWV + (size_t)slot * S::PWV_SLOT);
                const T *src = a.vb_ref + (long long)(a.zwrap ? wrap1(k, nz) : k) * a.rs2;
                if (!halo) {
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        cpa_elem<T>(dst + (c * NTM + ty * 32 + lane) * 2, src + c * a.rcs + pw_off);
                        cpa_elem<T>(dst + (c * NTM + ty * 32 + lane) * 2 + 1, src + c * a.rcs + pw_off + 1);
                    }
                } else {
                    cpa_elem<T>(dst + 6 * NTM + 2 * lane, src + a.rcs + pw_off);
                    cpa_elem<T>(dst + 6 * NTM + 2 * lane + 1, src + a.rcs + pw_off + 1);
                    if (lane < TY)
                        cpa_elem<T>(dst + 6 * NTM + 64 + lane, src + wrap1(i0 - 1, n0) + (long long)min(j0 + lane, n1) * a.N0);
                }
            }
        }
    };
    // ---- TMA group j = { vraw(j + 2), praw(j + 3) [, praw(j + 2) for the first group], pw(j) } ----
    const int bx0 = i0 + OX - 1, by0 = j0 + OY - 1;  // internal indices of the tile's first cell
    // group j lives in barrier / vraw / pointwise slot sb = (j - zfirst) mod (D + 1) and praw slot
    // sp = (j - zfirst) mod (D + 2)
    auto issue_group = [&](int j, int sb, int sp, bool first) {
        if (IN == 1) {
            void *bar = bars + sb;
            unsigned bytes = 0;
            const bool pw = j >= zs && j < zend;
            bytes += (unsigned)(NE * sizeof(T)) + (unsigned)(S::NEP * sizeof(T)) * (first ? 2u : 1u);
            if (pw && VB == 2) bytes += (unsigned)(3 * (TY + 1) * PX * sizeof(T));
            if (pw && S::PWA) bytes += (unsigned)(3 * TY * 64 * sizeof(T));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            const int qv = j + 2, qp = j + 3;
            tma_load_4d(smem + S::OFF_VRAW + (size_t)sb * S::VRAW_SLOT, &tm_v, bar, bx0 - 4, by0 - 2, qv + OZ - 1, 0);
            tma_load_3d(smem + S::OFF_PRAW + (size_t)sp * S::PRAW_SLOT, &tm_p, bar, bx0 - 4, by0 - 2, qp + OZ - 1);
            if (first)  // praw(j + 2) of the very first group: slot "-1"
                tma_load_3d(smem + S::OFF_PRAW + (size_t)(D + 1) * S::PRAW_SLOT, &tm_p, bar, bx0 - 4, by0 - 2, qp - 1 + OZ - 1);
            if (pw && VB == 2)
                tma_load_4d(smem + S::OFF_PWV + (size_t)sb * S::PWV_SLOT, &tm_vb, bar, bx0 - 4, by0 - 1, j + OZ - 1, 0);
            if (pw && S::PWA)
                tma_load_4d(smem + S::OFF_PWA + (size_t)sb * S::PWA_SLOT, &tm_acc, bar, bx0, by0, j + OZ - 1, 0);
        }
    };
    // ---- lazy projection of plane q: U(q) = vraw(q) - G praw   (applypressure!, src/operators.jl:306-311) ----
    // Done by the halo warp alone (its stencil share is a third of a main warp's): the plane ring is PX = 72 wide, so
    // flat pair index e <-> elements 2e, 2e + 1 of the [PYU][PX] plane; lane l takes pairs l, l + 32, ...  The
    // metric factors of every pair sit in shared memory (x: a pair of 1/dxu, y: 1/dyu of the row).
    constexpr int NPAIR = S::NPAIR, NPQ = (NPAIR + 31) / 32;
    static_assert(PX % 2 == 0, "pairs must not straddle rows");
    P *const prj_rx = reinterpret_cast<P *>(smem + S::OFF_PRJ);
    T *const prj_ry = reinterpret_cast<T *>(smem + S::OFF_PRJ) + 2 * NPAIR;
    if (IN == 1) {
        for (int e = tid; e < NPAIR; e += NT) {
            const int r = e / (PX / 2), col = 2 * (e - r * (PX / 2));
            prj_rx[e] = pmk(g.rdu[0][wrap1(min(i0 - 4 + col, n0 + 2), n0)], g.rdu[0][wrap1(min(i0 - 3 + col, n0 + 2), n0)]);
            prj_ry[e] = g.rdu[1][wrap1(min(j0 - 2 + r, n1 + 2), n1)];
        }
    }
    // separately rounded pair operations (no contraction: u - ((p1 - p0) * r), each step rounded like the scalar code)
    auto sub2 = [](P a, P b) { return psub_rn(a, b); };
    auto mul2 = [](P a, P b) { return pmul_rn(a, b); };
    // vraw(q) sits in slot sv, praw(q) / praw(q + 1) in slots spm / sp, the result goes to ring slot su.
    // Work split: the halo warp takes the first NQH * 32 pairs, the main threads the rest.  Measured at 512^3 Float32
    // (rhs + combine per step): scalar projection on the main threads 7.98 ms; packed pairs, NQH = 0 / 2 / 3 / 4 / 6:
    // 7.92 / 7.86 / 7.99 / 8.10 / 8.58 ms; all twelve on the halo warp 9.34 ms (it becomes the critical path).
    constexpr int NQH = 2;
    auto project_plane = [&](int sv, int spm, int sp, int su, const T rz) {
        if (IN == 1) {
            const T *prs0 = reinterpret_cast<const T *>(smem + S::OFF_PRAW + (size_t)spm * S::PRAW_SLOT);
            const P *vr0 = reinterpret_cast<const P *>(smem + S::OFF_VRAW + (size_t)sv * S::VRAW_SLOT);
            const P *pu0 = reinterpret_cast<const P *>(smem + S::OFF_PRAW + (size_t)sp * S::PRAW_SLOT);
            P *uo0 = reinterpret_cast<P *>(uslot(su));
            const P rz2 = pbc(rz);
            auto task = [&](int e) {  // pair e of the plane
                const T *prs = prs0 + 2 * e;
                const P pc = *reinterpret_cast<const P *>(prs);
                const P px = pmk(pc.y, prs[2]);
                const P py = *reinterpret_cast<const P *>(prs + PX);
                const P rx = prj_rx[e];
                const P ry = pbc(prj_ry[e]);
                uo0[e] = sub2(vr0[e], mul2(sub2(px, pc), rx));
                uo0[NPAIR + e] = sub2(vr0[NPAIR + e], mul2(sub2(py, pc), ry));
                uo0[2 * NPAIR + e] = sub2(vr0[2 * NPAIR + e], mul2(sub2(pu0[e], pc), rz2));
            };
            if (halo) {
#pragma unroll
                for (int q = 0; q < NQH; q++) task(lane + 32 * q);
            } else {
                for (int e = 32 * NQH + tid; e < NPAIR; e += NTM) task(e);
            }
        }
    };

    // ---- shared-memory readers of the plane ring (thread's pair at pl) ----
    auto AL = [&](const T *pl, int c, int dy) { return *reinterpret_cast<const P *>(pl + c * PLANE + dy * PX); };
    auto XP = [&](const T *pl, int c, int dy) { return pmk(pl[c * PLANE + dy * PX + 1], pl[c * PLANE + dy * PX + 2]); };
    auto XM = [&](const T *pl, int c, int dy) { return pmk(pl[c * PLANE + dy * PX - 1], pl[c * PLANE + dy * PX]); };

    // ---- carried state ----
    P Fb0 = pbc(T(0)), Fb1 = pbc(T(0)), Fb2 = pbc(T(0));  // fluxes through the bottom faces of the current cells
    P vn0 = pbc(T(0)), vn1 = pbc(T(0)), vn2 = pbc(T(0));  // stage output of the previous plane (divergence is one plane late)
    P vzm = pbc(T(0));                                    // its z component one plane further down
    T hFb = T(0);                                         // warp 8, lanes < 8: bottom flux of the halo column (x component)
    // fluxes through the top faces of the plane in slot slo (z-table row ztr) = bottom faces of the plane in slot shi
    auto init_bottom = [&](int slo, int shi, const T *ztr) {
        const T *lo = uslot(slo) + lidx, *hi = uslot(shi) + lidx;
        const P z_lg = pbc(ztr[0]), z_ug = pbc(ztr[ZT]);
        const P u2 = AL(lo, 2, 0);
        if (!halo) {
            Fb2 = flux_diag<P>(u2, AL(hi, 2, 0), z_lg, quarter);
            Fb0 = flux_off<P>(AL(lo, 0, 0), AL(hi, 0, 0), u2, XP(lo, 2, 0), x_hw2, x_hw1, z_ug);
        }
        Fb1 = flux_off<P>(AL(lo, 1, 0), AL(hi, 1, 0), u2, AL(lo, 2, 1), y_hw2, y_hw1, z_ug);
        if (halo && lane < TY) {
            const int hi_i = wrap1(i0 - 1, n0);
            const T *l2 = uslot(slo) + (lane + 2) * PX + 3, *h2 = uslot(shi) + (lane + 2) * PX + 3;
            const P w2 = pbc(T(0.5) * g.a2[0][hi_i]), w1 = pbc(T(0.5) * g.a1[0][hi_i + 1]);
            hFb = flux_off<P>(pbc(l2[0]), pbc(h2[0]), pbc(l2[2 * PLANE]), pbc(l2[2 * PLANE + 1]), w2, w1, z_ug).x;
        }
    };

    // this thread's pair in the internal layout (plane offset `zo` is added per iteration)
    const long long q_int = (long long)(min(i0 + 2 * lane, n0 - 1) + OX - 1) + (long long)(min(Yraw, n1) + OY - 1) * a.s1;
    T *const v_b = a.v_out + q_int;
    T *const ao_b = ACC ? a.acc_out + q_int : nullptr;
    T *const d_b = a.d_out + q_int;
    // periodic ghost images of this pair: offsets in x and y (0 = none)
    const int ci = i0 + 2 * lane, cj = Yraw;
    const long long gx = ci <= 2 ? n0 : (ci >= n0 - 1 ? -n0 : 0);
    const long long gy = (cj <= 2 ? (long long)n1 : (cj >= n1 - 1 ? -(long long)n1 : 0)) * a.s1;

    // divergence of the stage output of plane k - 1 (vn*, vzm), emitted at the start of iteration k
    auto emit_div = [&](int kprev, const T *ztr) {
        const int par = kprev & 1;
        T up = __shfl_up_sync(0xffffffffu, vn0.y, 1);
        if (lane == 0) up = EXX[par * 8 + ty];
        const P vxm = pmk(up, vn0.x);
        const P vym = *reinterpret_cast<const P *>(EXY + (par * (TY + 1) + ty) * 64 + 2 * lane);
        P dd = padd(pmul(psub(vn0, vxm), dv_rx), pmul(psub(vn1, vym), pbc(dv_ry)));
        dd = padd(dd, pmul(psub(vn2, vzm), pbc(ztr[6 * ZT])));
        const P om = pmul(pmul(dv_dx, pbc(dv_dy)), pbc(ztr[7 * ZT]));
        dd = pmul(dd, om);
        if (act && kprev >= zbeg) *reinterpret_cast<P *>(d_b + (long long)(kprev + OZ - 1) * a.s2) = dd;
    };

    // =========================== prologue ===========================
    const int zfirst = IN == 1 ? zs - 3 : zs;
    const T *ztz = zt + (zfirst - (zbeg - 2));  // z-table row of plane z (valid from the first computed plane on)
    int su0, su1, su2 = 0;                      // ring slots of planes z, z + 1 (read) and z + 2 (lazily projected)
    int wb = 0, wpar = 0, wp = 0;               // group z: barrier / vraw / pointwise slot, barrier parity, praw slot
    if (IN == 0) {
        // planes zs - 1, zs, zs + 1 (group 0), then D - 1 more groups
        fetch_uref();
        fetch_uref();
        fetch_uref();
        cpa_commit();
#pragma unroll
        for (int d = 1; d < D; d++) {
            fetch_uref();
            cpa_commit();
        }
        cpa_wait<D - 1>();
        __syncthreads();
        init_bottom(0, 1, ztz - 1);
        su0 = 1, su1 = 2;
    } else {
        __syncthreads();  // barriers initialised, z tables written
        if (tid == TY * 32) {
#pragma unroll
            for (int d = 0; d < D; d++) issue_group(zfirst + d, d, d, d == 0);
        }
#pragma unroll
        for (int d = 0; d < D; d++) {
            fetch_pw_ref(zfirst + d, d);
            cpa_commit();
        }
        su0 = 1, su1 = 2;  // (U(z + 2) goes to slot (z - zfirst) mod 3)
    }

    // =========================== march ===========================
    long long zo = (long long)(zfirst + OZ - 1) * a.s2;  // plane z of the internal arrays
    for (int z = zfirst; z < zend; z++) {
        if (IN == 1) {
            mbar_wait(bars + wb, (unsigned)wpar);
            if (VB == 1) cpa_wait<D - 1>();
        } else {
            cpa_wait<D - 1>();
        }
        __syncthreads();
        // next loads
        if (IN == 1) {
            const int ib = wb == 0 ? D : wb - 1;              // (wb + D) mod (D + 1)
            const int ip = wp >= 2 ? wp - 2 : wp + D;         // (wp + D) mod (D + 2)
            if (tid == TY * 32 && z + D < zend) issue_group(z + D, ib, ip, false);
            fetch_pw_ref(z + D, ib);
            cpa_commit();
        } else {
            fetch_uref();  // plane z + 1 + D
            cpa_commit();
        }
        const bool comp = z >= zs;
        if (!halo && comp && z > zs) emit_div(z - 1, ztz - 1);
        if (IN == 1) {
            project_plane(wb, wp == 0 ? D + 1 : wp - 1, wp, su2, ztz[2 * ZT + 2]);
            if (z == zs - 1) init_bottom(su0, su1, ztz);  // U(zs - 1), U(zs): projected two / one iterations ago
        }
        if (comp) {
            const T *p0 = uslot(su0) + lidx, *pp = uslot(su1) + lidx;
            const T *ztr = ztz;
            const P z_lg = pbc(ztr[0]), z_ug = pbc(ztr[ZT]);
            const P z_rdu = pbc(ztr[2 * ZT]), z_rdl = pbc(ztr[3 * ZT]), z_hw2 = pbc(ztr[4 * ZT]), z_hw1 = pbc(ztr[5 * ZT]);
            const P u0c = AL(p0, 0, 0), u1c = AL(p0, 1, 0), u2c = AL(p0, 2, 0);
            const P u0xm = XM(p0, 0, 0), u1xm = XM(p0, 1, 0), u1xp = XP(p0, 1, 0);
            const P u0yp = AL(p0, 0, 1), u1ym = AL(p0, 1, -1), u1yp = AL(p0, 1, 1);
            const P u1zp = AL(pp, 1, 0);
            // y component (all nine warps)
            const P Ft1 = flux_off<P>(u1c, u1zp, u2c, AL(p0, 2, 1), y_hw2, y_hw1, z_ug);
            P k1 = pmul(psub(flux_off<P>(u1xm, u1c, u0xm, XM(p0, 0, 1), y_hw2, y_hw1, x_ugm),
                             flux_off<P>(u1c, u1xp, u0c, u0yp, y_hw2, y_hw1, x_ug0)), x_rdl);
            k1 = pfma(psub(flux_diag<P>(u1ym, u1c, y_lg0, quarter), flux_diag<P>(u1c, u1yp, y_lg1, quarter)), y_rdu, k1);
            k1 = pfma(psub(Fb1, Ft1), z_rdl, k1);
            Fb1 = Ft1;
            k1 = padd(k1, pbc(a.bf[1]));
            const int slot_pw = wb;
            if (!halo) {
                const P u0xp = XP(p0, 0, 0), u2xm = XM(p0, 2, 0), u2xp = XP(p0, 2, 0);
                const P u0ym = AL(p0, 0, -1), u2ym = AL(p0, 2, -1), u2yp = AL(p0, 2, 1);
                const P u0zp = AL(pp, 0, 0);
                const P Ft2 = flux_diag<P>(u2c, AL(pp, 2, 0), z_lg, quarter);
                const P Ft0 = flux_off<P>(u0c, u0zp, u2c, u2xp, x_hw2, x_hw1, z_ug);
                P k0 = pmul(psub(flux_diag<P>(u0xm, u0c, x_lg0, quarter), flux_diag<P>(u0c, u0xp, x_lg1, quarter)), x_rdu);
                k0 = pfma(psub(flux_off<P>(u0ym, u0c, u1ym, XP(p0, 1, -1), x_hw2, x_hw1, y_ugm),
                               flux_off<P>(u0c, u0yp, u1c, u1xp, x_hw2, x_hw1, y_ug0)), y_rdl, k0);
                k0 = pfma(psub(Fb0, Ft0), z_rdl, k0);
                P k2 = pmul(psub(flux_off<P>(u2xm, u2c, u0xm, XM(pp, 0, 0), z_hw2, z_hw1, x_ugm),
                                 flux_off<P>(u2c, u2xp, u0c, u0zp, z_hw2, z_hw1, x_ug0)), x_rdl);
                k2 = pfma(psub(flux_off<P>(u2ym, u2c, u1ym, AL(pp, 1, -1), z_hw2, z_hw1, y_ugm),
                               flux_off<P>(u2c, u2yp, u1c, u1zp, z_hw2, z_hw1, y_ug0)), y_rdl, k2);
                k2 = pfma(psub(Fb2, Ft2), z_rdu, k2);
                Fb0 = Ft0, Fb2 = Ft2;
                k0 = padd(k0, pbc(a.bf[0]));
                k2 = padd(k2, pbc(a.bf[2]));
                // pointwise operands
                P vb0, vb1, vb2, ai0, ai1, ai2;
                if (VB == 0) {
                    vb0 = u0c, vb1 = u1c, vb2 = u2c;
                } else if (VB == 1) {
                    const P *pw = reinterpret_cast<const P *>(smem + S::OFF_PWV + (size_t)slot_pw * S::PWV_SLOT) + ty * 32 + lane;
                    vb0 = pw[0], vb1 = pw[NTM], vb2 = pw[2 * NTM];
                } else {
                    const T *pw = reinterpret_cast<const T *>(smem + S::OFF_PWV + (size_t)slot_pw * S::PWV_SLOT) + (ty + 1) * PX + 2 * lane + 4;
                    vb0 = *reinterpret_cast<const P *>(pw), vb1 = *reinterpret_cast<const P *>(pw + (TY + 1) * PX);
                    vb2 = *reinterpret_cast<const P *>(pw + 2 * (TY + 1) * PX);
                }
                vzm = vn2;
                vn0 = PairAxpy<T, CT>::run(vb0, a.cv, k0);
                vn1 = PairAxpy<T, CT>::run(vb1, a.cv, k1);
                vn2 = PairAxpy<T, CT>::run(vb2, a.cv, k2);
                *reinterpret_cast<P *>(EXY + ((z & 1) * (TY + 1) + ty + 1) * 64 + 2 * lane) = vn1;
                if (act && z >= zbeg) {
                    const long long gz = !a.zwrap ? 0 : (z <= 2 ? (long long)nz : (z >= nz - 1 ? -(long long)nz : 0)) * a.s2;
                    // a pair and its periodic ghost images (two layers; x, y always, z on one GPU)
                    auto store_img = [&](T *o, const P val) {
                        *reinterpret_cast<P *>(o) = val;
                        if (gx | gy | gz) {
                            if (gx) *reinterpret_cast<P *>(o + gx) = val;
                            if (gy) {
                                *reinterpret_cast<P *>(o + gy) = val;
                                if (gx) *reinterpret_cast<P *>(o + gy + gx) = val;
                            }
                            if (gz) {
                                *reinterpret_cast<P *>(o + gz) = val;
                                if (gx) *reinterpret_cast<P *>(o + gz + gx) = val;
                                if (gy) {
                                    *reinterpret_cast<P *>(o + gz + gy) = val;
                                    if (gx) *reinterpret_cast<P *>(o + gz + gy + gx) = val;
                                }
                            }
                        }
                    };
                    store_img(v_b + zo, vn0);
                    store_img(v_b + zo + a.cs, vn1);
                    store_img(v_b + zo + 2 * a.cs, vn2);
                    if (IN == 1 && VB == 0 && a.u0_out) {  // chained step: keep the projected input of the step
                        T *u0_b = a.u0_out + q_int + zo;
                        store_img(u0_b, u0c);
                        store_img(u0_b + a.cs, u1c);
                        store_img(u0_b + 2 * a.cs, u2c);
                    }
                    if (ACC) {
                        // (the accumulator is the vbase of the last stage, which reads it on its halo too)
                        if (VB == 0) {
                            ai0 = u0c, ai1 = u1c, ai2 = u2c;
                        } else {
                            const T *pa = reinterpret_cast<const T *>(smem + S::OFF_PWA + (size_t)slot_pw * S::PWA_SLOT) + ty * 64 + 2 * lane;
                            ai0 = *reinterpret_cast<const P *>(pa), ai1 = *reinterpret_cast<const P *>(pa + TY * 64);
                            ai2 = *reinterpret_cast<const P *>(pa + 2 * TY * 64);
                        }
                        store_img(ao_b + zo, PairAxpy<T, CT>::run(ai0, a.ca, k0));
                        store_img(ao_b + zo + a.cs, PairAxpy<T, CT>::run(ai1, a.ca, k1));
                        store_img(ao_b + zo + 2 * a.cs, PairAxpy<T, CT>::run(ai2, a.ca, k2));
                    }
                }
            } else {
                // warp 8, row part: stage output (y component) of the halo row j0 - 1
                P vb1;
                if (VB == 0) {
                    vb1 = u1c;
                } else if (VB == 1) {
                    vb1 = *reinterpret_cast<const P *>(reinterpret_cast<const T *>(smem + S::OFF_PWV + (size_t)slot_pw * S::PWV_SLOT) + 6 * NTM + 2 * lane);
                } else {
                    vb1 = *reinterpret_cast<const P *>(reinterpret_cast<const T *>(smem + S::OFF_PWV + (size_t)slot_pw * S::PWV_SLOT) + (TY + 1) * PX + 2 * lane + 4);
                }
                *reinterpret_cast<P *>(EXY + ((z & 1) * (TY + 1)) * 64 + 2 * lane) = PairAxpy<T, CT>::run(vb1, a.cv, k1);
                // column part: x component at (i0 - 1, j0 + lane), lanes 0..TY-1 (both pair slots hold the same cell)
                if (lane < TY) {
                    const int ic = wrap1(i0 - 1, n0), jc = min(j0 + lane, n1);
                    const T *c0 = uslot(su0) + (lane + 2) * PX + 3, *cp = uslot(su1) + (lane + 2) * PX + 3;
                    const P cx_lg0 = pbc(-nu * g.rdlg[0][ic]), cx_lg1 = pbc(-nu * g.rdlg[0][ic + 1]);
                    const P cx_rdu = pbc(g.rdu[0][ic]);
                    const P cx_hw2 = pbc(T(0.5) * g.a2[0][ic]), cx_hw1 = pbc(T(0.5) * g.a1[0][ic + 1]);
                    const P cy_ugm = pbc(-nu * g.rdug[1][jc - 1]), cy_ug0 = pbc(-nu * g.rdug[1][jc]), cy_rdl = pbc(g.rdl[1][jc]);
                    const P s0c = pbc(c0[0]), s0xm = pbc(c0[-1]), s0xp = pbc(c0[1]);
                    const P s0ym = pbc(c0[-PX]), s0yp = pbc(c0[PX]);
                    const P s1c = pbc(c0[PLANE]), s1xp = pbc(c0[PLANE + 1]);
                    const P s1ym = pbc(c0[PLANE - PX]), s1ymxp = pbc(c0[PLANE - PX + 1]);
                    const P s2c = pbc(c0[2 * PLANE]), s2xp = pbc(c0[2 * PLANE + 1]);
                    const P Ft = flux_off<P>(s0c, pbc(cp[0]), s2c, s2xp, cx_hw2, cx_hw1, z_ug);
                    P kx = pmul(psub(flux_diag<P>(s0xm, s0c, cx_lg0, quarter), flux_diag<P>(s0c, s0xp, cx_lg1, quarter)), cx_rdu);
                    kx = pfma(psub(flux_off<P>(s0ym, s0c, s1ym, s1ymxp, cx_hw2, cx_hw1, cy_ugm),
                                   flux_off<P>(s0c, s0yp, s1c, s1xp, cx_hw2, cx_hw1, cy_ug0)), cy_rdl, kx);
                    kx = pfma(psub(pbc(hFb), Ft), z_rdl, kx);
                    hFb = Ft.x;
                    kx = padd(kx, pbc(a.bf[0]));
                    T vbx;
                    if (VB == 0)
                        vbx = c0[0];
                    else if (VB == 1)
                        vbx = reinterpret_cast<const T *>(smem + S::OFF_PWV + (size_t)slot_pw * S::PWV_SLOT)[6 * NTM + 64 + lane];
                    else
                        vbx = reinterpret_cast<const T *>(smem + S::OFF_PWV + (size_t)slot_pw * S::PWV_SLOT)[(lane + 1) * PX + 3];
                    EXX[(z & 1) * 8 + lane] = PairAxpy<T, CT>::run(pbc(vbx), a.cv, kx).x;
                }
            }
        }
        zo += a.s2;
        ztz++;
        if (IN == 1) {
            const int t3 = su0;
            su0 = su1, su1 = su2, su2 = t3;
            wb = wb == D ? 0 : wb + 1;
            wpar ^= wb == 0;
            wp = wp == D + 1 ? 0 : wp + 1;
        } else {
            su0 = su1;
            su1 = su1 + 1 == S::NU ? 0 : su1 + 1;
        }
    }
    // divergence of the last plane
    __syncthreads();
    if (!halo) emit_div(zend - 1, ztz - 1);
}


// ---------------------------------------------------------------------------------------------
// explicit projection of the last stage into the caller's array (reference layout) together with
// its periodic ghost images: applypressure! + apply_bc_u!
// ---------------------------------------------------------------------------------------------
constexpr int PRB = 4;
template <typename T>
__global__ void __launch_bounds__(128) project_out_kernel(const Grid<T> g, FusedLayout L, const T *__restrict__ v,
                                                          const T *__restrict__ p, T *__restrict__ u, int N0, int N1,
                                                          long long rs2, long long rcs, int zwrap, int zoff, int n2g) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j0 = blockIdx.y * PRB + 1;
    const int k = blockIdx.z + 1;
    const int n0 = L.n[0], n1 = L.n[1], nz = L.n[2];
    if (i > n0) return;
    const T rx = g.rdu[0][i];
    const T rz = g.rdu[2][wrap1(zoff + k, n2g) - zoff];
    const long long gx = i == 1 ? n0 : (i == n0 ? -n0 : 0);
    const long long gz = !zwrap ? 0 : (k == 1 ? (long long)nz : (k == nz ? -(long long)nz : 0)) * rs2;
    T pc[PRB], px[PRB], py[PRB], pz[PRB], vv[3][PRB];
#pragma unroll
    for (int r = 0; r < PRB; r++) {
        const int j = min(j0 + r, n1);
        const long long q = L.at(i, j, k);
        pc[r] = __ldg(p + q), px[r] = __ldg(p + q + 1), py[r] = __ldg(p + q + L.s1), pz[r] = __ldg(p + q + L.s2);
#pragma unroll
        for (int c = 0; c < 3; c++) vv[c][r] = __ldg(v + q + c * L.cs);
    }
#pragma unroll
    for (int r = 0; r < PRB; r++) {
        const int j = j0 + r;
        if (j > n1) break;
        const T ry = g.rdu[1][j];
        const long long gy = (j == 1 ? (long long)n1 : (j == n1 ? -(long long)n1 : 0)) * N0;
        T out[3];
        out[0] = sub_rn(vv[0][r], mul_rn(sub_rn(px[r], pc[r]), rx));
        out[1] = sub_rn(vv[1][r], mul_rn(sub_rn(py[r], pc[r]), ry));
        out[2] = sub_rn(vv[2][r], mul_rn(sub_rn(pz[r], pc[r]), rz));
#pragma unroll
        for (int c = 0; c < 3; c++) {
            T *o = u + c * rcs + i + (long long)j * N0 + (long long)k * rs2;
            const T val = out[c];
            o[0] = val;
            if (gx) o[gx] = val;
            if (gy) {
                o[gy] = val;
                if (gx) o[gy + gx] = val;
            }
            if (gz) {
                o[gz] = val;
                if (gx) o[gz + gx] = val;
                if (gy) {
                    o[gz + gy] = val;
                    if (gx) o[gz + gy + gx] = val;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

enum MapKind { MAP_V = 0, MAP_P = 1, MAP_VB = 2, MAP_ACC = 3 };

}  // namespace

struct FusedMaps {
    struct Entry {
        const void *ptr;
        int kind;
        CUtensorMap map;
    };
    std::vector<Entry> entries;
};

int fused_maps_create(FusedMaps **out) {
    *out = new FusedMaps();
    return 0;
}
void fused_maps_destroy(FusedMaps *m) { delete m; }

namespace {

template <typename T>
int get_map(FusedMaps *m, const FusedLayout &L, const void *ptr, int kind, const CUtensorMap **out) {
    for (auto &e : m->entries)
        if (e.ptr == ptr && e.kind == kind) {
            *out = &e.map;
            return 0;
        }
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    const bool vec = kind != MAP_P;
    const cuuint32_t rank = vec ? 4 : 3;
    cuuint64_t dims[4] = {(cuuint64_t)L.NPX, (cuuint64_t)L.NPY, (cuuint64_t)L.NPZ, 3};
    cuuint64_t strides[3] = {(cuuint64_t)L.s1 * sizeof(T), (cuuint64_t)L.s2 * sizeof(T), (cuuint64_t)L.cs * sizeof(T)};
    constexpr int TY = FShape<T, 1, 1, true>::TY;
    cuuint32_t box[4] = {FShape<T, 1, 1, true>::PX, TY + 3, 1, 3};
    if (kind == MAP_P) box[1] = TY + 4;
    if (kind == MAP_VB) box[1] = TY + 1;
    if (kind == MAP_ACC) box[0] = 64, box[1] = TY;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    FusedMaps::Entry e;
    e.ptr = ptr;
    e.kind = kind;
    CUresult r = enc(&e.map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
        return 1;
    }
    m->entries.push_back(e);
    *out = &m->entries.back().map;
    return 0;
}

template <typename T, typename CT, int IN, int VB, bool ACC>
int launch_variant(const Grid<T> &g, const FusedCtx &c, const FusedStage<T> &s, const Grid<T> &gl, Launch L) {
    typedef FShape<T, IN, VB, ACC> S;
    auto kern = stage_fused_kernel<T, CT, IN, VB, ACC>;
    static bool attr = false;
    if (!attr) {
        INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    FusedDev<T, CT> a{};
    a.u_ref = s.u_ref;
    a.u_below = s.u_below;
    a.vb_ref = s.vbase_ref;
    a.N0 = g.N[0];
    a.N1 = g.N[1];
    a.rs2 = g.s[2];
    a.rcs = g.cs;
    a.v_out = s.v_out;
    a.acc_out = s.acc_out;
    a.d_out = s.d_out;
    a.u0_out = s.u0_out;
    a.s1 = c.L.s1, a.s2 = c.L.s2, a.cs = c.L.cs;
    a.n0 = c.L.n[0], a.n1 = c.L.n[1], a.nz = c.L.n[2];
    a.zwrap = c.zwrap, a.zoff = c.zoff, a.n2g = c.n2g;
    a.cv = (CT)s.cv, a.ca = (CT)s.ca;
    a.nu = (T)s.nu;
    for (int q = 0; q < 3; q++) a.bf[q] = (T)s.bodyforce[q];
    // FusedMaps::entries may reallocate while maps are added: copy the maps by value
    CUtensorMap tv, tp, tvb, tacc;
    std::memset(&tv, 0, sizeof(tv));
    tp = tvb = tacc = tv;
    const CUtensorMap *m = nullptr;
    if (IN == 1) {
        INSB_TRY(get_map<T>(c.maps, c.L, s.v_in, MAP_V, &m));
        tv = *m;
        INSB_TRY(get_map<T>(c.maps, c.L, s.p_in, MAP_P, &m));
        tp = *m;
    }
    if (VB == 2) {
        INSB_TRY(get_map<T>(c.maps, c.L, s.vbase_int, MAP_VB, &m));
        tvb = *m;
    }
    if (S::PWA) {
        INSB_TRY(get_map<T>(c.maps, c.L, s.acc_in, MAP_ACC, &m));
        tacc = *m;
    }
    dim3 grid((a.n0 + S::TXC - 1) / S::TXC, (a.n1 + S::TY - 1) / S::TY, (a.nz + S::ZC - 1) / S::ZC);
    kern<<<grid, S::NT, S::total, L.stream>>>(gl, a, tv, tp, tvb, tacc);
    if (L.counter) ++*L.counter;
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        cudaFuncAttributes fa{};
        cudaFuncGetAttributes(&fa, kern);
        set_error(std::string("fused stage kernel launch: ") + cudaGetErrorString(le) + " (regs " + std::to_string(fa.numRegs) +
                  ", static smem " + std::to_string(fa.sharedSizeBytes) + ", dynamic smem " + std::to_string(S::total) +
                  ", max threads " + std::to_string(fa.maxThreadsPerBlock) + ", local " + std::to_string(fa.localSizeBytes) + ")");
        return 1;
    }
    return 0;
}

template <typename T, typename CT>
int launch_stage_fused_ct(const Grid<T> &g, const FusedCtx &c, const FusedStage<T> &s, Launch L) {
    const int in = s.u_ref ? 0 : 1;
    const int vb = s.vbase_ref ? 1 : (s.vbase_int ? 2 : 0);
    const bool acc = s.acc_out != nullptr;
    if (in == 0 && vb == 0 && acc) return launch_variant<T, CT, 0, 0, true>(g, c, s, g, L);
    if (in == 0 && vb == 0 && !acc) return launch_variant<T, CT, 0, 0, false>(g, c, s, g, L);
    if (in == 1 && vb == 1 && acc) return launch_variant<T, CT, 1, 1, true>(g, c, s, g, L);
    if (in == 1 && vb == 0 && acc) return launch_variant<T, CT, 1, 0, true>(g, c, s, g, L);  // chained first stage
    if (in == 1 && vb == 2 && acc) return launch_variant<T, CT, 1, 2, true>(g, c, s, g, L);  // chained middle stages
    if (in == 1 && vb == 2 && !acc) return launch_variant<T, CT, 1, 2, false>(g, c, s, g, L);
    set_error("fused stage: unsupported operand combination");
    return 1;
}

}  // namespace

template <typename T>
int launch_stage_fused(const Grid<T> &g, const FusedCtx &c, const FusedStage<T> &s, bool ct_double, Launch L) {
    if (sizeof(T) == 8 || ct_double) return launch_stage_fused_ct<T, double>(g, c, s, L);
    return launch_stage_fused_ct<T, T>(g, c, s, L);
}
template int launch_stage_fused<float>(const Grid<float> &, const FusedCtx &, const FusedStage<float> &, bool, Launch);
template int launch_stage_fused<double>(const Grid<double> &, const FusedCtx &, const FusedStage<double> &, bool, Launch);

template <typename T>
int launch_project_out(const Grid<T> &g, const FusedCtx &c, const T *v, const T *p, T *u_ref, Launch L) {
    dim3 b(128, 1, 1);
    dim3 gr((c.L.n[0] + 127) / 128, (c.L.n[1] + PRB - 1) / PRB, c.L.n[2]);
    project_out_kernel<T><<<gr, b, 0, L.stream>>>(g, c.L, v, p, u_ref, g.N[0], g.N[1], g.s[2], g.cs, c.zwrap, c.zoff, c.n2g);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_project_out<float>(const Grid<float> &, const FusedCtx &, const float *, const float *, float *, Launch);
template int launch_project_out<double>(const Grid<double> &, const FusedCtx &, const double *, const double *, double *, Launch);

}  // namespace insb
