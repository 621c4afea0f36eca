// stencil.cu -- stage kernel (right-hand side + Runge-Kutta stage combine), divergence,
// pressure gradient / projection update, CFL and max-abs reductions.
//
// Replaces, per stage: fill!(k,0) + convectiondiffusion! (src/solver.jl:2-6,
// src/operators.jl:616-628) [+ applygravity!, convection_diffusion_temp!, dissipation!,
// src/solver.jl:29-34] + copyto!(state, statestart) + the `@. u += Δt*A[i,j]*k` broadcasts
// (src/time_steppers/step_explicit_runge_kutta.jl:24-32) with ONE pass over the fields.
#include <cstdlib>

#include "pair_math.cuh"

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
__global__ void __launch_bounds__(256) stage_direct_kernel(Grid<T> g, StageArgs<T> a, int xb, int xe) {
    int I[3];
    I[0] = xb + blockIdx.x * blockDim.x + threadIdx.x;  // cells xb <= I0 < xe
    I[1] = blockIdx.y * blockDim.y + threadIdx.y;
    I[2] = (D == 3) ? (int)blockIdx.z : 0;
    if (I[0] >= xe || I[1] >= g.N[1]) return;
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
// Shared-memory plane ring shared by the marching kernels: planes z-1, z, z+1 of a few fields
// (with a one-cell xy halo) sit in shared memory, so every value is read from HBM/L2 once per
// block column instead of dozens of times through L1 as in the direct variant.
// ------------------------------------------------------------------------------------
template <typename T, int PX, int PLANE>
struct TileLoad {
    const T *pl[3];  // component 0 at this thread's cell in planes z-1, z, z+1
    __device__ __forceinline__ T operator()(int c, int dx, int dy, int dz) const {
        return pl[dz + 1][c * PLANE + dy * PX + dx];
    }
};

// Device-side arguments of the marching kernel: per-component base pointers (so that every
// access is base + 32-bit cell offset) and coefficients already converted to the types they
// are used in.
template <typename T, typename CT>
struct StageDev {
    const T *u[3];
    const T *vbase[3];
    T *v[3];
    const T *acc_in[3];
    T *acc_out[3];
    T *k_out[3];
    const T *kprev[MAXSTAGE][3];
    CT cprev[MAXSTAGE];
    int nprev;
    CT cv, ca;
    T nu;
    T bf[3];
};

// ------------------------------------------------------------------------------------
// stage kernel v3 ("pair" kernel): like the marching kernel above, but
//   * every thread owns TWO x-adjacent cells and does its arithmetic on 2-wide values:
//     packed FP32 (FFMA2/FADD2/FMUL2, sm_100) halves the issue slots of the flop-heavy
//     stencil, which is what bounded v2 (307 instructions per cell, 63 % issue-active);
//   * the z-direction face fluxes are carried in registers from one plane to the next
//     (the bottom face of cell z is the top face of cell z-1), so only planes z and z+1
//     are read from shared memory (3-slot ring) and 3 of the 18 face fluxes disappear;
//   * -nu/Δ, -nu/Δu and A/2 are folded into the metric tables once per block.
// Face flux of component i through a j-face at J (src/operators.jl:378-393,556):
//   i == j: ((u_i[J-e_i] + u_i[J])/2)^2 - nu (u_i[J] - u_i[J-e_i]) / Δ_i[J_i]
//   i != j: (u_i[J] + u_i[J+e_j])/2 * (A2_i u_j[J] + A1_i u_j[J+e_i]) - nu (u_i[J+e_j] - u_i[J]) / Δu_j[J_j]
// ------------------------------------------------------------------------------------
// cp.async (LDGSTS) of one element, and group bookkeeping
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(T) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Shared-memory footprint of the pair kernel (elements of T), shared by kernel and launcher.
template <int TY, int ZC, int DEPTH, bool ACC>
struct PairSmem {
    static constexpr int TXC = 64, PX = TXC + 4, PY = TY + 2, NE = 3 * PX * PY, NT = 32 * TY;
    static constexpr int NS = DEPTH + 2;                 // velocity planes z .. z+1+DEPTH
    static constexpr int NPW = DEPTH + 1;                // pointwise planes z .. z+DEPTH
    static constexpr int PWE = (ACC ? 2 : 1) * 3 * NT * 2;  // elements per pointwise slot
    static constexpr int ZT = ZC + 2;
    static constexpr int total = NS * NE + NPW * PWE + 6 * ZT * 2;
};

// Memory pipeline: all global reads are cp.async copies issued DEPTH planes ahead of their use
// (velocity planes into a (DEPTH+2)-slot ring shared by the block; the pointwise operands
// vbase / acc_in of a thread's own two cells into a private (DEPTH+1)-slot ring), so that
// 2 x DEPTH x ~20 KB are in flight per SM with only 16 resident warps.
// One pair (two consecutive elements, naturally aligned) per cp.async
template <typename T>
__device__ __forceinline__ void cp_async_pair(T *smem_dst, const T *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(T) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// AL8: N0 is even and the arrays are pair-aligned, so x tiles start at EVEN cell indices
// (tile bx covers cells 64 bx .. 64 bx + 63) and every global access of a cell pair is one
// naturally aligned 8-/16-byte transaction (cp.async of pairs, STG.64/128).  Otherwise tiles
// start at cell 1 and pairs are moved as two scalars.
// SAME: vbase (and acc_in) alias the velocity the stencil reads -- the first stage of every step --
// so the pointwise operands are the centre cells already sitting in the plane ring: no separate
// fetch of them at all.
template <typename T, typename CT, int TY, int ZC, int DEPTH, bool ACC, bool GEN, bool AL8, bool SAME = false>
__global__ void __launch_bounds__(32 * TY, sizeof(T) == 4 ? 2 : 1) stage_pair_kernel(Grid<T> g, StageDev<T, CT> a) {
    typedef typename PairOf<T>::type P;
    typedef PairSmem<TY, ZC, DEPTH, ACC> SM;
    constexpr int TXC = SM::TXC, LX = TXC + 2, PX = SM::PX, PY = SM::PY, PLANE = PX * PY, NE = SM::NE;
    constexpr int NT = SM::NT, NLD = 3 * PY * LX, NL = (NLD + NT - 1) / NT, ZT = SM::ZT, NS = SM::NS, NPW = SM::NPW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s = reinterpret_cast<T *>(smem_raw);              // [NS slots][3 comps][PY][PX]; cell x0 sits at column 2
    T *pw = s + NS * NE;                                 // [NPW slots][arrays][3 comps][NT] pairs
    P *zt = reinterpret_cast<P *>(pw + NPW * SM::PWE);   // [6 tables][ZT] of (v, v) pairs
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x0 = (AL8 ? 0 : 1) + blockIdx.x * TXC, y0 = 1 + blockIdx.y * TY;
    const int zbeg = 1 + blockIdx.z * ZC;
    const int zend = min(zbeg + ZC, g.N[2] - 1);
    if (zbeg >= zend) return;
    const T nu = a.nu;
    // this thread's share of a haloed plane; loop-invariant 32-bit offsets from a running plane pointer
    // (the three components are contiguous, 3 * cs < 2^32 is checked by the host).
    //  !AL8: element e = (comp, py, px), px in [0, LX) <-> gx = x0 - 1 + px
    //   AL8: NLC pair chunks (comp, py, j) <-> cells x0 + 2j, x0 + 2j + 1, plus one halo single for
    //        the first 6 PY threads (comp, py, side)
    constexpr int NCH = 3 * PY * (TXC / 2), NLC = (NCH + NT - 1) / NT, NG = AL8 ? NLC + 1 : NL;
    unsigned goff[NG];
    int soff[NG];
    if (AL8) {
#pragma unroll
        for (int i = 0; i < NLC; i++) {
            int e = min(tid + i * NT, NCH - 1);
            int c = e / (PY * (TXC / 2)), r = e - c * (PY * (TXC / 2));
            int py = r / (TXC / 2), j = r - py * (TXC / 2);
            int gx = min(x0 + 2 * j, g.N[0] - 2), gy = min(y0 - 1 + py, g.N[1] - 1);
            goff[i] = (unsigned)(c * g.cs + gx + (long long)gy * g.s[1]);
            soff[i] = c * PLANE + py * PX + 2 + 2 * j;
        }
        int e = min(tid, 6 * PY - 1);
        int c = e / (2 * PY), r = e - c * (2 * PY);
        int py = r >> 1, side = r & 1;
        int gx = side ? min(x0 + TXC, g.N[0] - 1) : max(x0 - 1, 0), gy = min(y0 - 1 + py, g.N[1] - 1);
        goff[NLC] = (unsigned)(c * g.cs + gx + (long long)gy * g.s[1]);
        soff[NLC] = c * PLANE + py * PX + (side ? TXC + 2 : 1);
    } else {
#pragma unroll
        for (int i = 0; i < NL; i++) {
            int e = min(tid + i * NT, NLD - 1);
            int c = e / (PY * LX), r = e - c * (PY * LX);
            int py = r / LX, px = r - py * LX;
            int gx = min(x0 - 1 + px, g.N[0] - 1), gy = min(y0 - 1 + py, g.N[1] - 1);
            goff[i] = (unsigned)(c * g.cs + gx + (long long)gy * g.s[1]);
            soff[i] = c * PLANE + py * PX + px + 1;
        }
    }
    const long long zs = g.s[2], cs = g.cs;
    const int I0 = x0 + 2 * tx, I1 = y0 + ty;
    const bool act0 = I0 >= 1 && I0 < g.N[0] - 1 && I1 < g.N[1] - 1, act1 = I0 + 1 < g.N[0] - 1 && I1 < g.N[1] - 1;
    const int Xa = max(1, min(I0, g.N[0] - 2)), Xb = max(1, min(I0 + 1, g.N[0] - 2)), Y = min(I1, g.N[1] - 2);
    const int Xp = AL8 ? min(I0, g.N[0] - 2) : Xa;  // first element of the pair in memory (even when AL8)
    const int dxb = AL8 ? 1 : Xb - Xa;              // 1, or 0 when the second cell is clamped
    const long long q0 = Xp + (long long)Y * g.s[1] + (long long)zbeg * zs;
    // issue state: next velocity plane / pointwise plane to fetch
    const T *uplane = a.u[0] + (long long)(zbeg - 1) * zs;
    int zu = zbeg - 1;           // plane index uplane stands on
    const T *vb_f = a.vbase[0] + q0;
    const T *ai_f = ACC ? a.acc_in[0] + q0 : nullptr;
    int zp = zbeg;               // plane index the pointwise fetch pointers stand on
    auto fetch_plane = [&]() {   // velocity plane zu -> slot zu % NS (skipped beyond the chunk)
        if (zu <= zend) {
            T *dst = s + (zu % NS) * NE;
            if (AL8) {
#pragma unroll
                for (int i = 0; i < NLC; i++)
                    if ((i + 1) * NT <= NCH || tid + i * NT < NCH) cp_async_pair<T>(dst + soff[i], uplane + goff[i]);
                if (tid < 6 * PY) cp_async_elem<T>(dst + soff[NLC], uplane + goff[NLC]);
            } else {
#pragma unroll
                for (int i = 0; i < NL; i++)
                    if ((i + 1) * NT <= NLD || tid + i * NT < NLD) cp_async_elem<T>(dst + soff[i], uplane + goff[i]);
            }
        }
        uplane += zs;
        zu++;
    };
    auto fetch_pointwise = [&]() {  // vbase / acc_in of this thread's cells in plane zp -> slot zp % NPW
        if (SAME) return;
        if (zp < zend) {
            T *dst = pw + (zp % NPW) * SM::PWE + 2 * tid;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (AL8) {
                    cp_async_pair<T>(dst + c * 2 * NT, vb_f + c * cs);
                    if (ACC) cp_async_pair<T>(dst + (3 + c) * 2 * NT, ai_f + c * cs);
                } else {
                    cp_async_elem<T>(dst + c * 2 * NT, vb_f + c * cs);
                    cp_async_elem<T>(dst + c * 2 * NT + 1, vb_f + c * cs + dxb);
                    if (ACC) {
                        cp_async_elem<T>(dst + (3 + c) * 2 * NT, ai_f + c * cs);
                        cp_async_elem<T>(dst + (3 + c) * 2 * NT + 1, ai_f + c * cs + dxb);
                    }
                }
            }
        }
        vb_f += zs;
        if (ACC) ai_f += zs;
        zp++;
    };
    // prologue: group 0 = planes zbeg-1, zbeg, zbeg+1 + pointwise zbeg; then DEPTH-1 more groups
    fetch_plane();
    fetch_plane();
    fetch_plane();
    fetch_pointwise();
    cp_async_commit();
#pragma unroll
    for (int d = 1; d < DEPTH; d++) {
        fetch_plane();
        fetch_pointwise();
        cp_async_commit();
    }
    // z tables, row r <-> plane z = zbeg - 1 + r:
    //  0: -nu*rdlg[z+1]  1: -nu*rdug[z]  2: rdu[z]  3: rdl[z]  4: a2[z]/2  5: a1[z+1]/2
    for (int e = tid; e < 6 * ZT; e += NT) {
        int tb = e / ZT, r = e - tb * ZT;
        int z = min(zbeg - 1 + r, g.N[2] - 2);
        T v = tb == 0 ? -nu * g.rdlg[2][z + 1] : tb == 1 ? -nu * g.rdug[2][z] : tb == 2 ? g.rdu[2][z]
            : tb == 3 ? g.rdl[2][z] : tb == 4 ? T(0.5) * g.a2[2][z] : T(0.5) * g.a1[2][z + 1];
        zt[e] = pbc(v);
    }
    // loop-invariant x (pairs over the two cells) and y (broadcast) metrics
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
    const int lidx = (ty + 1) * PX + 2 * tx + 2;  // even: pairs are 8-/16-byte aligned
    auto AL = [&](const T *pl, int c, int dy) { return *reinterpret_cast<const P *>(pl + c * PLANE + dy * PX); };
    auto XP = [&](const T *pl, int c, int dy) { return pmk(pl[c * PLANE + dy * PX + 1], pl[c * PLANE + dy * PX + 2]); };
    auto XM = [&](const T *pl, int c, int dy) { return pmk(pl[c * PLANE + dy * PX - 1], pl[c * PLANE + dy * PX]); };
    // z-face fluxes through the top face of the cells in the plane at `lo` (`hi` = the plane above)
    auto top_fluxes = [&](const T *lo, const T *hi, const P *ztr, P &F0, P &F1, P &F2) {
        const P z_lg = ztr[0 * ZT], z_ug = ztr[1 * ZT];
        const P u2 = AL(lo, 2, 0);
        F2 = flux_diag<P>(u2, AL(hi, 2, 0), z_lg, quarter);
        F0 = flux_off<P>(AL(lo, 0, 0), AL(hi, 0, 0), u2, XP(lo, 2, 0), x_hw2, x_hw1, z_ug);
        F1 = flux_off<P>(AL(lo, 1, 0), AL(hi, 1, 0), u2, AL(lo, 2, 1), y_hw2, y_hw1, z_ug);
    };
    cp_async_wait<DEPTH - 1>();
    __syncthreads();
    P Fb0, Fb1, Fb2;  // fluxes through the bottom face of the current cells
    top_fluxes(s + ((zbeg - 1) % NS) * NE + lidx, s + (zbeg % NS) * NE + lidx, zt, Fb0, Fb1, Fb2);
    T *v_p = a.v[0] + q0;
    T *ao_p = ACC ? a.acc_out[0] + q0 : nullptr;
    unsigned q = (unsigned)q0;  // GEN path only
    const P *ztr = zt + 1;
    int sl0 = zbeg % NS, sl1 = (zbeg + 1) % NS, pws = zbeg % NPW;
    for (int z = zbeg; z < zend; z++) {
        cp_async_wait<DEPTH - 1>();  // plane z+1 and the pointwise operands of plane z have landed
        __syncthreads();             // ... for every thread; and everybody is done with plane z-1
        fetch_plane();               // plane z+1+DEPTH into the slot of plane z-1
        fetch_pointwise();           // pointwise operands of plane z+DEPTH
        cp_async_commit();
        const T *p0 = s + sl0 * NE + lidx, *pp = s + sl1 * NE + lidx;
        const P z_rdu = ztr[2 * ZT], z_rdl = ztr[3 * ZT], z_hw2 = ztr[4 * ZT], z_hw1 = ztr[5 * ZT];
        P Ft0, Ft1, Ft2;
        top_fluxes(p0, pp, ztr, Ft0, Ft1, Ft2);
        const P u0c = AL(p0, 0, 0), u1c = AL(p0, 1, 0), u2c = AL(p0, 2, 0);
        const P u0xm = XM(p0, 0, 0), u0xp = XP(p0, 0, 0), u1xm = XM(p0, 1, 0), u1xp = XP(p0, 1, 0);
        const P u2xm = XM(p0, 2, 0), u2xp = XP(p0, 2, 0);
        const P u0ym = AL(p0, 0, -1), u0yp = AL(p0, 0, 1), u1ym = AL(p0, 1, -1), u1yp = AL(p0, 1, 1);
        const P u2ym = AL(p0, 2, -1), u2yp = AL(p0, 2, 1);
        const P u0zp = AL(pp, 0, 0), u1zp = AL(pp, 1, 0);
        P k[3];
        // x component
        k[0] = pmul(psub(flux_diag<P>(u0xm, u0c, x_lg0, quarter), flux_diag<P>(u0c, u0xp, x_lg1, quarter)), x_rdu);
        k[0] = pfma(psub(flux_off<P>(u0ym, u0c, u1ym, XP(p0, 1, -1), x_hw2, x_hw1, y_ugm),
                         flux_off<P>(u0c, u0yp, u1c, u1xp, x_hw2, x_hw1, y_ug0)), y_rdl, k[0]);
        k[0] = pfma(psub(Fb0, Ft0), z_rdl, k[0]);
        // y component
        k[1] = pmul(psub(flux_off<P>(u1xm, u1c, u0xm, XM(p0, 0, 1), y_hw2, y_hw1, x_ugm),
                         flux_off<P>(u1c, u1xp, u0c, u0yp, y_hw2, y_hw1, x_ug0)), x_rdl);
        k[1] = pfma(psub(flux_diag<P>(u1ym, u1c, y_lg0, quarter), flux_diag<P>(u1c, u1yp, y_lg1, quarter)), y_rdu, k[1]);
        k[1] = pfma(psub(Fb1, Ft1), z_rdl, k[1]);
        // z component
        k[2] = pmul(psub(flux_off<P>(u2xm, u2c, u0xm, XM(pp, 0, 0), z_hw2, z_hw1, x_ugm),
                         flux_off<P>(u2c, u2xp, u0c, u0zp, z_hw2, z_hw1, x_ug0)), x_rdl);
        k[2] = pfma(psub(flux_off<P>(u2ym, u2c, u1ym, AL(pp, 1, -1), z_hw2, z_hw1, y_ugm),
                         flux_off<P>(u2c, u2yp, u1c, u1zp, z_hw2, z_hw1, y_ug0)), y_rdl, k[2]);
        k[2] = pfma(psub(Fb2, Ft2), z_rdu, k[2]);
        Fb0 = Ft0, Fb1 = Ft1, Fb2 = Ft2;
        const P *pwp = reinterpret_cast<const P *>(pw + pws * SM::PWE) + tid;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const P kc = padd(k[c], pbc(a.bf[c]));
            const P uc = c == 0 ? u0c : c == 1 ? u1c : u2c;
            P x = SAME ? uc : pwp[c * NT];
            if (GEN) {
                if (a.k_out[0]) {
                    if (act0) a.k_out[c][q] = kc.x;
                    if (act1) a.k_out[c][q + 1] = kc.y;
                }
#pragma unroll 1
                for (int j = 0; j < a.nprev; j++) {
                    P kp = pmk(act0 ? a.kprev[j][c][q] : T(0), act1 ? a.kprev[j][c][q + 1] : T(0));
                    x = PairAxpy<T, CT>::run(x, a.cprev[j], kp);
                }
            }
            const P vo = PairAxpy<T, CT>::run(x, a.cv, kc);
            if (AL8 && act0 && act1) {
                *reinterpret_cast<P *>(v_p + c * cs) = vo;
            } else {
                if (act0) v_p[c * cs] = vo.x;
                if (act1) v_p[c * cs + dxb] = vo.y;
            }
            if (ACC) {
                const P ao = PairAxpy<T, CT>::run(SAME ? uc : pwp[(3 + c) * NT], a.ca, kc);
                if (AL8 && act0 && act1) {
                    *reinterpret_cast<P *>(ao_p + c * cs) = ao;
                } else {
                    if (act0) ao_p[c * cs] = ao.x;
                    if (act1) ao_p[c * cs + dxb] = ao.y;
                }
            }
        }
        sl0 = sl1;
        sl1 = sl1 + 1 == NS ? 0 : sl1 + 1;
        pws = pws + 1 == NPW ? 0 : pws + 1;
        v_p += zs, q += (unsigned)zs, ztr++;
        if (ACC) ao_p += zs;
    }
}

template <typename T, typename CT, int D>
static int launch_stage_t(const Grid<T> &g, const StageArgs<T> &a, Launch L, int xb = 0, int xe = -1);

template <typename T, typename CT, bool AL8>
static int launch_stage_pair_a(const Grid<T> &g, const StageArgs<T> &a, Launch L, int nxt) {
    constexpr int TY = 8, ZC = 128, DEPTH = 2;  // (DEPTH = 3 measured: rhs+combine 5.64 -> 6.06 ms/step)
    const bool acc = a.acc_out != nullptr, gen = a.k_out != nullptr || a.nprev > 0;
    const size_t smem = (size_t)(acc ? PairSmem<TY, ZC, DEPTH, true>::total : PairSmem<TY, ZC, DEPTH, false>::total) * sizeof(T);
    const bool same = !gen && a.vbase == a.u && (!acc || a.acc_in == a.u) && getenv("INSB200_STAGE_NO_SAME") == nullptr;
    auto kern = acc ? (gen ? stage_pair_kernel<T, CT, TY, ZC, DEPTH, true, true, AL8>
                           : (same ? stage_pair_kernel<T, CT, TY, ZC, DEPTH, true, false, AL8, true>
                                   : stage_pair_kernel<T, CT, TY, ZC, DEPTH, true, false, AL8>))
                    : (gen ? stage_pair_kernel<T, CT, TY, ZC, DEPTH, false, true, AL8>
                           : (same ? stage_pair_kernel<T, CT, TY, ZC, DEPTH, false, false, AL8, true>
                                   : stage_pair_kernel<T, CT, TY, ZC, DEPTH, false, false, AL8>));
    INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageDev<T, CT> d{};
    for (int c = 0; c < 3; c++) {
        d.u[c] = a.u + c * g.cs;
        d.vbase[c] = a.vbase ? a.vbase + c * g.cs : nullptr;
        d.v[c] = a.v ? a.v + c * g.cs : nullptr;
        d.acc_in[c] = a.acc_in ? a.acc_in + c * g.cs : nullptr;
        d.acc_out[c] = a.acc_out ? a.acc_out + c * g.cs : nullptr;
        d.k_out[c] = a.k_out ? a.k_out + c * g.cs : nullptr;
        for (int j = 0; j < a.nprev; j++) d.kprev[j][c] = a.kprev[j] + c * g.cs;
        d.bf[c] = (T)a.bodyforce[c];
    }
    d.nprev = a.nprev;
    for (int j = 0; j < a.nprev; j++) d.cprev[j] = (CT)a.cprev[j];
    d.cv = (CT)a.cv;
    d.ca = (CT)a.ca;
    d.nu = (T)a.nu;
    int nz = g.N[2] - 2;
    dim3 b(32, TY, 1);
    dim3 gr(nxt, (g.N[1] - 2 + TY - 1) / TY, (nz + ZC - 1) / ZC);
    kern<<<gr, b, smem, L.stream>>>(g, d);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, typename CT>
static int launch_stage_pair(const Grid<T> &g, const StageArgs<T> &a, Launch L) {
    constexpr int TXC = 64;
    static const bool no_al8 = getenv("INSB200_STAGE_UNALIGNED") != nullptr;
    const uintptr_t align = (uintptr_t)a.u | (uintptr_t)a.vbase | (uintptr_t)a.v | (uintptr_t)a.acc_in | (uintptr_t)a.acc_out;
    const bool al8 = !no_al8 && g.N[0] % 2 == 0 && align % (2 * sizeof(T)) == 0;
    if (!al8) return launch_stage_pair_a<T, CT, false>(g, a, L, (g.N[0] - 2 + TXC - 1) / TXC);
    // even-aligned tiles cover cells [0, 64 nxt); a narrow remainder goes to the direct kernel
    int nxt = (g.N[0] - 1) / TXC;
    int rest = g.N[0] - 1 - nxt * TXC;
    if (rest > 8 || nxt == 0) nxt++, rest = 0;
    INSB_TRY((launch_stage_pair_a<T, CT, true>(g, a, L, nxt)));
    if (rest > 0) INSB_TRY((launch_stage_t<T, CT, 3>(g, a, L, nxt * TXC, g.N[0] - 1)));
    return 0;
}

template <typename T, typename CT, int D>
static int launch_stage_t(const Grid<T> &g, const StageArgs<T> &a, Launch L, int xb, int xe) {
    if (xe < 0) xe = g.N[0];
    // narrow strips (the remainder next to the marching kernel's tiles): no idle x lanes
    int bx = 64;
    if (xe - xb <= 8) {
        bx = 1;
        while (bx < xe - xb) bx <<= 1;
    }
    dim3 b(bx, 256 / bx, 1);
    dim3 gr((xe - xb + b.x - 1) / b.x, (g.N[1] + b.y - 1) / b.y, D == 3 ? g.N[2] : 1);
    bool full = a.conv && a.diff;
    StageArgs<T> aa = a;
    if (!a.diff) aa.nu = 0.0;
    if (a.bouss) {
        if (full)
            stage_direct_kernel<T, CT, D, true, true><<<gr, b, 0, L.stream>>>(g, aa, xb, xe);
        else
            stage_direct_kernel<T, CT, D, true, false><<<gr, b, 0, L.stream>>>(g, aa, xb, xe);
    } else {
        if (full)
            stage_direct_kernel<T, CT, D, false, true><<<gr, b, 0, L.stream>>>(g, aa, xb, xe);
        else
            stage_direct_kernel<T, CT, D, false, false><<<gr, b, 0, L.stream>>>(g, aa, xb, xe);
    }
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------
// Boussinesq terms as a second pass behind the marching momentum kernel (3-D): temperature
// right-hand side (convection_diffusion_temp! + dissipation!, src/operators.jl:669-681,772-781)
// with its stage combine, and the buoyancy term g*avg(temp) (applygravity!, :802-829) added to
// the gravity component of the velocity outputs the momentum kernel has just written.
// v = (vbase + cv k) + cv g avg(T) instead of vbase + cv (k + g avg(T)): a rounding-level
// difference, far inside the parity tolerance.
// ------------------------------------------------------------------------------------
template <typename T, typename CT>
__global__ void __launch_bounds__(256) bouss_extra_kernel(Grid<T> g, StageArgs<T> a) {
    int I[3];
    I[0] = blockIdx.x * blockDim.x + threadIdx.x + 1;
    I[1] = blockIdx.y * blockDim.y + threadIdx.y + 1;
    I[2] = (int)blockIdx.z + 1;
    if (I[0] >= g.N[0] - 1 || I[1] >= g.N[1] - 1) return;
    const long long idx = I[0] + I[1] * g.s[1] + I[2] * g.s[2];
    GLoad<T, 3> U{a.u + idx, g.s[1], g.s[2], g.cs};
    GLoad<T, 3> Tm{a.temp + idx, g.s[1], g.s[2], 0};
    const T kt = temperature_rhs<T, 3>(U, Tm, g, I, (T)a.kappa, (T)a.nu, a.dodiss != 0);
    if (a.tk_out) a.tk_out[idx] = kt;
    if (a.tv) {
        T x = a.tvbase[idx];
        for (int j = 0; j < a.nprev; j++) x = axpy_rn<T, CT>(x, a.cprev[j], a.tkprev[j][idx]);
        a.tv[idx] = axpy_rn<T, CT>(x, a.cv, kt);
    }
    if (a.tacc_out) a.tacc_out[idx] = axpy_rn<T, CT>(a.tacc_in[idx], a.ca, kt);
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (c != a.gdir) continue;
        const T gk = (T)a.grav * avg_face<3>(Tm, g, I, c, 0);
        const long long q = idx + c * g.cs;
        if (a.k_out) a.k_out[q] += gk;
        if (a.v) a.v[q] = axpy_rn<T, CT>(a.v[q], a.cv, gk);
        if (a.acc_out) a.acc_out[q] = axpy_rn<T, CT>(a.acc_out[q], a.ca, gk);
    }
}

// Shared-memory variant of the pass above: the direct kernel issues ~70 loads per cell through L1
// (the dissipation term averages every off-diagonal derivative over four corners); here a block
// marches a TX x TY column along z with the planes z-1, z, z+1 of (u_x, u_y, u_z, temp) and a
// one-cell halo in a 4-slot ring (next plane prefetched into registers during the arithmetic), so
// every value comes from HBM/L2 once per block column.
// TAB: the slices of the six metric tables the block needs sit in shared memory too (one LDS with an immediate
// offset per entry instead of a parameter load + 64-bit address + LDG: the pass is issue-bound).
template <typename T, typename CT, int TX, int TY, int ZC, bool PREF, bool TAB>
__global__ void __launch_bounds__(TX *TY, sizeof(T) == 4 ? 4 : 2) bouss_tile_kernel(Grid<T> g, StageArgs<T> a) {
    constexpr int PX = TX + 2, PY = TY + 2, PLANE = PX * PY, NT = TX * TY, NE = 4 * PLANE;
    constexpr int NL = (NE + NT - 1) / NT;
    constexpr int LX = TX + 2, LY = TY + 2, LZ = ZC + 2, LEN = LX + LY + LZ;  // table slices: x, y, z back to back
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s = reinterpret_cast<T *>(smem_raw);  // [4 slots][u_x, u_y, u_z, temp][PY][PX]
    T *mt = s + 4 * NE;                      // [rdu, rdl, rdlg, rdug, wa, wb][LEN]
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int x0 = 1 + blockIdx.x * TX, y0 = 1 + blockIdx.y * TY;
    const int zbeg = 1 + blockIdx.z * ZC;
    const int zend = min(zbeg + ZC, g.N[2] - 1);
    if (zbeg >= zend) return;
    if (TAB) {  // entry r of a slice <-> index base - 1 + r of the dimension (clamped; clamped entries are never used)
        for (int e = tid; e < 6 * LEN; e += NT) {
            const int tb = e / LEN, r = e - tb * LEN;
            const int d = r < LX ? 0 : r < LX + LY ? 1 : 2;
            const int rl = d == 0 ? r : d == 1 ? r - LX : r - LX - LY;
            const int gi = min((d == 0 ? x0 : d == 1 ? y0 : zbeg) - 1 + rl, g.N[d] - 1);
            const T *tp = tb == 0 ? g.rdu[d] : tb == 1 ? g.rdl[d] : tb == 2 ? g.rdlg[d] : tb == 3 ? g.rdug[d] : tb == 4 ? g.wa[d] : g.wb[d];
            mt[e] = __ldg(tp + gi);
        }
    }
    const T *lp[NL];
    bool lv[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        int e = min(tid + i * NT, NE - 1);
        int c = e / PLANE, r = e - c * PLANE;
        int py = r / PX, px = r - py * PX;
        int gx = min(x0 - 1 + px, g.N[0] - 1), gy = min(y0 - 1 + py, g.N[1] - 1);
        lv[i] = tid + i * NT < NE;
        const T *base = c < 3 ? a.u + c * g.cs : a.temp;
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
    __syncthreads();
    const int lidx = (ty + 1) * PX + tx + 1;
    long long idx = Ic0 + (long long)Ic1 * g.s[1] + (long long)zbeg * g.s[2];
    for (int z = zbeg; z < zend; z++, idx += zs) {
        const bool more = z + 1 < zend;
        if (more) load_plane();
        if (active) {
            const int I[3] = {Ic0, Ic1, z};
            TileLoad<T, PX, PLANE> U, Tm;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                U.pl[k] = s + ((z - 1 + k) & 3) * NE + lidx;
                Tm.pl[k] = U.pl[k] + 3 * PLANE;
            }
            // PREF: the old values of the combine are requested before the arithmetic (the warps of a block march
            // in lock-step behind one barrier per plane, so a load issued after it is exposed DRAM latency)
            const bool grav_ok = a.gdir >= 0 && a.gdir < 3;
            const long long qg = idx + (long long)(grav_ok ? a.gdir : 0) * g.cs;
            T o_tb = T(0), o_ta = T(0), o_v = T(0), o_acc = T(0);
            if (PREF) {
                if (a.tv) o_tb = a.tvbase[idx];
                if (a.tacc_out) o_ta = a.tacc_in[idx];
                if (grav_ok && a.v) o_v = a.v[qg];
                if (grav_ok && a.acc_out) o_acc = a.acc_out[qg];
            }
            // entries of this thread's cell in the table slices (x: tx + 1, y: ty + 1, z: z - zbeg + 1)
            SmemTab<T, LEN> MT;
            MT.p[0] = mt + tx + 1, MT.p[1] = mt + LX + ty + 1, MT.p[2] = mt + LX + LY + (z - zbeg) + 1;
            const T kt = TAB ? temperature_rhs_m<T, 3>(U, Tm, MT, (T)a.kappa, (T)a.nu, a.dodiss != 0)
                             : temperature_rhs<T, 3>(U, Tm, g, I, (T)a.kappa, (T)a.nu, a.dodiss != 0);
            if (a.tk_out) a.tk_out[idx] = kt;
            if (a.tv) {
                T x = PREF ? o_tb : a.tvbase[idx];
                for (int j = 0; j < a.nprev; j++) x = axpy_rn<T, CT>(x, a.cprev[j], a.tkprev[j][idx]);
                a.tv[idx] = axpy_rn<T, CT>(x, a.cv, kt);
            }
            if (a.tacc_out) a.tacc_out[idx] = axpy_rn<T, CT>(PREF ? o_ta : a.tacc_in[idx], a.ca, kt);
            if (grav_ok) {
                const int c = a.gdir;
                T av;
                if (TAB)
                    av = c == 0 ? avg_face_m<3, T>(Tm, MT, 0, 0) : c == 1 ? avg_face_m<3, T>(Tm, MT, 1, 0) : avg_face_m<3, T>(Tm, MT, 2, 0);
                else
                    av = c == 0 ? avg_face<3>(Tm, g, I, 0, 0) : c == 1 ? avg_face<3>(Tm, g, I, 1, 0) : avg_face<3>(Tm, g, I, 2, 0);
                const T gk = (T)a.grav * av;
                if (a.k_out) a.k_out[qg] += gk;
                if (a.v) a.v[qg] = axpy_rn<T, CT>(PREF ? o_v : a.v[qg], a.cv, gk);
                if (a.acc_out) a.acc_out[qg] = axpy_rn<T, CT>(PREF ? o_acc : a.acc_out[qg], a.ca, gk);
            }
        }
        if (more) store_plane(z + 2);
        __syncthreads();
    }
}

template <typename T, typename CT>
static int launch_stage_bouss_split(const Grid<T> &g, const StageArgs<T> &a, Launch L) {
    StageArgs<T> m = a;  // momentum without the temperature coupling
    m.bouss = 0;
    m.temp = nullptr;
    m.tv = m.tacc_out = m.tk_out = nullptr;
    INSB_TRY((launch_stage_pair<T, CT>(g, m, L)));
    // measured at 512x512x256 (cfg 4), rhs + combine per RK44 step: Float32 direct 11.1 -> tile 8.96 -> tile with the
    // operand prefetch 7.17 ms; Float64 direct 15.17, tile 17.64, tile with the prefetch 14.57 ms.  INSB200_BOUSS_TILE=0:
    // the one-thread-per-cell pass
    const char *te = getenv("INSB200_BOUSS_TILE");
    if (!te || atoi(te) != 0) {
        constexpr int TX = 32, TY = 8, ZC = 64;
        constexpr size_t smem_f = (size_t)4 * 4 * (TX + 2) * (TY + 2) * sizeof(T);
        constexpr size_t smem_t = smem_f + (size_t)6 * (TX + 2 + TY + 2 + ZC + 2) * sizeof(T);
        dim3 b(TX, TY, 1);
        dim3 gr((g.N[0] - 2 + TX - 1) / TX, (g.N[1] - 2 + TY - 1) / TY, (g.N[2] - 2 + ZC - 1) / ZC);
        const char *pe = getenv("INSB200_BOUSS_PREFETCH"), *se = getenv("INSB200_BOUSS_SMEMTAB");
        const bool pref = !pe || atoi(pe) != 0, tab = !se || atoi(se) != 0;
        if (tab && pref) {
            auto kern = bouss_tile_kernel<T, CT, TX, TY, ZC, true, true>;
            if (smem_t > 48 * 1024) INSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
            kern<<<gr, b, smem_t, L.stream>>>(g, a);
        } else if (pref) {
            bouss_tile_kernel<T, CT, TX, TY, ZC, true, false><<<gr, b, smem_f, L.stream>>>(g, a);
        } else {
            bouss_tile_kernel<T, CT, TX, TY, ZC, false, false><<<gr, b, smem_f, L.stream>>>(g, a);
        }
        if (L.counter) ++*L.counter;
        INSB_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 b(64, 4, 1);
    dim3 gr((g.N[0] - 2 + b.x - 1) / b.x, (g.N[1] - 2 + b.y - 1) / b.y, g.N[2] - 2);
    bouss_extra_kernel<T, CT><<<gr, b, 0, L.stream>>>(g, a);
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
    const bool no_split = getenv("INSB200_BOUSS_DIRECT") != nullptr;
    if (g.D == 3 && a.bouss && a.temp && a.conv && a.diff && !a.zero_outside && !a.add_to_k && !force_direct &&
        !no_split && a.v != nullptr && 3 * g.cs < (1LL << 32)) {
        if (sizeof(T) == 8 || ct_double) return launch_stage_bouss_split<T, double>(g, a, L);
        return launch_stage_bouss_split<T, T>(g, a, L);
    }
    if (g.D == 3 && !a.bouss && a.conv && a.diff && !a.zero_outside && !a.add_to_k && !force_direct &&
        a.v != nullptr && 3 * g.cs < (1LL << 32)) {
        if (sizeof(T) == 8 || ct_double) return launch_stage_pair<T, double>(g, a, L);
        return launch_stage_pair<T, T>(g, a, L);
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
// Each thread handles RB consecutive rows (y) of one x column: all loads are issued before the
// first store so that RB x (D+D) independent requests are in flight per thread.
constexpr int RB = 4;

// wrap: bit d set = dimension d is periodic and the lower neighbour of the first interior cell is
// taken from the last interior cell instead of the ghost cell (which then need not be filled).
template <typename T, int D>
__global__ void __launch_bounds__(128) divergence_kernel(Grid<T> g, const T *__restrict__ u,
                                                         T *__restrict__ div, int scale, int wrap) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int iy0 = blockIdx.y * RB + 1;
    const int iz = (D == 3) ? (int)blockIdx.z + 1 : 0;
    if (ix >= g.N[0] - 1) return;
    const long long s1 = g.s[1], s2 = g.s[2], cs = g.cs;
    const long long base = ix + iy0 * s1 + (D == 3 ? iz * s2 : 0);
    const int ny = g.N[1] - 1;
    const long long xm = ((wrap & 1) && ix == 1) ? g.N[0] - 3 : -1;                      // offset to ix-1
    const long long zm = ((wrap & 4) && iz == 1) ? (long long)(g.N[2] - 3) * s2 : -s2;   // offset to iz-1
    T ux[RB], uxm[RB], uy[RB + 1], uz[RB], uzm[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const bool ok = iy0 + r < ny;
        const long long q = base + r * s1;
        ux[r] = ok ? __ldg(u + q) : T(0);
        uxm[r] = ok ? __ldg(u + q + xm) : T(0);
        uy[r + 1] = ok ? __ldg(u + cs + q) : T(0);
        if (D == 3) {
            uz[r] = ok ? __ldg(u + 2 * cs + q) : T(0);
            uzm[r] = ok ? __ldg(u + 2 * cs + q + zm) : T(0);
        }
    }
    uy[0] = __ldg(u + cs + base + (((wrap & 2) && iy0 == 1) ? (long long)(g.N[1] - 3) * s1 : -s1));
    const T rx = __ldg(g.rdlg[0] + ix), dx = __ldg(g.dl[0] + ix);
    T rz = T(0), dz = T(1);
    if (D == 3) {
        rz = __ldg(g.rdlg[2] + iz);
        dz = __ldg(g.dl[2] + iz);
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (iy0 + r >= ny) break;
        const int iy = iy0 + r;
        T acc = (ux[r] - uxm[r]) * rx + (uy[r + 1] - uy[r]) * __ldg(g.rdlg[1] + iy);
        if (D == 3) acc = acc + (uz[r] - uzm[r]) * rz;
        if (scale) {
            T om = dx * __ldg(g.dl[1] + iy);
            if (D == 3) om = om * dz;
            acc *= om;
        }
        div[base + r * s1] = acc;
    }
}

template <typename T>
int launch_divergence(const Grid<T> &g, const T *u, T *div, bool scale, Launch L, int wrap) {
    dim3 b(128, 1, 1);
    dim3 gr((g.N[0] - 2 + b.x - 1) / b.x, (g.N[1] - 2 + RB - 1) / RB, g.D == 3 ? g.N[2] - 2 : 1);
    if (g.D == 3)
        divergence_kernel<T, 3><<<gr, b, 0, L.stream>>>(g, u, div, scale, wrap);
    else
        divergence_kernel<T, 2><<<gr, b, 0, L.stream>>>(g, u, div, scale, wrap);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_divergence<float>(const Grid<float> &, const float *, float *, bool, Launch, int);
template int launch_divergence<double>(const Grid<double> &, const double *, double *, bool, Launch, int);

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
// wrap: bit d set = dimension d is periodic: the upper neighbour of the last interior pressure is the
// first interior one (p ghosts need not be filled), and every updated velocity is also stored
// to its periodic ghost images (what apply_bc_u! would copy afterwards,
// src/boundary_conditions.jl:273-285, dims in order => edges and corners included).
template <typename T, int D>
__global__ void __launch_bounds__(128) pgrad_kernel(Grid<T> g, const T *__restrict__ p, T *G, T *usub, int wrap) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int iy0 = blockIdx.y * RB + 1;
    const int iz = (D == 3) ? (int)blockIdx.z + 1 : 0;
    if (ix >= g.N[0] - 1) return;
    const long long s1 = g.s[1], s2 = g.s[2], cs = g.cs;
    const long long base = ix + iy0 * s1 + (D == 3 ? iz * s2 : 0);
    const int ny = g.N[1] - 1;
    const long long xp = ((wrap & 1) && ix == g.N[0] - 2) ? -(long long)(g.N[0] - 3) : 1;
    const long long zp = ((wrap & 4) && iz == g.N[2] - 2) ? -(long long)(g.N[2] - 3) * s2 : s2;
    T pc[RB + 1], px[RB], pz[RB], uo[D][RB];
#pragma unroll
    for (int r = 0; r < RB; r++) {
        const bool ok = iy0 + r < ny;
        const long long q = base + r * s1;
        pc[r] = ok ? __ldg(p + q) : T(0);
        px[r] = ok ? __ldg(p + q + xp) : T(0);
        if (D == 3) pz[r] = ok ? __ldg(p + q + zp) : T(0);
        if (usub) {
#pragma unroll
            for (int c = 0; c < D; c++) uo[c][r] = ok ? usub[q + c * cs] : T(0);
        }
    }
    {
        int last = min(iy0 + RB, ny);  // row above the last one handled here (<= N1-1, in bounds)
        const int lastr = last;
        if ((wrap & 2) && last == ny) last = 1;
        const T top = __ldg(p + ix + last * s1 + (D == 3 ? iz * s2 : 0));
        pc[RB] = T(0);
#pragma unroll
        for (int r = 1; r <= RB; r++)
            if (iy0 + r == lastr) pc[r] = top;
    }
    const T rx = __ldg(g.rdu[0] + ix);
    const T rz = D == 3 ? __ldg(g.rdu[2] + iz) : T(0);
    // periodic ghost images of this column: offsets (0 = none)
    const long long gx = !(wrap & 1) ? 0 : ix == 1 ? g.N[0] - 2 : ix == g.N[0] - 2 ? -(long long)(g.N[0] - 2) : 0;
    const long long gz = !(D == 3 && (wrap & 4)) ? 0
                         : iz == 1 ? (long long)(g.N[2] - 2) * s2 : iz == g.N[2] - 2 ? -(long long)(g.N[2] - 2) * s2 : 0;
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (iy0 + r >= ny) break;
        const int iy = iy0 + r;
        const long long q = base + r * s1;
        T d[3];
        d[0] = (px[r] - pc[r]) * rx;
        d[1] = (pc[r + 1] - pc[r]) * __ldg(g.rdu[1] + iy);
        if (D == 3) d[2] = (pz[r] - pc[r]) * rz;
        const long long gy = !(wrap & 2) ? 0 : iy == 1 ? (long long)(g.N[1] - 2) * s1
                                               : iy == g.N[1] - 2 ? -(long long)(g.N[1] - 2) * s1 : 0;
#pragma unroll
        for (int c = 0; c < D; c++) {
            if (G) G[q + c * cs] = d[c];
            if (usub) {
                const T v = uo[c][r] - d[c];
                T *o = usub + q + c * cs;
                o[0] = v;
                if (wrap) {
                    if (gx) o[gx] = v;
                    if (gy) {
                        o[gy] = v;
                        if (gx) o[gy + gx] = v;
                    }
                    if (gz) {
                        o[gz] = v;
                        if (gx) o[gz + gx] = v;
                        if (gy) {
                            o[gz + gy] = v;
                            if (gx) o[gz + gy + gx] = v;
                        }
                    }
                }
            }
        }
    }
}

template <typename T>
int launch_pressuregradient(const Grid<T> &g, const T *p, T *G, T *u_sub, Launch L, int wrap) {
    dim3 b(128, 1, 1);
    dim3 gr((g.N[0] - 2 + b.x - 1) / b.x, (g.N[1] - 2 + RB - 1) / RB, g.D == 3 ? g.N[2] - 2 : 1);
    if (g.D == 3)
        pgrad_kernel<T, 3><<<gr, b, 0, L.stream>>>(g, p, G, u_sub, wrap);
    else
        pgrad_kernel<T, 2><<<gr, b, 0, L.stream>>>(g, p, G, u_sub, wrap);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}
template int launch_pressuregradient<float>(const Grid<float> &, const float *, float *, float *, Launch, int);
template int launch_pressuregradient<double>(const Grid<double> &, const double *, double *, double *,
                                             Launch, int);

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
