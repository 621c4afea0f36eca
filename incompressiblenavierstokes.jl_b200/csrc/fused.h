// fused.h -- the fused periodic time-step path (stage_fused.cu): internal padded field layout,
// arguments of the fused stage kernel and of the kernels around it.
//
// Per Runge-Kutta stage the reference runs navierstokes! -> stage combine -> apply_bc_u! -> divergence!
// -> scalewithvolume! -> psolver -> apply_bc_p! -> applypressure! -> apply_bc_u!
// (src/time_steppers/step_explicit_runge_kutta.jl:24-46, src/pressure.jl:42-55).  On periodic boxes the
// fused path keeps the UNPROJECTED stage velocity v and the pressure p of a stage and lets the NEXT
// stage kernel form u = v - G p while it loads its planes ("lazy projection"), and it lets the stage
// kernel emit the scaled divergence of its own output.  Per stage the field traffic is
//   stage kernel: R v 3, R p 1, R u0 3, R acc 3, W v 3, W acc 3, W d 1    (17 w, stages 2..s-1)
//   Poisson solve: 5 passes x (R 1 + W 1)                                  (10 w)
// instead of 15 w + 4 w (divergence) + 10 w + 7 w (projection); only the last stage of a step is
// projected explicitly (project_out_kernel) into the caller's array.
#pragma once
#include "internal.h"

namespace insb {

// Internal layout of the plan-owned fields of the fused path.  Interior cell (i, j, k), 1-based like
// the reference's ghosted arrays, sits at (i + OX - 1, j + OY - 1, k + OZ - 1):
//   x: 4 pad/ghost cells before cell 1 (16-byte aligned rows and tiles -- what TMA and 128-bit
//      accesses need; the reference layout puts cell 1 of a Float32 row at an odd offset),
//   y, z: 2 ghost layers on either side (the fused divergence needs the stage output at I - e, hence
//      the velocity at I - 2e).
// Ghost layers are periodic images written by the producing kernels (x, y always; z on one GPU) or
// exchanged between neighbouring slabs (z, world > 1).
struct FusedLayout {
    static constexpr int OX = 4, OY = 2, OZ = 2;
    int n[3];            // interior extents; n[2] = planes of this slab
    int NPX, NPY, NPZ;   // padded extents
    long long s1, s2, cs;  // row, plane and component strides (elements)
    __host__ __device__ long long scalar_elems() const { return (long long)NPX * NPY * NPZ; }
    __host__ __device__ long long at(int i, int j, int k) const { return (i + OX - 1) + (long long)(j + OY - 1) * s1 + (long long)(k + OZ - 1) * s2; }
};

inline FusedLayout make_fused_layout(int n0, int n1, int nz) {
    FusedLayout L;
    L.n[0] = n0, L.n[1] = n1, L.n[2] = nz;
    L.NPX = (n0 + 8 + 3) / 4 * 4;
    L.NPY = n1 + 4;
    L.NPZ = nz + 4;
    L.s1 = L.NPX;
    L.s2 = (long long)L.NPX * L.NPY;
    L.cs = L.s2 * L.NPZ;
    return L;
}

struct FusedMaps;  // tensor maps of a set of internal buffers (opaque; stage_fused.cu)

// One fused stage.  Velocity input: `u_ref` (reference layout, ghosts filled, the first stage of a step)
// or (`v_in`, `p_in`) (internal layout, lazily projected).  vbase / acc_in: nullptr = the input velocity
// itself (first stage); vbase_ref = the caller's array (reference layout); vbase_int / acc_in internal.
template <typename T>
struct FusedStage {
    const T *u_ref = nullptr;
    const T *u_below = nullptr;  // slab mode, first stage: plane "-1" of u_ref, [3][N1][N0] (from the rank below)
    const T *v_in = nullptr, *p_in = nullptr;
    const T *vbase_ref = nullptr, *vbase_int = nullptr, *acc_in = nullptr;
    T *v_out = nullptr, *acc_out = nullptr, *d_out = nullptr;
    // chained steps (insb200_solve_unsteady): the first stage of a step that starts from the lazy state of the previous
    // step also stores the projected velocity u0 = v_in - G p_in (internal layout): the vbase of the later stages
    T *u0_out = nullptr;
    bool acc_halo = false;  // slab mode: a later stage reads acc_out as vbase_int (on its z halo too): exchange its ghost planes
    double cv = 0, ca = 0, nu = 0;
    double bodyforce[3] = {0, 0, 0};
};

struct FusedCtx {
    FusedLayout L;
    int zwrap;   // 1: z periodic on this GPU (ghost planes are images); 0: slab (ghost planes exchanged)
    int zoff;    // global index of local plane 0 minus 0 (rank * nzl)
    int n2g;     // global interior planes
    FusedMaps *maps;
};

int fused_maps_create(FusedMaps **out);
void fused_maps_destroy(FusedMaps *m);

template <typename T>
int launch_stage_fused(const Grid<T> &g, const FusedCtx &c, const FusedStage<T> &s, bool ct_double, Launch L);

// u_ref = v - G p with its periodic ghost images (x, y; z when zwrap): the explicit projection of the
// last stage, applypressure! + apply_bc_u! (src/operators.jl:306-311, src/boundary_conditions.jl:273-285).
template <typename T>
int launch_project_out(const Grid<T> &g, const FusedCtx &c, const T *v, const T *p, T *u_ref, Launch L);

// divergence! + scalewithvolume! of a reference-layout velocity into the internal layout (projection of
// a caller-provided field through the fused Poisson plumbing; also the reference path of the tests).
bool fused_supported(const Grid<float> &g);

}  // namespace insb
