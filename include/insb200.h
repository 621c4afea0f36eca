/* insb200.h -- C ABI of libinsb200.so, the B200 (sm_100a) backend for the
 * per-time-step hot path of IncompressibleNavierStokes.jl v5.0.1.
 *
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference checkout).  A Julia package extension binds these with
 * `ccall`; the Python/ctypes harness in
 * incompressiblenavierstokes.jl_b200/insb200/ binds the very same symbols.
 *
 * Conventions
 *  - All field pointers are DEVICE pointers owned by the caller, laid out exactly
 *    like the reference's arrays: column-major, one ghost volume per side,
 *    scalar `N[0] x N[1] (x N[2])`, vector `N... x D` with the component as the
 *    slowest index (src/initializers.jl:2,5; src/grid.jl:101-112).
 *  - `stream` is a `cudaStream_t` passed as `void*`.  Calls enqueue work and
 *    return without synchronising, except where stated.
 *  - Return value: 0 on success, non-zero on failure; `insb200_last_error()`
 *    then returns a message.  No C++ exception crosses this boundary.
 *  - A plan is not re-entrant; distinct plans may be used from distinct threads.
 *  - There is NO CPU fallback: every function fails if no CUDA device is present.
 */
#ifndef INSB200_H
#define INSB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define INSB200_VERSION 100 /* 0.1.0 */

enum { INSB200_F32 = 0, INSB200_F64 = 1 };

/* src/boundary_conditions.jl:5-39 */
enum {
    INSB200_BC_PERIODIC = 0,
    INSB200_BC_DIRICHLET = 1,
    INSB200_BC_SYMMETRIC = 2,
    INSB200_BC_PRESSURE = 3
};

/* src/pressure.jl: psolver_spectral :321, psolver_transform :215, psolver_tridiagonal :401 */
enum {
    INSB200_PSOLVER_NONE = -1,
    INSB200_PSOLVER_SPECTRAL = 0,
    INSB200_PSOLVER_TRANSFORM = 1,
    INSB200_PSOLVER_TRIDIAGONAL = 2,
    INSB200_PSOLVER_DEFAULT = 3, /* default_psolver, src/pressure.jl:58-78 */
    INSB200_PSOLVER_CG = 4       /* psolver_cg, src/pressure.jl:155-203: any boundary conditions / grids.
                                  * default_psolver resolves to it where the reference would build the
                                  * sparse direct solver (psolver_direct is not on the GPU path).  In that
                                  * case -- and unless cg_reltol / cg_abstol are given -- the relative
                                  * tolerance is 1e-13 (Float64) / 1e-5 (Float32) instead of psolver_cg's
                                  * sqrt(eps), and the iteration stops when the residual has not fallen for
                                  * 200 iterations, so that the substitute solves to round-off level like the
                                  * direct solver; insb200_plan_psolver() tells which solver was chosen */
};

/* The subset of the `Setup` NamedTuple (src/grid.jl:91-208) the hot path consumes.
 * The 1-D metric tables are DATA computed by the caller exactly as the reference
 * does; the library never re-derives grid spacings. */
typedef struct insb200_setup {
    int D;              /* 2 or 3 */
    int N[3];           /* volumes per dim INCLUDING ghosts (setup.N) */
    int dtype;          /* INSB200_F32 / INSB200_F64 = eltype(x[1]) */
    int bc_u[3][2];     /* boundary_conditions.u[dim][left/right] */
    int has_temp;       /* boundary_conditions has a `temp` entry */
    int bc_temp[3][2];  /* boundary_conditions.temp[dim][left/right] */
    /* HOST pointers to 1-D tables of length N[dim], element type `dtype`:
     * setup.Δ, setup.Δu, setup.A_stag[dim][1], setup.A_stag[dim][2] */
    const void *Delta[3];
    const void *DeltaU[3];
    const void *A1[3];
    const void *A2[3];
    double xlims[3][2];             /* setup.xlims (psolver_transform uses it, pressure.jl:221-223) */
    double bc_u_value[3][2][3];     /* DirichletBC constants [dim][side][component]; 0 = no-slip */
    double bc_temp_value[3][2];     /* DirichletBC temperature constants [dim][side] */
    int psolver;                    /* INSB200_PSOLVER_* */
    int tridiagonal_dir;            /* wall-normal dir for psolver_tridiagonal, -1 = infer (pressure.jl:410-416) */
    /* Optional HOST eigenvalue tables (`ahat`, pressure.jl:234-246,336-342; `lam`, :453-460),
     * element type `dtype`; NULL = computed by the library the way the reference does. */
    const void *eig[3];
    int device;                     /* CUDA device ordinal, -1 = current device */
    /* z-slab decomposition (one process per GPU over NCCL; no counterpart in the reference,
     * SURVEY.md §8e).  world == 1 for a single GPU.  With world > 1 the setup still describes
     * the GLOBAL grid; every field pointer passed to the library is then this rank's slab:
     * N[0] x N[1] x ((N[2]-2)/world + 2) per component, i.e. the planes
     * [rank*nzl, rank*nzl + nzl + 2) of the global ghosted array.  Requires D == 3, periodic z,
     * (N[2]-2) % world == 0; the Poisson solve then needs psolver_spectral on power-of-two grids. */
    int rank, world;
    /* 128-byte NCCL unique id from insb200_comm_unique_id(), identical on all ranks (world > 1). */
    const void *nccl_id;
    /* psolver_cg keyword arguments (pressure.jl:157-161); 0 = the reference defaults
     * (abstol 0, reltol sqrt(eps(T)), maxiter prod(Np)) */
    double cg_abstol, cg_reltol;
    long long cg_maxiter;
} insb200_setup_t;

/* `params` NamedTuple forwarded to force! (src/solver.jl:2-35,176) */
typedef struct insb200_params {
    double viscosity;
    double conductivity;
    double gravity;
    int gdir;           /* 0-based */
    int dodissipation;
    int boussinesq;     /* 0: navierstokes! (solver.jl:2-6); 1: boussinesq! (solver.jl:16-35) */
    double bodyforce[3];/* constant body force added to k.u (examples/ChannelFlow.jl:172-176); 0 = none */
    /* eddy-viscosity closure added to k.u, i.e. the force_eddy! of examples/ChannelFlow.jl:177-182:
     * navierstokes! + eddy_viscosity_closure!(model, ...) (src/eddyviscosity.jl:356-378), 3-D only */
    int eddy_model;     /* 0 none, 1 Smagorinsky, 2 WALE, 3 Vreman, 4 QR (INSB200_EDDY_*) */
    double eddy_C;      /* model constant (Smagorinsky(C), WALE(C), ...) */
} insb200_params_t;
#define INSB200_EDDY_NONE 0
#define INSB200_EDDY_SMAGORINSKY 1
#define INSB200_EDDY_WALE 2
#define INSB200_EDDY_VREMAN 3
#define INSB200_EDDY_QR 4

typedef struct insb200_plan insb200_plan_t;

const char *insb200_last_error(void);
int insb200_version(void);

/* Plan = device copies of the Setup tables + psolver workspace (replaces the closures
 * built by psolver_spectral/transform/tridiagonal, pressure.jl:344-347,248-253,469-476)
 * + time-stepper caches (get_cache, time_steppers/time_stepper_caches.jl:11-24). */
int insb200_plan_create(const insb200_setup_t *setup, insb200_plan_t **out);
/* Rank 0 calls this and ships the 128 bytes to the other ranks (MPI.jl / Distributed in Julia,
 * torch.distributed in the Python harness) before every rank calls insb200_plan_create. */
int insb200_comm_unique_id(void *id128);
/* Direct peer-store transposes of the distributed FFT (NVLink P2P through CUDA IPC; optional,
 * the NCCL all-to-all is used otherwise): every rank exports a 192-byte blob, the host gathers
 * the blobs of all ranks in rank order (world x 192 bytes) and every rank imports them.  The
 * strided FFT passes then store each tile straight into the destination rank's buffer and
 * ranks synchronise through flags in peer memory. */
int insb200_plan_ipc_export(insb200_plan_t *plan, void *blob192);
int insb200_plan_ipc_import(insb200_plan_t *plan, const void *blobs);
/* Collective tear-down of decomposed plans: every rank first unmaps its peers' buffers (insb200_plan_ipc_close),
 * all ranks synchronise (a barrier of the caller's launcher), and only then the plans are destroyed -- memory that
 * another process still has mapped must not be freed (CUDA IPC).  After the call the plan keeps working over NCCL. */
int insb200_plan_ipc_close(insb200_plan_t *plan);
/* Extents (including ghosts) of the arrays this rank passes in: N of the setup, with the slab
 * extent in z when world > 1. */
int insb200_plan_local_n(const insb200_plan_t *plan, int N_local[3]);
int insb200_plan_destroy(insb200_plan_t *plan);
/* Which psolver the plan holds (default_psolver resolved), and bytes of device workspace. */
int insb200_plan_psolver(const insb200_plan_t *plan);
long long insb200_plan_workspace_bytes(const insb200_plan_t *plan);

/* apply_bc_u! / apply_bc_p! / apply_bc_temp! (src/boundary_conditions.jl:163-171,200-208,237-245).
 * dudt != 0 selects the `dudt = true` variant (:341-357). */
int insb200_apply_bc_u(insb200_plan_t *plan, void *u, double t, int dudt, void *stream);
int insb200_apply_bc_p(insb200_plan_t *plan, void *p, void *stream);
int insb200_apply_bc_temp(insb200_plan_t *plan, void *temp, double t, void *stream);

/* Time-dependent (function-valued) DirichletBC, src/boundary_conditions.jl:344-357: the host
 * evaluates the function on the plane coordinates and hands the plane (device pointer, full
 * extent of the other dims, column-major) for (dim, side, component); NULL restores the constant. */
int insb200_set_bc_u_plane(insb200_plan_t *plan, int dim, int side, int comp, const void *plane);

/* navierstokes! / boussinesq! (src/solver.jl:2-35): k.u (and k.temp) are OVERWRITTEN on
 * 2:N-1 and zeroed elsewhere, like fill!(k,0) followed by convectiondiffusion!
 * (src/operators.jl:616-628) [+ applygravity! :802, convection_diffusion_temp! :669, dissipation! :772]. */
int insb200_rhs(insb200_plan_t *plan, const void *u, const void *temp, void *ku, void *ktemp,
                const insb200_params_t *params, void *stream);
/* convection! / diffusion! (src/operators.jl:407-416,545-554): ADD to f like the reference. */
int insb200_convection(insb200_plan_t *plan, void *f, const void *u, void *stream);
int insb200_diffusion(insb200_plan_t *plan, void *f, const void *u, double viscosity, void *stream);

/* divergence! (src/operators.jl:242-247); scalewithvolume! (:218-231);
 * pressuregradient! (:277-280); applypressure! (:306-311). */
int insb200_divergence(insb200_plan_t *plan, void *div, const void *u, void *stream);
int insb200_scalewithvolume(insb200_plan_t *plan, void *p, void *stream);
int insb200_pressuregradient(insb200_plan_t *plan, void *G, const void *p, void *stream);
int insb200_applypressure(insb200_plan_t *plan, void *u, const void *p, void *stream);

/* poisson!(psolver, p) = psolver(p) (src/pressure.jl:22): in place on the ghosted array,
 * touches only view(p, Ip) (pressure.jl:349-381, 255-317, 507-544). */
int insb200_psolve(insb200_plan_t *plan, void *p, void *stream);
/* Iterations the last psolver_cg solve of this plan took (0 for the FFT-type solvers). */
int insb200_psolver_iterations(const insb200_plan_t *plan);
/* project!(u, setup; psolver, p) (src/pressure.jl:42-55). p is caller scratch (ghosted scalar). */
int insb200_project(insb200_plan_t *plan, void *u, void *p, void *stream);

/* The stage of a Runge-Kutta step composed on the host -- user-defined force! functions (src/solver.jl:55-61)
 * or Dirichlet functions that depend on time, whose planes the host re-evaluates at every stage time:
 *   copyto!(u, u_start); for j: @. u += coef[j] * k[j]      (step_explicit_runge_kutta.jl:27-32,
 *                                                            step_lmwray3.jl:41-50: one k, coef = a[i] dt)
 *   apply_bc_u!(u, t); project!(u, setup; psolver, p); apply_bc_u!(u, t); apply_bc_temp!(temp, t)   (:38-52)
 * coef[j] = dt * A[i, j] evaluated by the caller in promote(eltype(method), typeof(dt)); coef_dtype says which
 * type that was (the products and sums are rounded like the reference broadcast, no FMA contraction).
 * u may alias u_start.  temp / temp_start / ktemp may be NULL (no temperature).  p: ghosted scalar scratch.
 * insb200_stage_combine is the combine alone on `ncomp` scalar fields (D for a velocity, 1 for a scalar);
 * insb200_axpy is y += a * x (the statestart update of step_lmwray3.jl:86-90). */
int insb200_stage_combine(insb200_plan_t *plan, void *dst, const void *base, int nk, const void *const *k,
                          const double *coef, int coef_dtype, int ncomp, void *stream);
int insb200_stage_combine_project(insb200_plan_t *plan, void *u, const void *u_start, void *temp,
                                  const void *temp_start, int nk, const void *const *k, const void *const *ktemp,
                                  const double *coef, int coef_dtype, void *p, double t_stage, void *stream);
int insb200_axpy(insb200_plan_t *plan, void *y, double a, const void *x, int coef_dtype, int ncomp, void *stream);

/* timestep!(::ExplicitRungeKuttaMethod, ...) (src/time_steppers/step_explicit_runge_kutta.jl:4-57).
 * A: s*s row-major SHIFTED tableau (methods.jl:80-85), c: s stage times, both as double;
 * coef_dtype = eltype of the method (INSB200_F32/F64): the stage combine is evaluated in
 * promote(field dtype, coef_dtype) like the reference broadcast (:28-32).
 * u (and temp) are advanced in place from t to t + dt. */
int insb200_step_rk(insb200_plan_t *plan, const double *A, const double *c, int s, int coef_dtype,
                    void *u, void *temp, double t, double dt, const insb200_params_t *params,
                    void *stream);
/* timestep!(::LMWray3, ...) (src/time_steppers/step_lmwray3.jl:3-94). */
int insb200_step_lmwray3(insb200_plan_t *plan, void *u, void *temp, double t, double dt,
                         const insb200_params_t *params, void *stream);
/* Same as insb200_step_rk but with HOST buffers: H2D copy of u (and temp), nsteps steps,
 * D2H copy back, stream-synchronised on return (reference callers holding `Array`s). */
int insb200_step_rk_host(insb200_plan_t *plan, const double *A, const double *c, int s,
                         int coef_dtype, void *u_host, void *temp_host, double t, double dt,
                         int nsteps, const insb200_params_t *params, void *stream);

/* Ensemble variant of insb200_step_rk_host: `nmembers` independent states held in HOST buffers
 * (pinned memory for full copy bandwidth) are each advanced by nsteps steps.  Member m+1 is
 * copied in and member m-1 copied out (two extra streams, two device slots) while member m is
 * being stepped, so that the PCIe copies overlap the kernels; returns after the last copy.
 * Replaces a host-side loop over independent `Array` states, each advanced with timestep!
 * (src/time_steppers/step_explicit_runge_kutta.jl:4-57). */
int insb200_step_rk_host_ensemble(insb200_plan_t *plan, const double *A, const double *c, int s,
                                  int coef_dtype, void *const *u_hosts, void *const *temp_hosts,
                                  int nmembers, double t, double dt, int nsteps,
                                  const insb200_params_t *params, void *stream);

/* eddy_viscosity_closure!(model, f, u, cache, setup) (src/eddyviscosity.jl:356-378): f -= div of the
 * modelled stress; u with its ghost cells filled, f and u distinct arrays.  The cache is plan-owned scratch:
 * one scalar field (visc) on the default path, which recomputes the gradient tensor and the stress from u
 * instead of storing them; the nine tensor components G only with PressureBC sides or on slabs. */
int insb200_eddy_viscosity_closure(insb200_plan_t *plan, int model, double C, void *f, const void *u,
                                   void *stream);

/* ---- diagnostics feeding the processors (src/operators.jl:853-1066) -----------------------
 * One pass over the fields each; outputs have the layout of the reference's arrays (scalar or
 * vector field, ghosted), entries outside the reference's index range are left untouched.
 * vorticity! (:853-881): omega is a scalar field in 2-D, a vector field in 3-D. */
int insb200_vorticity(insb200_plan_t *plan, void *omega, const void *u, void *stream);
/* interpolate_u_p! (:921-940): velocity at the pressure points. */
int insb200_interpolate_u_p(insb200_plan_t *plan, void *up, const void *u, void *stream);
/* interpolate_ω_p! (:950-984): vorticity at the pressure points. */
int insb200_interpolate_omega_p(insb200_plan_t *plan, void *omegap, const void *omega, void *stream);
/* qcrit! (:989-1000): Q = -tr(G G)/2 of the collocated velocity gradient. */
int insb200_qcrit(insb200_plan_t *plan, void *q, const void *u, void *stream);
/* kinetic_energy! (:1027-1055), interpolate_first = 0/1. */
int insb200_kinetic_energy(insb200_plan_t *plan, void *ke, const void *u, int interpolate_first,
                           void *stream);
/* total_kinetic_energy (:1061-1066): sum over the pressure volumes of ke * cell volume, evaluated
 * without materialising ke (summed in double, deterministic order; all-reduced over the ranks of a
 * decomposed plan).  Synchronises the stream and returns the host scalar. */
int insb200_total_kinetic_energy(insb200_plan_t *plan, const void *u, int interpolate_first,
                                 double *out, void *stream);

/* random_field(setup, t; profile, totalenergy, rng) (src/initializers.jl:109-183) on the device, with the FFT
 * passes of the plan's psolver_spectral: white noise per component -> projection onto the divergence-free space of
 * the staggered grid (:133-151) -> shell-wise rescaling to the target spectrum (:153-172, shells of
 * src/spectral.jl:39-67) -> view(u, Ip, a), then apply_bc_u!.  shell_energy[kappa] (host, kdiag + 1 entries,
 * kdiag = isqrt(sum (Np/2)^2)) = totalenergy * profile(kappa) / sum(profile): the caller evaluates the profile
 * function, which cannot cross the ABI.  fill_noise == 0: the interior of u holds the caller's N(0, 1) samples
 * (the reference draws them with its own rng); != 0: they are generated here (Philox4x32-10 keyed with `seed`,
 * Box-Muller), so that fields too large for the host never exist there. */
int insb200_random_field(insb200_plan_t *plan, void *u, const double *shell_energy, int nshell, int fill_noise,
                         unsigned long long seed, double t, void *stream);

/* energyspectrum(u, setup) (src/spectral.jl:75-117) on a plan holding psolver_spectral (periodic,
 * power-of-two extents, single GPU): the solver's FFT passes transform every component, the modes
 * are binned into the shells kappa = 0..kmax (kmax < 0: minimum(Np) / 2, spectral_stuff :36); ehat
 * receives kmax + 1 host doubles whose sum is the mean kinetic energy density.  Synchronises. */
int insb200_energy_spectrum(insb200_plan_t *plan, const void *u, int kmax, double *ehat, void *stream);

/* The spectral sums turbulence_statistics(u, setup, viscosity) is made of (src/spectral.jl:144-257; same plans and
 * preconditions as above): moments[j] = sum over all wavenumbers (spectralsum :89-90, conjugate modes counted) of
 * |2 pi k / L|^(2 j) |u_hat(k)|^2 / 2, j = 0, 1, 2, with spectral_energy_field! (:75-87) and
 * squared_wavenumbers_dimensional (:121-135) evaluated on the fly behind the solver's FFT passes.  The caller forms
 * e = m0, diss = 2 nu m1, enstrophy_diss = 2 nu m2 and the derived scales.  3 host doubles; synchronises. */
int insb200_spectral_moments(insb200_plan_t *plan, const void *u, double *moments, void *stream);

/* solve_unsteady (src/solver.jl:77-149): the time loop itself, so that a run of many small steps
 * does not return to the host language per step.
 *   method 0: explicit Runge-Kutta with the tableau (A, c, s, coef_dtype) of insb200_step_rk;
 *   method 1: LMWray3 (A, c ignored).
 *   dt > 0: fixed step, rounded so that (tend - tstart) / dt is an integer (:123-132); with
 *     use_graph != 0 (single GPU) the step is captured into a CUDA graph once and replayed;
 *   dt <= 0: adaptive (:98-122): every n_adapt_dt steps dt = cfl * propose_timestep, raised to
 *     dt_min when dt_min > 0; a proposal < 1e-10, > 100 or NaN stops the loop with the status
 *     INSB200_STOPPED_BAD_DT (the reference warns and breaks); the last step is clipped to tend.
 *   time_dtype: INSB200_F32/F64, the type t and dt arithmetic is evaluated in (typeof(tlims)).
 *   callback: called after every step whose number n is a multiple of callback_every
 *     (0 = never) -- the reference's processors observe the state after every step and act every
 *     `nupdate` steps (src/processors.jl:45-80), so pass the gcd of their nupdate.  The state is
 *     ordered on `stream` when the callback runs; a non-zero return value ends the loop.
 * Outputs: final time and the number of steps taken. */
#define INSB200_STOPPED_BAD_DT 2
typedef int (*insb200_step_callback_t)(void *user, long long n, double t);
int insb200_solve_unsteady(insb200_plan_t *plan, int method, const double *A, const double *c, int s,
                           int coef_dtype, void *u, void *temp, double tstart, double tend, double dt,
                           double dt_min, double cfl, int n_adapt_dt, int time_dtype,
                           const insb200_params_t *params, int callback_every,
                           insb200_step_callback_t callback, void *user, int use_graph, double *t_out,
                           long long *n_out, void *stream);

/* right_hand_side!(dudt, u, params_ref, t) (src/sciml.jl:35-48). */
int insb200_right_hand_side(insb200_plan_t *plan, void *dudt, const void *u, double t,
                            const insb200_params_t *params, void *stream);

/* propose_timestep(force!, state, setup, params) (src/solver.jl:157-214). Synchronises the
 * stream and returns the proposed dt (WITHOUT the cfl factor) in *dt_out. */
int insb200_cfl(insb200_plan_t *plan, const void *u, const insb200_params_t *params,
                double *dt_out, void *stream);
/* maximum(abs, u) over the whole vector field (timelogger, src/processors.jl:75); syncs. */
int insb200_maxabs(insb200_plan_t *plan, const void *u, double *out, void *stream);

/* Debugging aid of the fused periodic time-step path (tools/debug_fused.py): copies one of the plan's
 * padded internal arrays to the host.  which: 0, 1 = unprojected stage velocities, 2 = running accumulator,
 * 3 = pressure, 4 = scaled divergence; dims receives {NPX, NPY, NPZ, components}; host may be NULL. */
int insb200_debug_fused_copy(insb200_plan_t *plan, int which, void *host, long long host_bytes, int dims[4]);

/* Number of kernel launches issued by this plan since creation (bench.py `gpu_launches`). */
long long insb200_launch_count(const insb200_plan_t *plan);

/* Device-timed kernel groups for roofline accounting: when enabled, CUDA events bracket the
 * stencil (A), Poisson (P) and projection (C) groups of every stage on `stream`. */
int insb200_profile_enable(insb200_plan_t *plan, int on);
/* ms accumulated per group [0]=rhs+combine, [1]=divergence, [2]=poisson, [3]=project, [4]=bc;
 * counts = number of brackets; resets the accumulators. Synchronises. */
int insb200_profile_read(insb200_plan_t *plan, double ms[5], long long counts[5]);
/* Same with the Poisson solve broken down: groups [5]=x_r2c, [6]=y_fwd, [7]=transpose_fwd,
 * [8]=z_mid (forward + eigenvalue divide + inverse), [9]=transpose_bwd, [10]=y_inv, [11]=x_c2r;
 * n <= 12 entries are written. */
int insb200_profile_read_ex(insb200_plan_t *plan, int n, double *ms, long long *counts);

#ifdef __cplusplus
}
#endif
#endif /* INSB200_H */
