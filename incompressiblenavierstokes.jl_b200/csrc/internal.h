// internal.h -- shared declarations of libinsb200 (host side + kernel launchers).
//
// Index conventions: 0-based; a scalar field has N[0]*N[1]*N[2] cells (N[2] = 1 in
// 2-D) with x fastest; a vector field is D such scalars back to back (the reference's
// column-major (N..., D) layout, src/initializers.jl:2,5).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/insb200.h"

namespace insb {

constexpr int MAXSTAGE = 8;

void set_error(const std::string &msg);

#define INSB_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            insb::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " at " \
                            + __FILE__ + ":" + std::to_string(__LINE__));               \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

#define INSB_TRY(call)          \
    do {                        \
        int r__ = (call);       \
        if (r__) return r__;    \
    } while (0)

// Device view of the Setup tables (all pointers are device pointers, length N[a]).
template <typename T>
struct Grid {
    int D;
    int N[3];
    long long s[3];  // strides of a scalar field: 1, N0, N0*N1
    long long cs;    // component stride N0*N1*N2
    const T *dl[3];    // Δ            (src/grid.jl:167-171)
    const T *du[3];    // Δu           (src/grid.jl:173-177)
    const T *rdl[3];   // 1/Δ
    const T *rdu[3];   // 1/Δu
    const T *rdlg[3];  // 1/Δ  where Δ  > 2 eps, else 0 (guard of δ, src/operators.jl:181-187)
    const T *rdug[3];  // 1/Δu where Δu > 2 eps, else 0
    const T *a1[3];    // A_stag[a][1] (src/grid.jl:180-186)
    const T *a2[3];    // A_stag[a][2]
    const T *wa[3];    // Δ[i+1]/(Δ[i]+Δ[i+1])  weights of avg (src/operators.jl:197-200)
    const T *wb[3];    // Δ[i]  /(Δ[i]+Δ[i+1])
};

// ---- comm.cu: z-slab decomposition, one process per GPU over NCCL ----
struct Comm;
int comm_unique_id(void *id128);
int comm_create(int rank, int world, const void *id128, Comm **out);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);   // 0 for nullptr
int comm_world(const Comm *c);  // 1 for nullptr
int comm_halo_z(Comm *c, void *field, size_t esz, long long plane, int nplanes, int ncomp, long long cs,
                cudaStream_t st, bool ring = true);
// Generic neighbour exchange along the decomposed direction: chunk `send_up` (+ c * src_cs_bytes for component
// c) goes to the rank above, which receives it in `recv_below` (+ c * dst_cs_bytes); `send_down` / `recv_above`
// likewise downwards.  Null pointers skip a direction (all ranks alike).
int comm_exchange(Comm *c, const void *send_up, const void *send_down, void *recv_below, void *recv_above,
                  size_t chunk_bytes, int ncomp, size_t src_cs_bytes, size_t dst_cs_bytes, cudaStream_t st);
int comm_exchange_launches(const Comm *c, size_t chunk_bytes, int ncomp);
// `ng` ghost planes on either side of `nplanes` contiguous planes per component
int comm_halo_planes(Comm *c, void *field, size_t esz, long long plane, int nplanes, int ng, int ncomp, long long cs,
                     cudaStream_t st, bool ring = true);
// Halo exchange by direct peer stores: `windows[r]` is rank r's window (peer-mapped, windows[rank]
// local) of comm_halo_window_bytes(slot_bytes) bytes, zero-initialised; slot_bytes >= bytes of
// the ghost planes of one exchange (all components).  Until enabled comm_halo_z uses NCCL.
size_t comm_halo_window_bytes(size_t slot_bytes);
int comm_enable_p2p_halo(Comm *c, void *const *windows, size_t slot_bytes);
void comm_disable_p2p_halo(Comm *c, const void *my_window);
int comm_halo_launches(const Comm *c);  // kernels of this library per halo exchange (0 on the NCCL path)
int comm_all_to_all(Comm *c, const void *send, void *recv, size_t block_bytes, cudaStream_t st);
int comm_reduce_scratch(Comm *c, double *scratch, bool is_min, cudaStream_t st);
int comm_sum_double(Comm *c, double *value, cudaStream_t st);  // in-place all-reduce(sum) of one device double
int comm_sum_doubles(Comm *c, double *values, int n, cudaStream_t st);
// cross-process flags in peer-mapped memory: every rank stores `epoch` into slot `rank` of the
// flag row `row` of all ranks (after everything enqueued before on `st`), resp. waits until all
// slots of its own row hold a value >= epoch.
int comm_signal_peers(unsigned *const *peer_flags, int row, int rank, int world, unsigned epoch, cudaStream_t st);
int comm_wait_flags(unsigned *my_flags, int row, int world, unsigned epoch, unsigned *timeout_flag, cudaStream_t st);

// Boundary description used by the ghost-fill kernels.
struct BcDesc {
    Comm *comm;  // non-null: dimension 2 is decomposed into slabs, its ghost planes come from the neighbours
    int bc_u[3][2];
    int bc_temp[3][2];
    double u_val[3][2][3];
    double temp_val[3][2];
    const void *u_plane[3][2][3];  // optional device planes for function-valued Dirichlet BC
};

// Arguments of the fused stage kernel (right-hand side + stage combine).
template <typename T>
struct StageArgs {
    const T *u;     // velocity the stencil reads (ghosts filled)
    const T *temp;  // temperature the stencil reads (Boussinesq) or nullptr
    // velocity outputs (any may be nullptr)
    T *v;            // v = vbase + sum_j cprev[j]*kprev[j] + cv*k
    const T *vbase;
    double cv;
    T *acc_out;      // acc_out = acc_in + ca*k
    const T *acc_in;
    double ca;
    T *k_out;        // k itself
    int zero_outside;  // k_out: also write zeros outside 2:N-1 (insb200_rhs semantics)
    int nprev;
    const T *kprev[MAXSTAGE];
    double cprev[MAXSTAGE];
    // temperature twins
    T *tv;
    const T *tvbase;
    T *tacc_out;
    const T *tacc_in;
    T *tk_out;
    const T *tkprev[MAXSTAGE];
    // parameters
    double nu, kappa, grav;
    int gdir, dodiss, bouss;
    double bodyforce[3];
    int eddy_model;  // eddy-viscosity closure added to k (0 = none), see les.cu
    double eddy_C;
    int conv, diff;  // which stresses to include (1,1 = convectiondiffusion!)
    int add_to_k;    // k_out += instead of = (convection!/diffusion! add semantics)
};

// Optional device-timing hook (CUDA events on the launching stream, see insb200_profile_*).
struct Profiler {
    virtual void begin(int group, cudaStream_t st) = 0;
    virtual void end(int group, cudaStream_t st) = 0;
    virtual ~Profiler() {}
};
constexpr int NPROF = 12;  // 0 rhs+combine 1 divergence 2 poisson 3 project 4 bc 5 x_r2c 6 y_fwd 7 transpose_fwd
                           // 8 z_mid 9 transpose_bwd 10 y_inv 11 x_c2r

struct Launch {
    cudaStream_t stream;
    long long *counter;
    Profiler *prof = nullptr;
};

struct ProfScope {
    const Launch &L;
    int group;
    ProfScope(const Launch &l, int g) : L(l), group(g) {
        if (L.prof) L.prof->begin(group, L.stream);
    }
    ~ProfScope() {
        if (L.prof) L.prof->end(group, L.stream);
    }
};

// ---- stencil.cu ----
template <typename T>
int launch_stage(const Grid<T> &g, const StageArgs<T> &a, bool ct_double, Launch L);
template <typename T>
int launch_divergence(const Grid<T> &g, const T *u, T *div, bool scale, Launch L, int wrap = 0);
template <typename T>
int launch_scalewithvolume(const Grid<T> &g, T *p, Launch L);
template <typename T>
int launch_pressuregradient(const Grid<T> &g, const T *p, T *G, T *u_sub, Launch L, int wrap = 0);
template <typename T>
int launch_cfl(const Grid<T> &g, const int Iu[3][3][2], const T *u, double *scratch, Launch L);
template <typename T>
int launch_maxabs(const T *u, long long n, double *scratch, Launch L);

// ---- diag.cu ----
// op: 0 vorticity, 1 interpolate_u_p, 2 interpolate_ω_p, 3 qcrit, 4 kinetic energy, 5 kinetic energy
// (interpolate first); Ip = pressure DOF range (0-based, half open)
template <typename T>
int launch_diag(const Grid<T> &g, const int Ip[3][2], int op, const T *in, T *out, Launch L);
template <typename T>
int launch_total_ke(const Grid<T> &g, const int Ip[3][2], const T *u, bool first, double *partial, int npartial,
                    double *out, Launch L);

// ---- bc.cu ----
// dims: bit mask of the dimensions to fill (default all)
template <typename T>
int launch_bc_u(const Grid<T> &g, const BcDesc &b, T *u, int dudt, Launch L, int dims = 7);
template <typename T>
int launch_bc_p(const Grid<T> &g, const BcDesc &b, T *p, Launch L, int dims = 7);
template <typename T>
int launch_bc_temp(const Grid<T> &g, const BcDesc &b, T *temp, Launch L);

template <typename T>
int launch_bc_p_fields(const Grid<T> &g, const BcDesc &b, T *fields, int nf, Launch L);
template <typename T>
int launch_zero_out_wall(const Grid<T> &g, const BcDesc &b, T *fields, int nf, Launch L);

// ---- les.cu ----
// eddy_viscosity_closure! (src/eddyviscosity.jl:356-378).  model: 1 Smagorinsky, 2 WALE, 3 Vreman,
// 4 QR.  G: 9 scalar fields, visc: 1 (scratch).  With fl = -(div sigma): k_out += fl, v += cv fl,
// acc += ca fl (null pointers are skipped).  Default: the two fused kernels (nu_t from u, then stress
// divergence from u and nu_t; G unused and may be null); the four-pass path with the stored tensor is taken
// when les_needs_tensor says so (PressureBC sides, decomposed grids, u aliasing an output, INSB200_LES_UNFUSED).
bool les_needs_tensor(int D, const int N[3], const BcDesc &b, const void *u, const void *k_out, const void *v,
                      const void *acc);
template <typename T>
int launch_les_closure(const Grid<T> &g, const BcDesc &b, int model, double C, const T *u, T *G, T *visc, T *k_out, T *v,
                       double cv, T *acc, double ca, bool ct_double, Launch L);

// ---- poisson.cu ----
struct Poisson;  // opaque
int poisson_create(const insb200_setup_t &s, const int Ip[3][2], int kind, Comm *comm, Poisson **out);
void poisson_destroy(Poisson *p);
long long poisson_bytes(const Poisson *p);
int poisson_iterations(const Poisson *p);
// energyspectrum (src/spectral.jl:75-117): ehat = kmax + 1 device doubles
int poisson_energy_spectrum(Poisson *ps, const void *u, long long cs, int kmax, double *ehat, Launch L);
// spectral sums of turbulence_statistics (src/spectral.jl:166-257): mom = 3 device doubles, sum of |k|^(2j) e(k), j = 0..2,
// with the dimensional wavenumbers 2 pi k / lengths[a]
int poisson_spectral_moments(Poisson *ps, const void *u, long long cs, const double lengths[3], double *mom, Launch L);  // psolver_cg: iterations of the last solve
// random_field (src/initializers.jl:109-183): the interior of u holds white noise (fill == 0) or is filled with
// Philox N(0, 1) from `seed` first; e0_host[kappa] = target energy of shell kappa (kdiag + 1 entries)
int poisson_random_field(Poisson *ps, void *u, long long cs, const double *e0_host, int nshell, int fill,
                         unsigned long long seed, const double xlims[3][2], Launch L);
// ---- cg.cu ----
struct CgSolver;
int cg_create(const insb200_setup_t &s, const int Ip[3][2], double reltol, double abstol, long long maxiter,
              CgSolver **out);
void cg_destroy(CgSolver *c);
long long cg_bytes(const CgSolver *c);
int cg_iterations(const CgSolver *c);
int cg_solve(CgSolver *c, void *p_ghosted, Launch L);
// Optional source of the right-hand side: when given (and poisson_fuses_divergence), the first
// pass of the solve evaluates divergence! + scalewithvolume! of the velocity `u` itself and
// p_ghosted is only written (with the solution).  Tables are the ghosted Setup tables.
struct DivSource {
    const void *u;
    long long cs;
    const void *rdlg[3];
    const void *dl[3];
    int wrap;
};
bool poisson_fuses_divergence(const Poisson *ps);
// Separate input / output arrays in the internal layout of the fused time-step path (fused.h): the right-hand
// side is read from `rhs`, the solution written to `sol` together with its periodic ghost images (img, see
// FastGeom::img); both point at the first interior cell, strides rs1 / rs2.  Spectral fast path only.
struct PoissonIO {
    const void *rhs;
    void *sol;
    long long rs1, rs2;
    int img;
};
bool poisson_supports_io(const Poisson *ps);
int poisson_solve(Poisson *ps, void *p_ghosted, Launch L, const DivSource *div = nullptr, const PoissonIO *io = nullptr);
// direct peer-store transposes: export this rank's IPC handles / import everybody's
constexpr int POISSON_IPC_BYTES = 192;
int poisson_ipc_export(Poisson *ps, void *blob);
int poisson_ipc_import(Poisson *ps, const void *blobs);
void poisson_ipc_close(Poisson *ps);

}  // namespace insb
