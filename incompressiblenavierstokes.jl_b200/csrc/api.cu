// api.cu -- plan management, time-step orchestration and the extern "C" surface of
// libinsb200.so (declared in include/insb200.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "fused.h"
#include "stencil_math.cuh"
#include "internal.h"

namespace insb {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

}  // namespace insb

using namespace insb;

struct insb200_plan : public insb::Profiler {
    insb200_setup_t s;
    bool f64 = false;
    size_t esz = 4;
    int D = 3;
    int dev = 0;
    long long ncell = 0;
    int Ip[3][2];
    int Iu[3][3][2];  // [comp][dim][lo/hi)
    std::vector<std::vector<unsigned char>> host_tables;  // keeps copies of the caller's tables
    void *dev_tables = nullptr;
    Grid<float> gf;
    Grid<double> gd;
    BcDesc bc;
    Comm *comm = nullptr;  // z-slab decomposition (world > 1)
    int nzl = 0;           // interior planes owned by this rank
    int Iu_loc[3][3][2];   // Iu with the z range of this slab (cfl reduction)
    int wrap = 0;          // dims whose periodic ghost handling is fused into the divergence / projection kernels
    Poisson *ps = nullptr;
    int psolver = INSB200_PSOLVER_NONE;
    // work buffers (device): two stage buffers and the accumulator for u (and temp), p scratch,
    // lazily allocated k storage for general tableaus
    void *V[2] = {nullptr, nullptr}, *ACC = nullptr, *P = nullptr;
    void *TV[2] = {nullptr, nullptr}, *TACC = nullptr;
    std::vector<void *> K, TK;
    void *HU = nullptr, *HT = nullptr;  // device state of the host-buffer entry point
    void *HU2 = nullptr, *HT2 = nullptr;  // second and
    void *HU3 = nullptr, *HT3 = nullptr;  // third slot of the pipelined ensemble entry point
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the ensemble entry point
    cudaStream_t s_graph = nullptr;                // capture / replay stream of insb200_solve_unsteady
    double *scratch = nullptr;  // reductions
    double *ke_partial = nullptr;  // block partials of total_kinetic_energy
    void *LES_G = nullptr, *LES_visc = nullptr;  // eddy-viscosity scratch: 9 tensor components + viscosity
    // fused periodic time-step path (fused.h): padded internal arrays -- two unprojected stage velocities,
    // the running accumulator, the pressure and the scaled divergence -- and their tensor maps
    void *FV[2] = {nullptr, nullptr}, *FACC = nullptr, *FP = nullptr, *FD = nullptr;
    void *FUB = nullptr;  // slab mode: plane "-1" of the caller's velocity (from the rank below), [3][N1][N0]
    // chained steps (insb200_solve_unsteady): between two steps of the native loop the state stays in its lazy form
    // (unprojected velocity FV[chain_slot] + pressure FP) -- no explicit projection into the caller's array, and the
    // first stage of the next step reads it by TMA; FU0 keeps the step's projected input for the stage combine.
    // The caller's array is materialised before anything else looks at it (call-backs, CFL, the end of the loop).
    void *FU0 = nullptr;
    bool chain_in = false, chain_out = false, chain_valid = false;
    int chain_slot = 1;
    FusedCtx fctx{};
    bool fused_ready = false;
    long long bytes = 0;
    long long launches = 0;
    // diffusive / thermal step limits (host, from the tables)
    double dmin_u[3], dmin_p[3];
    // profiling
    bool prof = false;
    bool prof_legacy = false;  // keep the stage-by-stage path (per-kernel profiling groups of round 1)
    double prof_ms[NPROF] = {};
    long long prof_cnt[NPROF] = {};
    struct Bracket {
        cudaEvent_t a, b;
        int group;
    };
    std::vector<Bracket> brackets;
    std::vector<cudaEvent_t> evpool;
    std::vector<std::pair<int, cudaEvent_t>> open_;  // begin events waiting for their end
    cudaEvent_t get_event() {
        cudaEvent_t e;
        if (!evpool.empty()) {
            e = evpool.back();
            evpool.pop_back();
        } else
            cudaEventCreate(&e);
        return e;
    }
    void begin(int group, cudaStream_t st) override {
        if (!prof) return;
        cudaEvent_t e = get_event();
        cudaEventRecord(e, st);
        open_.push_back({group, e});
    }
    void end(int group, cudaStream_t st) override {
        if (!prof) return;
        for (size_t i = open_.size(); i-- > 0;)
            if (open_[i].first == group) {
                cudaEvent_t b = get_event();
                cudaEventRecord(b, st);
                brackets.push_back({open_[i].second, b, group});
                open_.erase(open_.begin() + i);
                return;
            }
    }
};

namespace {

int offset_u(int bc, bool isright, bool isnormal) {
    switch (bc) {
        case INSB200_BC_PERIODIC: return 1;
        case INSB200_BC_DIRICHLET:
        case INSB200_BC_SYMMETRIC: return 1 + (isright && isnormal ? 1 : 0);
        default: return 1 + (!isright && !isnormal ? 1 : 0);
    }
}
int offset_p(int bc, bool isright) { return bc == INSB200_BC_PRESSURE ? 1 + (!isright ? 1 : 0) : 1; }

template <typename T>
int build_grid(insb200_plan *pl, Grid<T> &g) {
    const insb200_setup_t &s = pl->s;
    const int D = s.D;
    g.D = D;
    size_t total = 0;
    for (int a = 0; a < 3; a++) {
        g.N[a] = a < D ? s.N[a] : 1;
        total += a < D ? (size_t)s.N[a] : 0;
    }
    g.s[0] = 1;
    g.s[1] = g.N[0];
    g.s[2] = (long long)g.N[0] * g.N[1];
    g.cs = (long long)g.N[0] * g.N[1] * g.N[2];
    const int NT = 10;
    std::vector<T> h(total * NT);
    const T eps = std::is_same<T, float>::value ? (T)1.1920929e-07f : (T)2.220446049250313e-16;
    const T tol = 2 * eps;
    size_t off = 0;
    std::vector<size_t> offs(D);
    for (int a = 0; a < D; a++) {
        int n = s.N[a];
        offs[a] = off;
        const T *dl = (const T *)s.Delta[a], *du = (const T *)s.DeltaU[a];
        const T *a1 = (const T *)s.A1[a], *a2 = (const T *)s.A2[a];
        for (int i = 0; i < n; i++) {
            T *row = h.data() + off * NT;
            row[0 * n + i] = dl[i];
            row[1 * n + i] = du[i];
            row[2 * n + i] = T(1) / dl[i];
            row[3 * n + i] = T(1) / du[i];
            row[4 * n + i] = dl[i] > tol ? T(1) / dl[i] : T(0);
            row[5 * n + i] = du[i] > tol ? T(1) / du[i] : T(0);
            row[6 * n + i] = a1[i];
            row[7 * n + i] = a2[i];
            if (i + 1 < n) {
                row[8 * n + i] = dl[i + 1] / (dl[i] + dl[i + 1]);
                row[9 * n + i] = dl[i] / (dl[i] + dl[i + 1]);
            } else {
                row[8 * n + i] = T(0.5);
                row[9 * n + i] = T(0.5);
            }
        }
        off += n;
    }
    INSB_CUDA(cudaMalloc(&pl->dev_tables, h.size() * sizeof(T)));
    INSB_CUDA(cudaMemcpy(pl->dev_tables, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    pl->bytes += (long long)(h.size() * sizeof(T));
    const T *base = (const T *)pl->dev_tables;
    for (int a = 0; a < 3; a++) {
        int aa = a < D ? a : 0;
        int n = s.N[aa];
        const T *row = base + offs[aa] * NT;
        g.dl[a] = row + 0 * n;
        g.du[a] = row + 1 * n;
        g.rdl[a] = row + 2 * n;
        g.rdu[a] = row + 3 * n;
        g.rdlg[a] = row + 4 * n;
        g.rdug[a] = row + 5 * n;
        g.a1[a] = row + 6 * n;
        g.a2[a] = row + 7 * n;
        g.wa[a] = row + 8 * n;
        g.wb[a] = row + 9 * n;
    }
    if (pl->comm) {  // this rank's slab: planes [rank*nzl, rank*nzl + nzl + 2) of the global grid
        const long long zoff = (long long)comm_rank(pl->comm) * pl->nzl;
        g.N[2] = pl->nzl + 2;
        g.cs = (long long)g.N[0] * g.N[1] * g.N[2];
        g.dl[2] += zoff, g.du[2] += zoff, g.rdl[2] += zoff, g.rdu[2] += zoff, g.rdlg[2] += zoff;
        g.rdug[2] += zoff, g.a1[2] += zoff, g.a2[2] += zoff, g.wa[2] += zoff, g.wb[2] += zoff;
    }
    // step limits (solver.jl:157-166, :188-197)
    for (int a = 0; a < D; a++) {
        const T *dl = (const T *)s.Delta[a], *du = (const T *)s.DeltaU[a];
        T mu = du[pl->Iu[a][a][0]], mp = dl[pl->Ip[a][0]];
        for (int i = pl->Iu[a][a][0]; i < pl->Iu[a][a][1]; i++) mu = std::min(mu, du[i]);
        for (int i = pl->Ip[a][0]; i < pl->Ip[a][1]; i++) mp = std::min(mp, dl[i]);
        pl->dmin_u[a] = (double)mu;
        pl->dmin_p[a] = (double)mp;
    }
    return 0;
}

int dev_alloc(insb200_plan *pl, void **p, size_t bytes) {
    INSB_CUDA(cudaMalloc(p, bytes));
    INSB_CUDA(cudaMemset(*p, 0, bytes));
    pl->bytes += (long long)bytes;
    return 0;
}

int resolve_default_psolver(const insb200_plan *pl) {
    // default_psolver (src/pressure.jl:58-78)
    const insb200_setup_t &s = pl->s;
    const int D = s.D;
    bool isper[3], iswall[3], isuni[3];
    for (int a = 0; a < D; a++) {
        isper[a] = s.bc_u[a][0] == INSB200_BC_PERIODIC && s.bc_u[a][1] == INSB200_BC_PERIODIC;
        iswall[a] = s.bc_u[a][0] == INSB200_BC_DIRICHLET && s.bc_u[a][1] == INSB200_BC_DIRICHLET;
        double tol = std::sqrt(pl->f64 ? 2.220446049250313e-16 : 1.1920929e-07);
        auto get = [&](int i) {
            return pl->f64 ? ((const double *)s.Delta[a])[i] : (double)((const float *)s.Delta[a])[i];
        };
        double ref = get(pl->Ip[a][0]);
        isuni[a] = true;
        for (int i = pl->Ip[a][0]; i < pl->Ip[a][1]; i++)
            if (std::fabs(get(i) - ref) > tol * std::fmax(std::fabs(get(i)), std::fabs(ref))) isuni[a] = false;
    }
    bool allper = true;
    for (int a = 0; a < D; a++) allper = allper && isper[a] && isuni[a];
    if (allper) return INSB200_PSOLVER_SPECTRAL;
    for (int a = 0; a < D; a++) {
        bool ch = iswall[a];
        for (int b = 0; b < D; b++) ch = ch && (b == a || (isper[b] && isuni[b]));
        if (ch) return INSB200_PSOLVER_TRIDIAGONAL;
    }
    // the reference builds the sparse direct solver here (pressure.jl:71-77); on the GPU path the
    // general fallback is psolver_cg (single GPU)
    return pl->comm ? INSB200_PSOLVER_NONE : INSB200_PSOLVER_CG;
}

Launch L_of(insb200_plan *pl, void *stream) { return Launch{(cudaStream_t)stream, &pl->launches, pl}; }

// ---- profiling brackets -------------------------------------------------------------
struct Scope {
    insb200_plan *pl;
    cudaStream_t st;
    int group;
    Scope(insb200_plan *p, cudaStream_t s, int g) : pl(p), st(s), group(g) { pl->begin(group, st); }
    ~Scope() { pl->end(group, st); }
};

template <typename T>
const Grid<T> &grid_of(const insb200_plan *pl);
template <>
const Grid<float> &grid_of<float>(const insb200_plan *pl) { return pl->gf; }
template <>
const Grid<double> &grid_of<double>(const insb200_plan *pl) { return pl->gd; }

// fused = the time-stepper variant: the periodic dimensions in pl->wrap need no ghost fills of
// v and p (the kernels wrap their neighbour indices) and the projected velocity is stored
// together with its periodic ghost images, i.e. project! followed by apply_bc_u!.
template <typename T>
int project_t(insb200_plan *pl, T *u, T *p, void *stream, bool fused = false) {
    // project! (src/pressure.jl:42-55)
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    const int wrap = fused ? pl->wrap : 0, rest = 7 & ~wrap;
    if (!pl->ps) {
        set_error("this plan has no Poisson solver (psolver = NONE / unsupported BC combination)");
        return 1;
    }
    if (poisson_fuses_divergence(pl->ps)) {
        // divergence! + scalewithvolume! evaluated inside the first FFT pass (no d array)
        DivSource ds{};
        ds.u = u;
        ds.cs = g.cs;
        for (int a = 0; a < g.D; a++) {
            ds.rdlg[a] = g.rdlg[a];
            ds.dl[a] = g.dl[a];
        }
        ds.wrap = wrap;
        Scope sc(pl, L.stream, 2);
        INSB_TRY(poisson_solve(pl->ps, p, L, &ds));
    } else {
        {
            Scope sc(pl, L.stream, 1);
            INSB_TRY(launch_divergence<T>(g, u, p, true, L, wrap));
        }
        Scope sc(pl, L.stream, 2);
        INSB_TRY(poisson_solve(pl->ps, p, L));
    }
    if (rest & ((1 << g.D) - 1)) {
        Scope sc(pl, L.stream, 4);
        INSB_TRY(launch_bc_p<T>(g, pl->bc, p, L, rest));
    }
    {
        Scope sc(pl, L.stream, 3);
        INSB_TRY(launch_pressuregradient<T>(g, p, (T *)nullptr, u, L, wrap));
    }
    return 0;
}

template <typename T>
void fill_params(StageArgs<T> &a, const insb200_params_t *pr, const insb200_plan *pl) {
    a.nu = pr->viscosity;
    a.kappa = pr->conductivity;
    a.grav = pr->gravity;
    a.gdir = pr->gdir;
    a.dodiss = pr->dodissipation;
    a.bouss = pr->boussinesq;
    for (int c = 0; c < 3; c++) a.bodyforce[c] = pr->bodyforce[c];
    a.eddy_model = pr->eddy_model;
    a.eddy_C = pr->eddy_C;
    a.conv = 1;
    a.diff = 1;
    a.add_to_k = 0;
    a.zero_outside = 0;
}

int ensure_work(insb200_plan *pl, bool temp) {
    size_t vb = (size_t)pl->ncell * pl->D * pl->esz, sb = (size_t)pl->ncell * pl->esz;
    if (!pl->V[0]) {
        INSB_TRY(dev_alloc(pl, &pl->V[0], vb));
        INSB_TRY(dev_alloc(pl, &pl->V[1], vb));
        INSB_TRY(dev_alloc(pl, &pl->ACC, vb));
        INSB_TRY(dev_alloc(pl, &pl->P, sb));
    }
    if (temp && !pl->TV[0]) {
        INSB_TRY(dev_alloc(pl, &pl->TV[0], sb));
        INSB_TRY(dev_alloc(pl, &pl->TV[1], sb));
        INSB_TRY(dev_alloc(pl, &pl->TACC, sb));
    }
    return 0;
}

int ensure_k(insb200_plan *pl, int count, bool temp) {
    size_t vb = (size_t)pl->ncell * pl->D * pl->esz, sb = (size_t)pl->ncell * pl->esz;
    while ((int)pl->K.size() < count) {
        void *p;
        INSB_TRY(dev_alloc(pl, &p, vb));
        pl->K.push_back(p);
    }
    while (temp && (int)pl->TK.size() < count) {
        void *p;
        INSB_TRY(dev_alloc(pl, &p, sb));
        pl->TK.push_back(p);
    }
    return 0;
}


// ---- fused periodic path (fused.h) -----------------------------------------------------
// 3-D, all-periodic, psolver_spectral on the register-FFT fast path, plain navierstokes! forcing (+ constant
// body force), one GPU.  Everything else keeps the stage-by-stage path below.
bool fused_applicable(const insb200_plan *pl, const insb200_params_t *pr) {
    const bool off = getenv("INSB200_NO_FUSED_STEP") != nullptr;  // (read per call: the tests toggle it)
    if (off || pl->D != 3 || pl->prof_legacy) return false;
    if (pl->comm && getenv("INSB200_NO_FUSED_SLAB") != nullptr) return false;
    for (int a = 0; a < 3; a++)
        if (pl->s.bc_u[a][0] != INSB200_BC_PERIODIC || pl->s.bc_u[a][1] != INSB200_BC_PERIODIC) return false;
    if (pr->boussinesq || pr->eddy_model != 0) return false;
    if (!pl->ps || pl->psolver != INSB200_PSOLVER_SPECTRAL || !poisson_supports_io(pl->ps)) return false;
    const int n0 = pl->s.N[0] - 2, n1 = pl->s.N[1] - 2, n2 = pl->comm ? pl->nzl : pl->s.N[2] - 2;
    if (n0 % 2 || n0 < 8 || n1 < 8 || n2 < 8) return false;
    const FusedLayout L = make_fused_layout(n0, n1, n2);
    return 3 * L.cs < (1LL << 31) && 3 * (long long)pl->ncell < (1LL << 32);
}

int ensure_fused(insb200_plan *pl) {
    if (pl->fused_ready) return 0;
    const int n0 = pl->s.N[0] - 2, n1 = pl->s.N[1] - 2, n2g = pl->s.N[2] - 2, n2 = pl->comm ? pl->nzl : n2g;
    pl->fctx.L = make_fused_layout(n0, n1, n2);
    pl->fctx.zwrap = pl->comm ? 0 : 1;
    pl->fctx.zoff = pl->comm ? comm_rank(pl->comm) * pl->nzl : 0;
    pl->fctx.n2g = n2g;
    if (pl->comm) INSB_TRY(dev_alloc(pl, &pl->FUB, (size_t)3 * pl->s.N[0] * pl->s.N[1] * pl->esz));
    INSB_TRY(fused_maps_create(&pl->fctx.maps));
    const size_t sb = (size_t)pl->fctx.L.cs * pl->esz;
    INSB_TRY(dev_alloc(pl, &pl->FV[0], 3 * sb));
    INSB_TRY(dev_alloc(pl, &pl->FV[1], 3 * sb));
    INSB_TRY(dev_alloc(pl, &pl->FACC, 3 * sb));
    INSB_TRY(dev_alloc(pl, &pl->FP, sb));
    INSB_TRY(dev_alloc(pl, &pl->FD, sb));
    pl->fused_ready = true;
    return 0;
}

// One fused stage followed by its Poisson solve: the stage kernel leaves the unprojected velocity in v_out
// and the scaled divergence in FD; the solve turns FD into the pressure FP (with its ghost images).
// Slab mode: the two z ghost planes on either side of `ncomp` padded internal arrays come from the neighbouring
// ranks (whole padded planes: their x / y ghost images travel along).
template <typename T>
int fused_halo(insb200_plan *pl, T *f, int ncomp, Launch L) {
    const FusedLayout &FL = pl->fctx.L;
    if (L.counter) *L.counter += comm_exchange_launches(pl->comm, (size_t)FusedLayout::OZ * FL.s2 * sizeof(T), ncomp);
    return comm_halo_planes(pl->comm, f, sizeof(T), FL.s2, FL.NPZ, FusedLayout::OZ, ncomp, FL.cs, L.stream);
}

template <typename T>
int fused_stage_and_solve(insb200_plan *pl, const FusedStage<T> &st_in, bool ct_double, void *stream) {
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    FusedStage<T> st = st_in;
    if (pl->comm && st.u_ref) {
        // first stage of a step on a slab: the stage output on ghost plane 0 needs the caller's velocity on
        // plane -1 = the top interior plane but one of the rank below
        Scope sc(pl, L.stream, 4);
        const size_t pb = (size_t)g.s[2] * sizeof(T);
        if (L.counter) *L.counter += comm_exchange_launches(pl->comm, pb, 3);
        INSB_TRY(comm_exchange(pl->comm, (const char *)st.u_ref + (size_t)(pl->nzl - 1) * pb, nullptr, pl->FUB, nullptr, pb, 3,
                               (size_t)g.cs * sizeof(T), pb, L.stream));
        st.u_below = (const T *)pl->FUB;
    }
    {
        Scope sc(pl, L.stream, 0);
        INSB_TRY(launch_stage_fused<T>(g, pl->fctx, st, ct_double, L));
    }
    if (pl->comm) {
        Scope sc(pl, L.stream, 4);
        INSB_TRY(fused_halo<T>(pl, st.v_out, 3, L));
        if (st.acc_out && st.acc_halo) INSB_TRY(fused_halo<T>(pl, st.acc_out, 3, L));
        if (st.u0_out) INSB_TRY(fused_halo<T>(pl, st.u0_out, 3, L));
    }
    const FusedLayout &FL = pl->fctx.L;
    PoissonIO io{};
    io.rhs = (const T *)pl->FD + FL.at(1, 1, 1);
    io.sol = (T *)pl->FP + FL.at(1, 1, 1);
    io.rs1 = FL.s1;
    io.rs2 = FL.s2;
    io.img = 1 | (pl->fctx.zwrap ? 2 : 0);
    {
        Scope sc(pl, L.stream, 2);
        INSB_TRY(poisson_solve(pl->ps, nullptr, L, nullptr, &io));
    }
    if (pl->comm) {
        Scope sc(pl, L.stream, 4);
        INSB_TRY(fused_halo<T>(pl, (T *)pl->FP, 1, L));
    }
    return 0;
}

template <typename T>
int fused_finish(insb200_plan *pl, const T *v, T *U, void *stream) {
    Launch L = L_of(pl, stream);
    {
        Scope sc(pl, L.stream, 3);
        INSB_TRY(launch_project_out<T>(grid_of<T>(pl), pl->fctx, v, (const T *)pl->FP, U, L));
    }
    if (pl->comm) {  // z ghost planes of the result (apply_bc_u! along the decomposed direction)
        Scope sc(pl, L.stream, 4);
        INSB_TRY(launch_bc_u<T>(grid_of<T>(pl), pl->bc, U, 0, L, 4));
    }
    return 0;
}

template <typename T>
void fused_params(FusedStage<T> &st, const insb200_params_t *pr) {
    st.nu = pr->viscosity;
    for (int c = 0; c < 3; c++) st.bodyforce[c] = pr->bodyforce[c];
}

// Stage kernel followed by the eddy-viscosity closure when params ask for one: the closure term is
// applied to the outputs the stage kernel has just written (k, v, acc), like the Boussinesq pass.
template <typename T>
int stage_with_closures(insb200_plan *pl, const Grid<T> &g, const StageArgs<T> &a, bool ct_double, Launch L) {
    INSB_TRY(launch_stage<T>(g, a, ct_double, L));
    if (a.eddy_model != 0) {
        size_t sb = (size_t)pl->ncell * pl->esz;
        if (!pl->LES_G && les_needs_tensor(g.D, g.N, pl->bc, a.u, a.k_out, a.v, a.acc_out))
            INSB_TRY(dev_alloc(pl, &pl->LES_G, 9 * sb));
        if (!pl->LES_visc) INSB_TRY(dev_alloc(pl, &pl->LES_visc, sb));
        INSB_TRY(launch_les_closure<T>(g, pl->bc, a.eddy_model, a.eddy_C, a.u, (T *)pl->LES_G, (T *)pl->LES_visc, a.k_out,
                                       a.v, a.cv, a.acc_out, a.ca, ct_double, L));
    }
    return 0;
}

// One projected stage: v (already combined, interior only) -> BC, project, BC (+ temp BC).
template <typename T>
int finish_stage(insb200_plan *pl, T *v, T *tv, double t, void *stream) {
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    const int rest = 7 & ~pl->wrap & ((1 << g.D) - 1);  // dimensions that still need explicit ghost fills
    if (rest) {
        Scope sc(pl, L.stream, 4);
        INSB_TRY(launch_bc_u<T>(g, pl->bc, v, 0, L, rest));
    }
    INSB_TRY(project_t<T>(pl, v, (T *)pl->P, stream, true));
    {
        Scope sc(pl, L.stream, 4);
        if (rest) INSB_TRY(launch_bc_u<T>(g, pl->bc, v, 0, L, rest));
        if (tv) INSB_TRY(launch_bc_temp<T>(g, pl->bc, tv, L));
    }
    (void)t;
    return 0;
}

// timestep!(::ExplicitRungeKuttaMethod) -- step_explicit_runge_kutta.jl:4-57
template <typename T>
int step_rk_t(insb200_plan *pl, const double *A, const double *c, int s, int coef_dtype, T *U, T *TU,
              double t, double dt, const insb200_params_t *pr, void *stream) {
    if (s < 1 || s > MAXSTAGE) {
        set_error("number of stages must be in 1..8");
        return 1;
    }
    const bool bouss = pr->boussinesq != 0;
    if (bouss && !TU) {
        set_error("boussinesq forcing needs a temperature field");
        return 1;
    }
    INSB_TRY(ensure_work(pl, bouss));
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    const bool ct_double = pl->f64 || coef_dtype == INSB200_F64;
    // Δt * A[i,j] evaluated in the coefficient type (see include/insb200.h)
    auto coef = [&](int i, int j) -> double {
        if (coef_dtype == INSB200_F32) return (double)((float)dt * (float)A[i * s + j]);
        return dt * A[i * s + j];
    };
    // "fast" tableau: rows 0..s-2 only use their own stage; the last row is accumulated on the fly
    bool fast = true;
    for (int i = 0; i < s - 1; i++)
        for (int j = 0; j < i; j++) fast = fast && A[i * s + j] == 0.0;
    if (fast && s >= 2 && fused_applicable(pl, pr)) {
        // fused periodic path: per stage ONE stencil launch (lazy projection of the previous stage on load,
        // stage combine, divergence of its own output) + the Poisson solve; explicit projection only at the end
        INSB_TRY(ensure_fused(pl));
        // chained steps: the input is the lazy state the previous step left (cin), the output stays lazy (cout)
        const bool cin = pl->chain_in && pl->chain_valid, cout = pl->chain_out;
        // (allocated by the step that starts a chain: the next one may already run under stream capture)
        if ((cin || cout) && !pl->FU0) INSB_TRY(dev_alloc(pl, &pl->FU0, 3 * (size_t)pl->fctx.L.cs * pl->esz));
        const int cs = cin ? pl->chain_slot : 1;  // stage i writes FV[(cs + 1 + i) & 1]: never the buffer it reads
        auto slot = [&](int i) { return (cs + 1 + i) & 1; };
        for (int i = 0; i < s; i++) {
            const bool last = i == s - 1;
            FusedStage<T> st;
            fused_params(st, pr);
            st.cv = coef(i, i);
            st.v_out = (T *)pl->FV[slot(i)];
            st.d_out = (T *)pl->FD;
            if (i == 0 && !cin) {
                st.u_ref = U;
            } else {
                st.v_in = (const T *)pl->FV[i == 0 ? cs : slot(i - 1)];
                st.p_in = (const T *)pl->FP;
                if (i == 0) st.u0_out = (T *)pl->FU0;
            }
            if (!last) {
                st.acc_out = (T *)pl->FACC;
                st.acc_halo = i == s - 2;  // the last stage reads the accumulator as its vbase
                st.ca = coef(s - 1, i);
                if (i > 0) {
                    if (cin)
                        st.vbase_int = (const T *)pl->FU0;
                    else
                        st.vbase_ref = U;
                    st.acc_in = (const T *)pl->FACC;
                }
            } else {
                st.vbase_int = (const T *)pl->FACC;
            }
            INSB_TRY(fused_stage_and_solve<T>(pl, st, ct_double, stream));
            if (const char *e = getenv("INSB200_FUSED_STOP_AFTER"))  // debugging aid (tools/debug_fused.py)
                if (atoi(e) == i + 1) return 0;
        }
        pl->chain_slot = slot(s - 1);
        pl->chain_valid = cout;
        if (cout) return 0;
        return fused_finish<T>(pl, (const T *)pl->FV[slot(s - 1)], U, stream);
    }
    if (!fast) INSB_TRY(ensure_k(pl, s - 1, bouss));
    size_t vb = (size_t)pl->ncell * pl->D * pl->esz, sb = (size_t)pl->ncell * pl->esz;

    if (s == 1) {  // single-stage methods go through V[0]; carry the never-written cells along
        INSB_CUDA(cudaMemcpyAsync(pl->V[0], U, vb, cudaMemcpyDeviceToDevice, L.stream));
        if (bouss) INSB_CUDA(cudaMemcpyAsync(pl->TV[0], TU, sb, cudaMemcpyDeviceToDevice, L.stream));
    }
    const T *src = U;
    const T *tsrc = TU;
    for (int i = 0; i < s; i++) {
        const bool last = i == s - 1;
        T *dst = (last && s > 1) ? U : (T *)pl->V[i & 1];
        T *tdst = bouss ? ((last && s > 1) ? TU : (T *)pl->TV[i & 1]) : nullptr;
        StageArgs<T> a{};
        fill_params(a, pr, pl);
        a.u = src;
        a.temp = bouss ? tsrc : nullptr;
        a.v = dst;
        a.tv = tdst;
        a.cv = coef(i, i);
        if (fast) {
            if (last && s > 1) {
                a.vbase = (const T *)pl->ACC;
                a.tvbase = (const T *)pl->TACC;
            } else {
                a.vbase = U;
                a.tvbase = TU;
                if (s > 1) {
                    a.acc_in = i == 0 ? U : (const T *)pl->ACC;
                    a.acc_out = (T *)pl->ACC;
                    a.tacc_in = i == 0 ? TU : (const T *)pl->TACC;
                    a.tacc_out = bouss ? (T *)pl->TACC : nullptr;
                    a.ca = coef(s - 1, i);
                }
            }
        } else {
            a.vbase = U;
            a.tvbase = TU;
            a.nprev = i;
            for (int j = 0; j < i; j++) {
                a.kprev[j] = (const T *)pl->K[j];
                a.tkprev[j] = bouss ? (const T *)pl->TK[j] : nullptr;
                a.cprev[j] = coef(i, j);
            }
            if (!last) {
                a.k_out = (T *)pl->K[i];
                a.tk_out = bouss ? (T *)pl->TK[i] : nullptr;
            }
        }
        {
            Scope sc(pl, L.stream, 0);
            INSB_TRY(stage_with_closures<T>(pl, g, a, ct_double, L));
        }
        double ti = t + c[i] * dt;
        INSB_TRY(finish_stage<T>(pl, dst, tdst, ti, stream));
        src = dst;
        tsrc = tdst;
    }
    if (s == 1) {  // single-stage methods wrote into V[0]
        INSB_CUDA(cudaMemcpyAsync(U, pl->V[0], vb, cudaMemcpyDeviceToDevice, L.stream));
        if (bouss) INSB_CUDA(cudaMemcpyAsync(TU, pl->TV[0], sb, cudaMemcpyDeviceToDevice, L.stream));
    }
    return 0;
}

// timestep!(::LMWray3) -- step_lmwray3.jl:3-94
template <typename T>
int step_wray_t(insb200_plan *pl, T *U, T *TU, double t, double dt, const insb200_params_t *pr,
                void *stream) {
    const bool bouss = pr->boussinesq != 0;
    if (bouss && !TU) {
        set_error("boussinesq forcing needs a temperature field");
        return 1;
    }
    INSB_TRY(ensure_work(pl, bouss));
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    // coefficients in the field type (step_lmwray3.jl:60-62); axpy! converts the scalar to T
    const T a3[3] = {(T)(8.0 / 15.0), (T)(5.0 / 12.0), (T)(3.0 / 4.0)};
    const T b3[2] = {(T)(1.0 / 4.0), (T)0};
    const T c3[3] = {(T)(8.0 / 15.0), (T)(2.0 / 3.0), (T)1};
    if (fused_applicable(pl, pr)) {
        INSB_TRY(ensure_fused(pl));
        const bool cin = pl->chain_in && pl->chain_valid, cout = pl->chain_out;  // chained steps, see step_rk_t
        const int cs = cin ? pl->chain_slot : 1;
        auto slot = [&](int i) { return (cs + 1 + i) & 1; };
        for (int i = 0; i < 3; i++) {
            FusedStage<T> st;
            fused_params(st, pr);
            st.cv = (double)(T)((double)a3[i] * dt);
            st.v_out = (T *)pl->FV[slot(i)];
            st.d_out = (T *)pl->FD;
            if (i == 0) {  // statestart += b[0] dt k; v = u + a[0] dt k
                if (cin) {
                    st.v_in = (const T *)pl->FV[cs];
                    st.p_in = (const T *)pl->FP;
                } else {
                    st.u_ref = U;
                }
                st.acc_out = (T *)pl->FACC;
                st.acc_halo = true;  // stages 2 and 3 read it as their vbase
                st.ca = (double)(T)((double)b3[0] * dt);
            } else {
                st.v_in = (const T *)pl->FV[slot(i - 1)];
                st.p_in = (const T *)pl->FP;
                st.vbase_int = (const T *)pl->FACC;
            }
            INSB_TRY(fused_stage_and_solve<T>(pl, st, pl->f64, stream));
        }
        pl->chain_slot = slot(2);
        pl->chain_valid = cout;
        if (cout) return 0;
        return fused_finish<T>(pl, (const T *)pl->FV[slot(2)], U, stream);
    }
    const T *src = U;
    const T *tsrc = TU;
    for (int i = 0; i < 3; i++) {
        const bool last = i == 2;
        T *dst = last ? U : (T *)pl->V[i & 1];
        T *tdst = bouss ? (last ? TU : (T *)pl->TV[i & 1]) : nullptr;
        StageArgs<T> a{};
        fill_params(a, pr, pl);
        a.u = src;
        a.temp = bouss ? tsrc : nullptr;
        a.v = dst;
        a.tv = tdst;
        a.cv = (double)(T)((double)a3[i] * dt);
        a.vbase = i == 0 ? U : (const T *)pl->ACC;
        a.tvbase = i == 0 ? TU : (const T *)pl->TACC;
        if (i == 0) {  // statestart += b[0] Δt k ; (b[1] = 0 leaves it unchanged)
            a.acc_in = U;
            a.acc_out = (T *)pl->ACC;
            a.tacc_in = TU;
            a.tacc_out = bouss ? (T *)pl->TACC : nullptr;
            a.ca = (double)(T)((double)b3[0] * dt);
        }
        {
            Scope sc(pl, L.stream, 0);
            INSB_TRY(stage_with_closures<T>(pl, g, a, pl->f64, L));
        }
        double ti = t + (double)c3[i] * dt;
        INSB_TRY(finish_stage<T>(pl, dst, tdst, ti, stream));
        src = dst;
        tsrc = tdst;
    }
    // the trailing apply_bc_u!/apply_bc_temp! at t + Δt (:90-91) repeats the fills above for
    // constant-in-time BCs; function-valued BCs are refreshed by the host through
    // insb200_set_bc_u_plane + insb200_apply_bc_u.
    return 0;
}

template <typename T>
static int rhs_t(insb200_plan *pl, const T *u, const T *temp, T *ku, T *ktemp, const insb200_params_t *pr,
                 int conv, int diff, int add, void *stream) {
    StageArgs<T> a{};
    fill_params(a, pr, pl);
    a.u = u;
    a.temp = pr->boussinesq ? temp : nullptr;
    a.k_out = ku;
    a.tk_out = pr->boussinesq ? ktemp : nullptr;
    a.zero_outside = add ? 0 : 1;
    a.add_to_k = add;
    a.conv = conv;
    a.diff = diff;
    Scope sc(pl, (cudaStream_t)stream, 0);
    if (!(conv && diff)) a.eddy_model = 0;  // convection! / diffusion! alone never carry the closure
    return stage_with_closures<T>(pl, grid_of<T>(pl), a, pl->f64, L_of(pl, stream));
}

template <typename T>
static int rhs_sciml_t(insb200_plan *pl, T *dudt, const T *u, double t, const insb200_params_t *pr,
                       void *stream) {
    // right_hand_side! (src/sciml.jl:35-48)
    INSB_TRY(ensure_work(pl, false));
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    size_t vb = (size_t)pl->ncell * pl->D * pl->esz;
    T *tmp = (T *)pl->V[0];
    INSB_CUDA(cudaMemcpyAsync(tmp, u, vb, cudaMemcpyDeviceToDevice, L.stream));
    INSB_TRY(launch_bc_u<T>(g, pl->bc, tmp, 0, L));
    insb200_params_t p2 = *pr;
    p2.boussinesq = 0;
    INSB_TRY(rhs_t<T>(pl, tmp, nullptr, dudt, nullptr, &p2, 1, 1, 0, stream));
    INSB_TRY(launch_bc_u<T>(g, pl->bc, dudt, 1, L));
    INSB_TRY(project_t<T>(pl, dudt, (T *)pl->P, stream));
    (void)t;
    return 0;
}

}  // namespace

// =========================================================================================
// extern "C"
// =========================================================================================
// ---- host-composed stages (insb200.h: insb200_stage_combine / _project / insb200_axpy) ----------------
namespace {
template <typename T>
struct CombineArgs {
    const T *k[MAXSTAGE];
    double c[MAXSTAGE];
    int nk;
};

template <typename T, typename CT>
__global__ void __launch_bounds__(256) combine_kernel(T *__restrict__ dst, const T *base, CombineArgs<T> a, long long n) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    T x = base[q];
    for (int j = 0; j < a.nk; j++) x = axpy_rn<T, CT>(x, a.c[j], a.k[j][q]);
    dst[q] = x;
}

template <typename T>
int combine_t(insb200_plan *pl, T *dst, const T *base, int nk, const void *const *k, const double *coef, int coef_dtype,
              int ncomp, void *stream) {
    if (nk < 0 || nk > MAXSTAGE) {
        set_error("stage combine: 0..8 stage arrays");
        return 1;
    }
    CombineArgs<T> a{};
    a.nk = nk;
    for (int j = 0; j < nk; j++) a.k[j] = (const T *)k[j], a.c[j] = coef[j];
    const long long n = pl->ncell * ncomp;
    Launch L = L_of(pl, stream);
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (pl->f64 || coef_dtype == INSB200_F64)
        combine_kernel<T, double><<<nb, 256, 0, L.stream>>>(dst, base, a, n);
    else
        combine_kernel<T, T><<<nb, 256, 0, L.stream>>>(dst, base, a, n);
    if (L.counter) ++*L.counter;
    INSB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
int combine_project_t(insb200_plan *pl, T *u, const T *u0, T *temp, const T *temp0, int nk, const void *const *k,
                      const void *const *ktemp, const double *coef, int coef_dtype, T *p, double t, void *stream) {
    if (!pl->ps) {
        set_error("stage_combine_project: the plan has no pressure solver");
        return 1;
    }
    if (temp && (!temp0 || !ktemp)) {
        set_error("stage_combine_project: temp needs temp_start and ktemp");
        return 1;
    }
    INSB_TRY(combine_t<T>(pl, u, u0, nk, k, coef, coef_dtype, pl->D, stream));
    if (temp) INSB_TRY(combine_t<T>(pl, temp, temp0, nk, ktemp, coef, coef_dtype, 1, stream));
    const Grid<T> &g = grid_of<T>(pl);
    Launch L = L_of(pl, stream);
    INSB_TRY(launch_bc_u<T>(g, pl->bc, u, 0, L));
    INSB_TRY(project_t<T>(pl, u, p, stream));
    INSB_TRY(launch_bc_u<T>(g, pl->bc, u, 0, L));
    if (temp) INSB_TRY(launch_bc_temp<T>(g, pl->bc, temp, L));
    (void)t;  // function-valued planes are refreshed by the host before the call (insb200_set_bc_u_plane)
    return 0;
}
}  // namespace


#define API_BEGIN try {
#define API_END                                        \
    }                                                  \
    catch (const std::exception &e) {                  \
        insb::set_error(std::string("exception: ") + e.what()); \
        return 1;                                      \
    }                                                  \
    catch (...) {                                      \
        insb::set_error("unknown exception");          \
        return 1;                                      \
    }

extern "C" {

const char *insb200_last_error(void) { return insb::g_err.c_str(); }
int insb200_version(void) { return INSB200_VERSION; }

int insb200_plan_create(const insb200_setup_t *setup, insb200_plan_t **out) {
    API_BEGIN
    if (!setup || !out) {
        set_error("null argument");
        return 1;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libinsb200 has no CPU fallback");
        return 1;
    }
    const insb200_setup_t &s = *setup;
    if (s.D != 2 && s.D != 3) {
        set_error("D must be 2 or 3");
        return 1;
    }
    if (s.dtype != INSB200_F32 && s.dtype != INSB200_F64) {
        set_error("dtype must be INSB200_F32 or INSB200_F64");
        return 1;
    }
    for (int a = 0; a < s.D; a++) {
        if (s.N[a] < 4) {
            set_error("N[a] must be >= 4");
            return 1;
        }
        if (!s.Delta[a] || !s.DeltaU[a] || !s.A1[a] || !s.A2[a]) {
            set_error("metric tables must be given");
            return 1;
        }
        bool perl = s.bc_u[a][0] == INSB200_BC_PERIODIC, perr = s.bc_u[a][1] == INSB200_BC_PERIODIC;
        if (perl != perr) {
            set_error("Periodic boundary conditions must be periodic on both sides");
            return 1;
        }
    }
    if (s.world > 1) {
        // z periodic, or closed by walls / symmetry planes (the physical boundary planes then live on the first
        // and the last rank, bc.cu); a pressure boundary shifts the DOF ranges and is not decomposed
        if (s.D != 3 || s.bc_u[2][0] == INSB200_BC_PRESSURE || s.bc_u[2][1] == INSB200_BC_PRESSURE ||
            (s.bc_u[2][0] == INSB200_BC_PERIODIC) != (s.bc_u[2][1] == INSB200_BC_PERIODIC)) {
            set_error("world > 1: the slab decomposition needs a 3-D grid that is periodic, wall-bounded or symmetric in z");
            return 1;
        }
        if ((s.N[2] - 2) % s.world) {
            set_error("world > 1: the number of z volumes must be divisible by the number of ranks");
            return 1;
        }
    }
    insb200_plan *pl = new insb200_plan();
    pl->s = s;
    pl->f64 = s.dtype == INSB200_F64;
    pl->esz = pl->f64 ? 8 : 4;
    pl->D = s.D;
    if (s.device >= 0) INSB_CUDA(cudaSetDevice(s.device));
    INSB_CUDA(cudaGetDevice(&pl->dev));
    if (s.world > 1) {
        int rc = comm_create(s.rank, s.world, s.nccl_id, &pl->comm);
        if (rc) {
            delete pl;
            return rc;
        }
        pl->nzl = (s.N[2] - 2) / s.world;
    }
    // keep private copies of the caller's tables
    pl->host_tables.resize(4 * 3);
    for (int a = 0; a < s.D; a++) {
        const void *src[4] = {s.Delta[a], s.DeltaU[a], s.A1[a], s.A2[a]};
        const void **dst[4] = {&pl->s.Delta[a], &pl->s.DeltaU[a], &pl->s.A1[a], &pl->s.A2[a]};
        for (int q = 0; q < 4; q++) {
            auto &v = pl->host_tables[a * 4 + q];
            v.assign((const unsigned char *)src[q], (const unsigned char *)src[q] + s.N[a] * pl->esz);
            *dst[q] = v.data();
        }
    }
    pl->ncell = 1;
    for (int a = 0; a < 3; a++) {
        for (int c = 0; c < 3; c++) pl->Iu[c][a][0] = 0, pl->Iu[c][a][1] = 1;
        pl->Ip[a][0] = 0;
        pl->Ip[a][1] = 1;
    }
    for (int a = 0; a < s.D; a++) {
        pl->ncell *= s.N[a];
        pl->Ip[a][0] = offset_p(s.bc_u[a][0], false);
        pl->Ip[a][1] = s.N[a] - offset_p(s.bc_u[a][1], true);
        for (int c = 0; c < s.D; c++) {
            pl->Iu[c][a][0] = offset_u(s.bc_u[a][0], false, c == a);
            pl->Iu[c][a][1] = s.N[a] - offset_u(s.bc_u[a][1], true, c == a);
        }
    }
    std::memcpy(pl->Iu_loc, pl->Iu, sizeof(pl->Iu));
    if (pl->comm) {
        pl->ncell = (long long)s.N[0] * s.N[1] * (pl->nzl + 2);
        const int zoff = s.rank * pl->nzl;  // local plane = global plane - zoff
        for (int c = 0; c < 3; c++) {
            pl->Iu_loc[c][2][0] = std::max(1, pl->Iu[c][2][0] - zoff);
            pl->Iu_loc[c][2][1] = std::min(pl->nzl + 1, pl->Iu[c][2][1] - zoff);
        }
    }
    int r = pl->f64 ? build_grid<double>(pl, pl->gd) : build_grid<float>(pl, pl->gf);
    if (r) {
        insb200_plan_destroy(pl);
        return r;
    }
    std::memset(&pl->bc, 0, sizeof(pl->bc));
    pl->bc.comm = pl->comm;
    if (getenv("INSB200_NO_BC_FUSION") == nullptr)
        for (int a = 0; a < s.D; a++)
            if (s.bc_u[a][0] == INSB200_BC_PERIODIC && !(a == 2 && pl->comm)) pl->wrap |= 1 << a;
    std::memcpy(pl->bc.bc_u, s.bc_u, sizeof(s.bc_u));
    std::memcpy(pl->bc.bc_temp, s.has_temp ? s.bc_temp : s.bc_u, sizeof(s.bc_u));
    std::memcpy(pl->bc.u_val, s.bc_u_value, sizeof(s.bc_u_value));
    std::memcpy(pl->bc.temp_val, s.bc_temp_value, sizeof(s.bc_temp_value));
    int kind = s.psolver;
    if (kind == INSB200_PSOLVER_DEFAULT) {
        kind = resolve_default_psolver(pl);
        // Where the reference's default is the exact sparse direct solver, the stand-in is CG: make it
        // converge to round-off level (not the sqrt(eps) of psolver_cg's own default) unless the caller
        // chose tolerances, so that projected fields stay divergence free to about the same level.
        if (kind == INSB200_PSOLVER_CG && pl->s.cg_reltol == 0 && pl->s.cg_abstol == 0)
            pl->s.cg_reltol = -(pl->f64 ? 1e-13 : 1e-5);  // negative: "tight default", enables the stagnation stop
    }
    pl->psolver = kind;
    if (kind != INSB200_PSOLVER_NONE) {
        r = poisson_create(pl->s, pl->Ip, kind, pl->comm, &pl->ps);
        if (r) {
            insb200_plan_destroy(pl);
            return r;
        }
        pl->bytes += poisson_bytes(pl->ps);
    }
    if (cudaMalloc((void **)&pl->scratch, 4 * sizeof(double)) != cudaSuccess) {
        set_error("cudaMalloc failed");
        insb200_plan_destroy(pl);
        return 1;
    }
    *out = pl;
    return 0;
    API_END
}

int insb200_plan_destroy(insb200_plan_t *pl) {
    if (!pl) return 0;
    cudaDeviceSynchronize();
    if (pl->ps) poisson_destroy(pl->ps);
    for (void *p : {pl->V[0], pl->V[1], pl->ACC, pl->P, pl->TV[0], pl->TV[1], pl->TACC, pl->dev_tables,
                    (void *)pl->scratch, (void *)pl->ke_partial, pl->LES_G, pl->LES_visc, pl->HU, pl->HT, pl->HU2, pl->HT2, pl->HU3, pl->HT3,
                    pl->FV[0], pl->FV[1], pl->FACC, pl->FP, pl->FD, pl->FUB, pl->FU0})
        if (p) cudaFree(p);
    if (pl->fctx.maps) fused_maps_destroy(pl->fctx.maps);
    if (pl->s_in) cudaStreamDestroy(pl->s_in);
    if (pl->s_out) cudaStreamDestroy(pl->s_out);
    if (pl->s_graph) cudaStreamDestroy(pl->s_graph);
    for (void *p : pl->K) cudaFree(p);
    for (void *p : pl->TK) cudaFree(p);
    for (auto &b : pl->brackets) {
        cudaEventDestroy(b.a);
        cudaEventDestroy(b.b);
    }
    for (auto e : pl->evpool) cudaEventDestroy(e);
    comm_destroy(pl->comm);
    delete pl;
    return 0;
}

int insb200_comm_unique_id(void *id128) {
    API_BEGIN
    if (!id128) {
        set_error("null argument");
        return 1;
    }
    return comm_unique_id(id128);
    API_END
}
int insb200_plan_local_n(const insb200_plan_t *pl, int N_local[3]) {
    if (!pl || !N_local) return 1;
    for (int a = 0; a < 3; a++) N_local[a] = a < pl->D ? pl->s.N[a] : 1;
    if (pl->comm) N_local[2] = pl->nzl + 2;
    return 0;
}

int insb200_plan_ipc_export(insb200_plan_t *pl, void *blob192) {
    API_BEGIN
    if (!pl || !pl->ps) {
        set_error("ipc export: the plan has no Poisson solver");
        return 1;
    }
    return poisson_ipc_export(pl->ps, blob192);
    API_END
}
int insb200_plan_ipc_import(insb200_plan_t *pl, const void *blobs) {
    API_BEGIN
    if (!pl || !pl->ps) {
        set_error("ipc import: the plan has no Poisson solver");
        return 1;
    }
    return poisson_ipc_import(pl->ps, blobs);
    API_END
}

int insb200_plan_ipc_close(insb200_plan_t *pl) {
    API_BEGIN
    if (pl && pl->ps) poisson_ipc_close(pl->ps);
    return 0;
    API_END
}

int insb200_plan_psolver(const insb200_plan_t *pl) { return pl ? pl->psolver : INSB200_PSOLVER_NONE; }
int insb200_psolver_iterations(const insb200_plan_t *pl) { return pl ? poisson_iterations(pl->ps) : 0; }
long long insb200_plan_workspace_bytes(const insb200_plan_t *pl) { return pl ? pl->bytes : 0; }
long long insb200_launch_count(const insb200_plan_t *pl) { return pl ? pl->launches : 0; }

#define DISPATCH(expr_f, expr_d) (pl->f64 ? (expr_d) : (expr_f))

int insb200_apply_bc_u(insb200_plan_t *pl, void *u, double t, int dudt, void *stream) {
    API_BEGIN(void) t;
    return DISPATCH(launch_bc_u<float>(pl->gf, pl->bc, (float *)u, dudt, L_of(pl, stream)),
                    launch_bc_u<double>(pl->gd, pl->bc, (double *)u, dudt, L_of(pl, stream)));
    API_END
}
int insb200_apply_bc_p(insb200_plan_t *pl, void *p, void *stream) {
    API_BEGIN
    return DISPATCH(launch_bc_p<float>(pl->gf, pl->bc, (float *)p, L_of(pl, stream)),
                    launch_bc_p<double>(pl->gd, pl->bc, (double *)p, L_of(pl, stream)));
    API_END
}
int insb200_apply_bc_temp(insb200_plan_t *pl, void *temp, double t, void *stream) {
    API_BEGIN(void) t;
    return DISPATCH(launch_bc_temp<float>(pl->gf, pl->bc, (float *)temp, L_of(pl, stream)),
                    launch_bc_temp<double>(pl->gd, pl->bc, (double *)temp, L_of(pl, stream)));
    API_END
}
int insb200_set_bc_u_plane(insb200_plan_t *pl, int dim, int side, int comp, const void *plane) {
    if (!pl || dim < 0 || dim >= pl->D || side < 0 || side > 1 || comp < 0 || comp >= pl->D) {
        set_error("bad (dim, side, comp)");
        return 1;
    }
    pl->bc.u_plane[dim][side][comp] = plane;
    return 0;
}

int insb200_rhs(insb200_plan_t *pl, const void *u, const void *temp, void *ku, void *ktemp,
                const insb200_params_t *pr, void *stream) {
    API_BEGIN
    if (pr->boussinesq && (!temp || !ktemp)) {
        set_error("boussinesq forcing needs temp and ktemp");
        return 1;
    }
    return DISPATCH(rhs_t<float>(pl, (const float *)u, (const float *)temp, (float *)ku, (float *)ktemp, pr, 1,
                                 1, 0, stream),
                    rhs_t<double>(pl, (const double *)u, (const double *)temp, (double *)ku, (double *)ktemp,
                                  pr, 1, 1, 0, stream));
    API_END
}
int insb200_convection(insb200_plan_t *pl, void *f, const void *u, void *stream) {
    API_BEGIN
    insb200_params_t pr{};
    return DISPATCH(rhs_t<float>(pl, (const float *)u, nullptr, (float *)f, nullptr, &pr, 1, 0, 1, stream),
                    rhs_t<double>(pl, (const double *)u, nullptr, (double *)f, nullptr, &pr, 1, 0, 1, stream));
    API_END
}
int insb200_diffusion(insb200_plan_t *pl, void *f, const void *u, double viscosity, void *stream) {
    API_BEGIN
    insb200_params_t pr{};
    pr.viscosity = viscosity;
    return DISPATCH(rhs_t<float>(pl, (const float *)u, nullptr, (float *)f, nullptr, &pr, 0, 1, 1, stream),
                    rhs_t<double>(pl, (const double *)u, nullptr, (double *)f, nullptr, &pr, 0, 1, 1, stream));
    API_END
}

int insb200_divergence(insb200_plan_t *pl, void *div, const void *u, void *stream) {
    API_BEGIN
    return DISPATCH(launch_divergence<float>(pl->gf, (const float *)u, (float *)div, false, L_of(pl, stream)),
                    launch_divergence<double>(pl->gd, (const double *)u, (double *)div, false, L_of(pl, stream)));
    API_END
}
int insb200_scalewithvolume(insb200_plan_t *pl, void *p, void *stream) {
    API_BEGIN
    return DISPATCH(launch_scalewithvolume<float>(pl->gf, (float *)p, L_of(pl, stream)),
                    launch_scalewithvolume<double>(pl->gd, (double *)p, L_of(pl, stream)));
    API_END
}
int insb200_pressuregradient(insb200_plan_t *pl, void *G, const void *p, void *stream) {
    API_BEGIN
    return DISPATCH(
        launch_pressuregradient<float>(pl->gf, (const float *)p, (float *)G, (float *)nullptr, L_of(pl, stream)),
        launch_pressuregradient<double>(pl->gd, (const double *)p, (double *)G, (double *)nullptr,
                                        L_of(pl, stream)));
    API_END
}
int insb200_applypressure(insb200_plan_t *pl, void *u, const void *p, void *stream) {
    API_BEGIN
    return DISPATCH(
        launch_pressuregradient<float>(pl->gf, (const float *)p, (float *)nullptr, (float *)u, L_of(pl, stream)),
        launch_pressuregradient<double>(pl->gd, (const double *)p, (double *)nullptr, (double *)u,
                                        L_of(pl, stream)));
    API_END
}

int insb200_psolve(insb200_plan_t *pl, void *p, void *stream) {
    API_BEGIN
    if (!pl->ps) {
        set_error("this plan has no Poisson solver");
        return 1;
    }
    Scope sc(pl, (cudaStream_t)stream, 2);
    INSB_TRY(poisson_solve(pl->ps, p, L_of(pl, stream)));
    if (pl->psolver == INSB200_PSOLVER_CG)  // the reference's CG returns p with its ghosts filled (p += alpha q, q padded)
        return DISPATCH(launch_bc_p<float>(pl->gf, pl->bc, (float *)p, L_of(pl, stream)),
                        launch_bc_p<double>(pl->gd, pl->bc, (double *)p, L_of(pl, stream)));
    return 0;
    API_END
}
int insb200_project(insb200_plan_t *pl, void *u, void *p, void *stream) {
    API_BEGIN
    return DISPATCH(project_t<float>(pl, (float *)u, (float *)p, stream),
                    project_t<double>(pl, (double *)u, (double *)p, stream));
    API_END
}

int insb200_stage_combine(insb200_plan_t *pl, void *dst, const void *base, int nk, const void *const *k, const double *coef,
                          int coef_dtype, int ncomp, void *stream) {
    API_BEGIN
    return DISPATCH(combine_t<float>(pl, (float *)dst, (const float *)base, nk, k, coef, coef_dtype, ncomp, stream),
                    combine_t<double>(pl, (double *)dst, (const double *)base, nk, k, coef, coef_dtype, ncomp, stream));
    API_END
}
int insb200_stage_combine_project(insb200_plan_t *pl, void *u, const void *u_start, void *temp, const void *temp_start,
                                  int nk, const void *const *k, const void *const *ktemp, const double *coef,
                                  int coef_dtype, void *p, double t_stage, void *stream) {
    API_BEGIN
    return DISPATCH(combine_project_t<float>(pl, (float *)u, (const float *)u_start, (float *)temp, (const float *)temp_start,
                                             nk, k, ktemp, coef, coef_dtype, (float *)p, t_stage, stream),
                    combine_project_t<double>(pl, (double *)u, (const double *)u_start, (double *)temp,
                                              (const double *)temp_start, nk, k, ktemp, coef, coef_dtype, (double *)p,
                                              t_stage, stream));
    API_END
}
int insb200_axpy(insb200_plan_t *pl, void *y, double a, const void *x, int coef_dtype, int ncomp, void *stream) {
    API_BEGIN
    const void *k[1] = {x};
    return DISPATCH(combine_t<float>(pl, (float *)y, (const float *)y, 1, k, &a, coef_dtype, ncomp, stream),
                    combine_t<double>(pl, (double *)y, (const double *)y, 1, k, &a, coef_dtype, ncomp, stream));
    API_END
}

int insb200_step_rk(insb200_plan_t *pl, const double *A, const double *c, int s, int coef_dtype, void *u,
                    void *temp, double t, double dt, const insb200_params_t *pr, void *stream) {
    API_BEGIN
    return DISPATCH(step_rk_t<float>(pl, A, c, s, coef_dtype, (float *)u, (float *)temp, t, dt, pr, stream),
                    step_rk_t<double>(pl, A, c, s, coef_dtype, (double *)u, (double *)temp, t, dt, pr, stream));
    API_END
}
int insb200_step_lmwray3(insb200_plan_t *pl, void *u, void *temp, double t, double dt,
                         const insb200_params_t *pr, void *stream) {
    API_BEGIN
    return DISPATCH(step_wray_t<float>(pl, (float *)u, (float *)temp, t, dt, pr, stream),
                    step_wray_t<double>(pl, (double *)u, (double *)temp, t, dt, pr, stream));
    API_END
}

int insb200_step_rk_host(insb200_plan_t *pl, const double *A, const double *c, int s, int coef_dtype,
                         void *u_host, void *temp_host, double t, double dt, int nsteps,
                         const insb200_params_t *pr, void *stream) {
    API_BEGIN
    cudaStream_t st = (cudaStream_t)stream;
    size_t vb = (size_t)pl->ncell * pl->D * pl->esz, sb = (size_t)pl->ncell * pl->esz;
    // device-resident state lives in the k-storage slots (allocated on first use)
    bool bouss = pr->boussinesq != 0;
    if (!pl->HU) INSB_TRY(dev_alloc(pl, &pl->HU, vb));
    if (bouss && !pl->HT) INSB_TRY(dev_alloc(pl, &pl->HT, sb));
    void *dU = pl->HU;
    void *dT = bouss ? pl->HT : nullptr;
    INSB_CUDA(cudaMemcpyAsync(dU, u_host, vb, cudaMemcpyHostToDevice, st));
    if (bouss) INSB_CUDA(cudaMemcpyAsync(dT, temp_host, sb, cudaMemcpyHostToDevice, st));
    for (int n = 0; n < nsteps; n++) {
        INSB_TRY(insb200_step_rk(pl, A, c, s, coef_dtype, dU, dT, t + n * dt, dt, pr, stream));
    }
    INSB_CUDA(cudaMemcpyAsync(u_host, dU, vb, cudaMemcpyDeviceToHost, st));
    if (bouss) INSB_CUDA(cudaMemcpyAsync(temp_host, dT, sb, cudaMemcpyDeviceToHost, st));
    INSB_CUDA(cudaStreamSynchronize(st));
    return 0;
    API_END
}

int insb200_step_rk_host_ensemble(insb200_plan_t *pl, const double *A, const double *c, int s, int coef_dtype,
                                  void *const *u_hosts, void *const *temp_hosts, int nmembers, double t,
                                  double dt, int nsteps, const insb200_params_t *pr, void *stream) {
    API_BEGIN
    if (nmembers < 1 || !u_hosts) {
        set_error("step_rk_host_ensemble: need at least one member");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t vb = (size_t)pl->ncell * pl->D * pl->esz, sb = (size_t)pl->ncell * pl->esz;
    const bool bouss = pr->boussinesq != 0;
    if (bouss && !temp_hosts) {
        set_error("boussinesq forcing needs a temperature field");
        return 1;
    }
    if (!pl->HU) INSB_TRY(dev_alloc(pl, &pl->HU, vb));
    if (!pl->HU2) INSB_TRY(dev_alloc(pl, &pl->HU2, vb));
    if (nmembers > 2 && !pl->HU3) INSB_TRY(dev_alloc(pl, &pl->HU3, vb));
    if (bouss && !pl->HT) INSB_TRY(dev_alloc(pl, &pl->HT, sb));
    if (bouss && !pl->HT2) INSB_TRY(dev_alloc(pl, &pl->HT2, sb));
    if (bouss && nmembers > 2 && !pl->HT3) INSB_TRY(dev_alloc(pl, &pl->HT3, sb));
    // a slot is busy for H2D + steps + D2H; with the copies near the PCIe rate (2 x 29 ms for 1.6 GB each way) and a
    // 14 ms step, two slots cap the pipeline at 37 ms per member, three let it run at the rate of one copy direction
    const int nslot = nmembers > 2 ? 3 : 2;
    void *slotU[3] = {pl->HU, pl->HU2, pl->HU3}, *slotT[3] = {pl->HT, pl->HT2, pl->HT3};
    if (!pl->s_in) INSB_CUDA(cudaStreamCreateWithFlags(&pl->s_in, cudaStreamNonBlocking));
    if (!pl->s_out) INSB_CUDA(cudaStreamCreateWithFlags(&pl->s_out, cudaStreamNonBlocking));
    // pipeline over the members: H2D of the next members (s_in) and D2H of the previous ones (s_out) run while
    // member m is being stepped on the caller's stream; three device slots (two for a pair of members)
    std::vector<cudaEvent_t> ev(3 * (size_t)nmembers);
    for (auto &e : ev) INSB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = 0;
    {
        cudaEvent_t start;
        INSB_CUDA(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
        INSB_CUDA(cudaEventRecord(start, st));  // work already queued on the caller's stream comes first
        INSB_CUDA(cudaStreamWaitEvent(pl->s_in, start, 0));
        INSB_CUDA(cudaEventDestroy(start));
    }
    for (int m = 0; m < nmembers && rc == 0; m++) {
        void *dU = slotU[m % nslot];
        void *dT = bouss ? slotT[m % nslot] : nullptr;
        cudaEvent_t in_done = ev[3 * m], comp_done = ev[3 * m + 1], out_done = ev[3 * m + 2];
        if (m >= nslot) INSB_CUDA(cudaStreamWaitEvent(pl->s_in, ev[3 * (m - nslot) + 2], 0));  // slot free again
        INSB_CUDA(cudaMemcpyAsync(dU, u_hosts[m], vb, cudaMemcpyHostToDevice, pl->s_in));
        if (bouss) INSB_CUDA(cudaMemcpyAsync(dT, temp_hosts[m], sb, cudaMemcpyHostToDevice, pl->s_in));
        INSB_CUDA(cudaEventRecord(in_done, pl->s_in));
        INSB_CUDA(cudaStreamWaitEvent(st, in_done, 0));
        for (int n = 0; n < nsteps && rc == 0; n++)
            rc = insb200_step_rk(pl, A, c, s, coef_dtype, dU, dT, t + n * dt, dt, pr, stream);
        INSB_CUDA(cudaEventRecord(comp_done, st));
        INSB_CUDA(cudaStreamWaitEvent(pl->s_out, comp_done, 0));
        INSB_CUDA(cudaMemcpyAsync(u_hosts[m], dU, vb, cudaMemcpyDeviceToHost, pl->s_out));
        if (bouss) INSB_CUDA(cudaMemcpyAsync(temp_hosts[m], dT, sb, cudaMemcpyDeviceToHost, pl->s_out));
        INSB_CUDA(cudaEventRecord(out_done, pl->s_out));
    }
    cudaError_t e1 = cudaStreamSynchronize(pl->s_out), e2 = cudaStreamSynchronize(st);
    for (auto &e : ev) cudaEventDestroy(e);
    if (rc) return rc;
    INSB_CUDA(e1);
    INSB_CUDA(e2);
    return 0;
    API_END
}

// solve_unsteady (src/solver.jl:77-149): the whole time loop inside the library.
int insb200_solve_unsteady(insb200_plan_t *pl, int method, const double *A, const double *c, int s, int coef_dtype,
                           void *u, void *temp, double tstart, double tend, double dt, double dt_min, double cfl,
                           int n_adapt_dt, int time_dtype, const insb200_params_t *pr, int callback_every,
                           insb200_step_callback_t cb, void *user, int use_graph, double *t_out, long long *n_out,
                           void *stream) {
    API_BEGIN
    if (method != 0 && method != 1) {
        set_error("solve_unsteady: method must be 0 (explicit Runge-Kutta tableau) or 1 (LMWray3)");
        return 1;
    }
    if (method == 0 && (!A || !c)) {
        set_error("solve_unsteady: the Runge-Kutta method needs its tableau");
        return 1;
    }
    const bool tf32 = time_dtype == INSB200_F32;
    // t, Δt arithmetic in the caller's time type (the reference keeps typeof(tstart), typeof(Δt))
    auto R = [&](double v) { return tf32 ? (double)(float)v : v; };
    // the limits live in the time type too: a Float32 clock can never reach a Float64 tend that is not representable
    // (t < tend would hold forever with steps that round to nothing)
    tstart = R(tstart);
    tend = R(tend);
    auto step = [&](void *st, double t, double h) -> int {
        if (method == 1) return insb200_step_lmwray3(pl, u, temp, t, h, pr, st);
        return insb200_step_rk(pl, A, c, s, coef_dtype, u, temp, t, h, pr, st);
    };
    // chained steps on the fused periodic path: between two steps the state stays lazy (no explicit projection into
    // u); u is materialised before a call-back, before a CFL evaluation and at the end of the loop
    const bool can_chain = getenv("INSB200_NO_CHAINED_STEPS") == nullptr && fused_applicable(pl, pr);
    auto materialise = [&](void *st) -> int {
        if (!pl->chain_valid) return 0;
        pl->chain_valid = false;
        const int sl = pl->chain_slot;
        return pl->f64 ? fused_finish<double>(pl, (const double *)pl->FV[sl], (double *)u, st)
                       : fused_finish<float>(pl, (const float *)pl->FV[sl], (float *)u, st);
    };
    struct ChainReset {  // whatever happens, the plan leaves the loop without chaining
        insb200_plan *p;
        ~ChainReset() { p->chain_in = p->chain_out = p->chain_valid = false; }
    } chain_reset{pl};
    auto cb_due = [&](long long nn) { return cb && callback_every > 0 && nn % callback_every == 0; };
    double t = tstart;
    long long n = 0;
    int rc = 0;
    bool stop = false;
    auto notify = [&]() {
        if (cb && callback_every > 0 && n % callback_every == 0 && cb(user, n, t) != 0) stop = true;
    };
    if (dt <= 0) {  // adaptive: CFL every n_adapt_dt steps, last step clipped to tend (:98-122)
        if (n_adapt_dt < 1) n_adapt_dt = 1;
        double h = 0;
        while (t < tend && !stop) {
            if (n % n_adapt_dt == 0) {
                double prop;
                INSB_TRY(insb200_cfl(pl, u, pr, &prop, stream));
                h = pl->f64 ? cfl * prop : (double)((float)cfl * (float)prop);  // cfl and the proposal are eltype(u)
                if (dt_min > 0) h = std::fmax(h, dt_min);
                if (h < 1e-10 || h > 100 || std::isnan(h)) {
                    set_error("Proposed time step " + std::to_string(h) + " is out of bounds. Stopping simulation.");
                    rc = INSB200_STOPPED_BAD_DT;
                    break;
                }
            }
            const double hs = std::fmin(h, R(tend - t));
            const bool lastone = !(R(t + hs) < tend);
            pl->chain_in = can_chain;
            pl->chain_out = can_chain && !lastone && !cb_due(n + 1) && (n + 1) % n_adapt_dt != 0;
            INSB_TRY(step(stream, t, hs));
            t = R(t + hs);
            n++;
            notify();
        }
        INSB_TRY(materialise(stream));
    } else {  // fixed: Δt rounded so that (tend - tstart) / Δt is an integer (:123-132)
        const long long nstep = std::llround(R(tend - tstart) / dt);
        if (nstep < 1) {
            set_error("solve_unsteady: (tend - tstart) / dt rounds to zero steps");
            return 1;
        }
        const double h = R(R(tend - tstart) / (double)nstep);
        cudaStream_t user_st = (cudaStream_t)stream;
        // (psolver_cg reads its residual back while iterating: not capturable)
        const bool graph = use_graph && !pl->comm && !pl->prof && nstep > 2 && pl->psolver != INSB200_PSOLVER_CG;
        if (!graph) {
            for (long long it = 0; it < nstep && !stop; it++) {
                pl->chain_in = can_chain;
                pl->chain_out = can_chain && it < nstep - 1 && !cb_due(n + 1);
                INSB_TRY(step(stream, t, h));
                t = R(t + h);
                n++;
                notify();
            }
            INSB_TRY(materialise(stream));
        } else {
            // one eager step (allocations, function attributes), then the step is captured once and
            // replayed: the launch-bound small grids (config 1: 256^2) stop paying per-kernel launches.
            // The legacy default stream cannot be captured: run on a private stream ordered after it.
            if (!pl->s_graph) INSB_CUDA(cudaStreamCreateWithFlags(&pl->s_graph, cudaStreamNonBlocking));
            cudaEvent_t ev;
            INSB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            auto handoff = [&](cudaStream_t from, cudaStream_t to) -> int {
                INSB_CUDA(cudaEventRecord(ev, from));
                INSB_CUDA(cudaStreamWaitEvent(to, ev, 0));
                return 0;
            };
            // chained replay: every step must use the buffers in the same roles (an even number of stages) and nothing
            // may look at u in between (no call-backs); the last step runs eagerly and materialises u
            const bool gchain = can_chain && method == 0 && s % 2 == 0 && !(cb && callback_every > 0) && nstep > 3;
            pl->chain_in = false;
            pl->chain_out = gchain;
            INSB_TRY(step(stream, t, h));
            t = R(t + h);
            n++;
            notify();
            INSB_TRY(handoff(user_st, pl->s_graph));
            pl->chain_in = gchain;
            const long long l0 = pl->launches;
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ge = nullptr;
            INSB_CUDA(cudaStreamBeginCapture(pl->s_graph, cudaStreamCaptureModeThreadLocal));
            int src = step((void *)pl->s_graph, t, h);
            cudaError_t ce = cudaStreamEndCapture(pl->s_graph, &g);
            const long long per_step = pl->launches - l0;
            pl->launches = l0;
            if (src || ce != cudaSuccess || !g) {
                if (g) cudaGraphDestroy(g);
                cudaEventDestroy(ev);
                if (!src) set_error(std::string("CUDA graph capture of the time step failed: ") + cudaGetErrorString(ce));
                return 1;
            }
            ce = cudaGraphInstantiate(&ge, g, 0);
            if (ce != cudaSuccess) {
                cudaGraphDestroy(g);
                cudaEventDestroy(ev);
                set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
                return 1;
            }
            pl->chain_valid = gchain;  // (the capture did not execute: the eager first step left the lazy state)
            for (long long it = 1; it < nstep && !stop; it++) {
                if (gchain && it == nstep - 1) {  // the last step: eager, into u
                    pl->chain_out = false;
                    rc = step((void *)pl->s_graph, t, h);
                    if (rc) break;
                    t = R(t + h);
                    n++;
                    break;
                }
                ce = cudaGraphLaunch(ge, pl->s_graph);
                if (ce != cudaSuccess) break;
                pl->launches += per_step;
                t = R(t + h);
                n++;
                if (cb && callback_every > 0 && n % callback_every == 0) {
                    rc = handoff(pl->s_graph, user_st);  // the callback sees the state on the caller's stream
                    if (rc) break;
                    notify();
                }
            }
            int hrc = handoff(pl->s_graph, user_st);
            cudaGraphExecDestroy(ge);
            cudaGraphDestroy(g);
            cudaEventDestroy(ev);
            INSB_CUDA(ce);
            if (rc || hrc) return 1;
            INSB_TRY(materialise(stream));
        }
    }
    if (t_out) *t_out = t;
    if (n_out) *n_out = n;
    return rc;
    API_END
}

int insb200_right_hand_side(insb200_plan_t *pl, void *dudt, const void *u, double t,
                            const insb200_params_t *pr, void *stream) {
    API_BEGIN
    return DISPATCH(rhs_sciml_t<float>(pl, (float *)dudt, (const float *)u, t, pr, stream),
                    rhs_sciml_t<double>(pl, (double *)dudt, (const double *)u, t, pr, stream));
    API_END
}

int insb200_cfl(insb200_plan_t *pl, const void *u, const insb200_params_t *pr, double *dt_out, void *stream) {
    API_BEGIN
    Launch L = L_of(pl, stream);
    INSB_TRY(DISPATCH(launch_cfl<float>(pl->gf, pl->Iu_loc, (const float *)u, pl->scratch, L),
                      launch_cfl<double>(pl->gd, pl->Iu_loc, (const double *)u, pl->scratch, L)));
    if (pl->comm) INSB_TRY(comm_reduce_scratch(pl->comm, pl->scratch, true, L.stream));
    unsigned long long h[2];
    INSB_CUDA(cudaMemcpyAsync(h, pl->scratch, sizeof(h), cudaMemcpyDeviceToHost, L.stream));
    INSB_CUDA(cudaStreamSynchronize(L.stream));
    double conv;
    std::memcpy(&conv, &h[0], 8);
    if (h[1]) conv = std::nan("");
    // values were computed in T; the diffusive limits below too (solver.jl:157-166, 188-197)
    double dt = conv;
    const int D = pl->D;
    for (int a = 0; a < D; a++) {
        double d;
        if (pl->f64)
            d = pl->dmin_u[a] * pl->dmin_u[a] / pr->viscosity / (2 * D);
        else
            d = (double)((float)pl->dmin_u[a] * (float)pl->dmin_u[a] / (float)pr->viscosity / (float)(2 * D));
        dt = std::isnan(dt) ? dt : std::fmin(dt, d);
    }
    if (pr->boussinesq) {
        for (int a = 0; a < D; a++) {
            double d;
            if (pl->f64)
                d = pl->dmin_p[a] * pl->dmin_p[a] / pr->conductivity / (2 * D);
            else
                d = (double)((float)pl->dmin_p[a] * (float)pl->dmin_p[a] / (float)pr->conductivity /
                             (float)(2 * D));
            dt = std::isnan(dt) ? dt : std::fmin(dt, d);
        }
    }
    *dt_out = dt;
    return 0;
    API_END
}

int insb200_eddy_viscosity_closure(insb200_plan_t *pl, int model, double C, void *f, const void *u, void *stream) {
    API_BEGIN
    size_t sb = (size_t)pl->ncell * pl->esz;
    const int *Ng = pl->f64 ? pl->gd.N : pl->gf.N;
    if (!pl->LES_G && les_needs_tensor(pl->D, Ng, pl->bc, u, f, nullptr, nullptr)) INSB_TRY(dev_alloc(pl, &pl->LES_G, 9 * sb));
    if (!pl->LES_visc) INSB_TRY(dev_alloc(pl, &pl->LES_visc, sb));
    Launch L = L_of(pl, stream);
    return DISPATCH(launch_les_closure<float>(pl->gf, pl->bc, model, C, (const float *)u, (float *)pl->LES_G,
                                              (float *)pl->LES_visc, (float *)f, (float *)nullptr, 0.0, (float *)nullptr,
                                              0.0, false, L),
                    launch_les_closure<double>(pl->gd, pl->bc, model, C, (const double *)u, (double *)pl->LES_G,
                                               (double *)pl->LES_visc, (double *)f, (double *)nullptr, 0.0,
                                               (double *)nullptr, 0.0, true, L));
    API_END
}

// ---- diagnostics (src/operators.jl:853-1066), diag.cu --------------------------------
static void local_ip(const insb200_plan *pl, int Ip[3][2]) {
    std::memcpy(Ip, pl->Ip, sizeof(pl->Ip));
    if (pl->comm) Ip[2][0] = 1, Ip[2][1] = pl->nzl + 1;
}

static int diag_call(insb200_plan_t *pl, int op, const void *in, void *out, void *stream) {
    int Ip[3][2];
    local_ip(pl, Ip);
    return DISPATCH(launch_diag<float>(pl->gf, Ip, op, (const float *)in, (float *)out, L_of(pl, stream)),
                    launch_diag<double>(pl->gd, Ip, op, (const double *)in, (double *)out, L_of(pl, stream)));
}

int insb200_vorticity(insb200_plan_t *pl, void *omega, const void *u, void *stream) {
    API_BEGIN
    return diag_call(pl, 0, u, omega, stream);
    API_END
}

int insb200_interpolate_u_p(insb200_plan_t *pl, void *up, const void *u, void *stream) {
    API_BEGIN
    return diag_call(pl, 1, u, up, stream);
    API_END
}

int insb200_interpolate_omega_p(insb200_plan_t *pl, void *omegap, const void *omega, void *stream) {
    API_BEGIN
    return diag_call(pl, 2, omega, omegap, stream);
    API_END
}

int insb200_qcrit(insb200_plan_t *pl, void *q, const void *u, void *stream) {
    API_BEGIN
    return diag_call(pl, 3, u, q, stream);
    API_END
}

int insb200_kinetic_energy(insb200_plan_t *pl, void *ke, const void *u, int interpolate_first, void *stream) {
    API_BEGIN
    return diag_call(pl, interpolate_first ? 5 : 4, u, ke, stream);
    API_END
}

int insb200_total_kinetic_energy(insb200_plan_t *pl, const void *u, int interpolate_first, double *out,
                                 void *stream) {
    API_BEGIN
    constexpr int NPART = 148 * 4;
    if (!pl->ke_partial) INSB_CUDA(cudaMalloc((void **)&pl->ke_partial, NPART * sizeof(double)));
    int Ip[3][2];
    local_ip(pl, Ip);
    Launch L = L_of(pl, stream);
    INSB_TRY(DISPATCH(
        launch_total_ke<float>(pl->gf, Ip, (const float *)u, interpolate_first != 0, pl->ke_partial, NPART, pl->scratch, L),
        launch_total_ke<double>(pl->gd, Ip, (const double *)u, interpolate_first != 0, pl->ke_partial, NPART, pl->scratch, L)));
    if (pl->comm) INSB_TRY(comm_sum_double(pl->comm, pl->scratch, L.stream));
    INSB_CUDA(cudaMemcpyAsync(out, pl->scratch, sizeof(double), cudaMemcpyDeviceToHost, L.stream));
    INSB_CUDA(cudaStreamSynchronize(L.stream));
    return 0;
    API_END
}

int insb200_random_field(insb200_plan_t *pl, void *u, const double *shell_energy, int nshell, int fill_noise,
                         unsigned long long seed, double t, void *stream) {
    API_BEGIN
    if (!pl->ps) {
        set_error("random_field needs a plan with psolver_spectral");
        return 1;
    }
    Launch L = L_of(pl, stream);
    INSB_TRY(poisson_random_field(pl->ps, u, pl->ncell, shell_energy, nshell, fill_noise, seed, pl->s.xlims, L));
    (void)t;
    return DISPATCH(launch_bc_u<float>(pl->gf, pl->bc, (float *)u, 0, L), launch_bc_u<double>(pl->gd, pl->bc, (double *)u, 0, L));
    API_END
}

int insb200_energy_spectrum(insb200_plan_t *pl, const void *u, int kmax, double *ehat, void *stream) {
    API_BEGIN
    if (!pl->ps) {
        set_error("energyspectrum needs a plan with psolver_spectral");
        return 1;
    }
    if (kmax < 0) {  // spectral_stuff default: minimum(Np) / 2 (spectral.jl:36)
        kmax = pl->s.N[0] - 2;
        for (int a = 0; a < pl->D; a++) kmax = std::min(kmax, (pl->Ip[a][1] - pl->Ip[a][0]) / 2);
    }
    double *dev = nullptr;
    INSB_CUDA(cudaMalloc((void **)&dev, (size_t)(kmax + 1) * sizeof(double)));
    Launch L = L_of(pl, stream);
    int rc = poisson_energy_spectrum(pl->ps, u, pl->ncell, kmax, dev, L);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(ehat, dev, (size_t)(kmax + 1) * sizeof(double), cudaMemcpyDeviceToHost, L.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
        if (e != cudaSuccess) {
            set_error(std::string("energyspectrum: ") + cudaGetErrorString(e));
            rc = 1;
        }
    }
    cudaFree(dev);
    return rc;
    API_END
}

int insb200_spectral_moments(insb200_plan_t *pl, const void *u, double *moments, void *stream) {
    API_BEGIN
    if (!pl->ps) {
        set_error("turbulence_statistics needs a plan with psolver_spectral");
        return 1;
    }
    double lengths[3] = {0.0, 0.0, 0.0};
    for (int a = 0; a < pl->D; a++) lengths[a] = pl->s.xlims[a][1] - pl->s.xlims[a][0];
    double *dev = nullptr;
    INSB_CUDA(cudaMalloc((void **)&dev, 3 * sizeof(double)));
    Launch L = L_of(pl, stream);
    int rc = poisson_spectral_moments(pl->ps, u, pl->ncell, lengths, dev, L);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(moments, dev, 3 * sizeof(double), cudaMemcpyDeviceToHost, L.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
        if (e != cudaSuccess) {
            set_error(std::string("turbulence_statistics: ") + cudaGetErrorString(e));
            rc = 1;
        }
    }
    cudaFree(dev);
    return rc;
    API_END
}

int insb200_maxabs(insb200_plan_t *pl, const void *u, double *out, void *stream) {
    API_BEGIN
    Launch L = L_of(pl, stream);
    long long n = pl->ncell * pl->D;
    INSB_TRY(DISPATCH(launch_maxabs<float>((const float *)u, n, pl->scratch, L),
                      launch_maxabs<double>((const double *)u, n, pl->scratch, L)));
    if (pl->comm) INSB_TRY(comm_reduce_scratch(pl->comm, pl->scratch, false, L.stream));
    unsigned long long h[2];
    INSB_CUDA(cudaMemcpyAsync(h, pl->scratch, sizeof(h), cudaMemcpyDeviceToHost, L.stream));
    INSB_CUDA(cudaStreamSynchronize(L.stream));
    double v;
    std::memcpy(&v, &h[0], 8);
    *out = h[1] ? std::nan("") : v;
    return 0;
    API_END
}

// Debugging aid (not part of the reference-facing surface): copy one padded internal array of the fused
// path to the host.  which: 0, 1 = stage velocities, 2 = accumulator, 3 = pressure, 4 = scaled divergence;
// dims receives {NPX, NPY, NPZ, components}.
int insb200_debug_fused_copy(insb200_plan_t *pl, int which, void *host, long long host_bytes, int dims[4]) {
    API_BEGIN
    if (!pl->fused_ready) {
        set_error("the fused path has not run on this plan");
        return 1;
    }
    void *src[5] = {pl->FV[0], pl->FV[1], pl->FACC, pl->FP, pl->FD};
    if (which < 0 || which > 4) return 1;
    const int nc = which < 3 ? 3 : 1;
    dims[0] = pl->fctx.L.NPX, dims[1] = pl->fctx.L.NPY, dims[2] = pl->fctx.L.NPZ, dims[3] = nc;
    const size_t bytes = (size_t)pl->fctx.L.cs * nc * pl->esz;
    if (host && (long long)bytes <= host_bytes) {
        INSB_CUDA(cudaDeviceSynchronize());
        INSB_CUDA(cudaMemcpy(host, src[which], bytes, cudaMemcpyDeviceToHost));
    }
    return 0;
    API_END
}

int insb200_profile_enable(insb200_plan_t *pl, int on) {
    pl->prof = on != 0;
    return 0;
}
static int profile_collect(insb200_plan_t *pl) {
    INSB_CUDA(cudaDeviceSynchronize());
    for (auto &b : pl->brackets) {
        float t = 0;
        cudaEventElapsedTime(&t, b.a, b.b);
        pl->prof_ms[b.group] += t;
        pl->prof_cnt[b.group] += 1;
        pl->evpool.push_back(b.a);
        pl->evpool.push_back(b.b);
    }
    pl->brackets.clear();
    return 0;
}
int insb200_profile_read_ex(insb200_plan_t *pl, int n, double *ms, long long *counts) {
    API_BEGIN
    INSB_TRY(profile_collect(pl));
    for (int i = 0; i < NPROF; i++) {
        if (i < n) {
            ms[i] = pl->prof_ms[i];
            counts[i] = pl->prof_cnt[i];
        }
        pl->prof_ms[i] = 0;
        pl->prof_cnt[i] = 0;
    }
    return 0;
    API_END
}
int insb200_profile_read(insb200_plan_t *pl, double ms[5], long long counts[5]) {
    return insb200_profile_read_ex(pl, 5, ms, counts);
}

}  // extern "C"
