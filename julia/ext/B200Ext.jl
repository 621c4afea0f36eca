"""
    B200Ext

Package extension that routes the per-time-step hot path of IncompressibleNavierStokes
(right-hand side, Runge-Kutta stage combine, ghost fills, pressure projection) to
`libinsb200.so`, the hand-written sm_100a CUDA backend, through `ccall`.

Trigger: `CUDA` (like `ext/CUDSSExt.jl`).  Add to the reference's `Project.toml`:

    [extensions]
    B200Ext = ["CUDA"]

Usage -- the only change to a user script is the pressure solver:

    setup   = Setup(; x, boundary_conditions, backend = CUDABackend())
    psolver = psolver_b200(setup)                    # or psolver_b200(setup; kind = :transform)
    solve_unsteady(; setup, psolver, method = RKMethods.RK44(; T), ...)

`psolver_b200` returns a callable `B200PSolver` (it solves in place like the closures of
`src/pressure.jl`), and because the stepper NamedTuple carries the psolver's type,
`timestep!`, `project!` and `propose_timestep` dispatch to the fused library calls below.
Everything the extension does not override (processors, `Observable` updates, `Array(u)`,
user-defined `force!`) keeps working on the same `CuArray`s.

This file cannot be executed in the build container (no Julia); the Python/ctypes harness in
`incompressiblenavierstokes.jl_b200/insb200/` binds the very same symbols and is what the
tests drive.  See INTEGRATION.md.
"""
module B200Ext

using CUDA
using IncompressibleNavierStokes
using IncompressibleNavierStokes:
    PeriodicBC,
    DirichletBC,
    SymmetricBC,
    PressureBC,
    ExplicitRungeKuttaMethod,
    LMWray3,
    navierstokes!,
    boussinesq!,
    create_stepper

const INS = IncompressibleNavierStokes
const libinsb200 = get(ENV, "INSB200_LIB", "libinsb200.so")

# ---- C structs of include/insb200.h (field order and types must match exactly) ---------
struct CSetup
    D::Cint
    N::NTuple{3,Cint}
    dtype::Cint
    bc_u::NTuple{6,Cint}              # [dim][side]
    has_temp::Cint
    bc_temp::NTuple{6,Cint}
    Delta::NTuple{3,Ptr{Cvoid}}
    DeltaU::NTuple{3,Ptr{Cvoid}}
    A1::NTuple{3,Ptr{Cvoid}}
    A2::NTuple{3,Ptr{Cvoid}}
    xlims::NTuple{6,Cdouble}
    bc_u_value::NTuple{18,Cdouble}    # [dim][side][component]
    bc_temp_value::NTuple{6,Cdouble}
    psolver::Cint
    tridiagonal_dir::Cint
    eig::NTuple{3,Ptr{Cvoid}}
    device::Cint
    rank::Cint
    world::Cint
    nccl_id::Ptr{Cvoid}
    cg_abstol::Cdouble              # psolver_cg keyword arguments, 0 = reference defaults
    cg_reltol::Cdouble
    cg_maxiter::Clonglong
end

struct CParams
    viscosity::Cdouble
    conductivity::Cdouble
    gravity::Cdouble
    gdir::Cint
    dodissipation::Cint
    boussinesq::Cint
    bodyforce::NTuple{3,Cdouble}
    eddy_model::Cint            # 0 none, 1 Smagorinsky, 2 WALE, 3 Vreman, 4 QR
    eddy_C::Cdouble
end

eddycode(::Nothing) = (Cint(0), 0.0)
eddycode(e::INS.Smagorinsky) = (Cint(1), Float64(e.C))
eddycode(e::INS.WALE) = (Cint(2), Float64(e.C))
eddycode(e::INS.Vreman) = (Cint(3), Float64(e.C))
eddycode(e::INS.QR) = (Cint(4), Float64(e.C))

bccode(::PeriodicBC) = Cint(0)
bccode(::DirichletBC) = Cint(1)
bccode(::SymmetricBC) = Cint(2)
bccode(::PressureBC) = Cint(3)

lasterror() = unsafe_string(ccall((:insb200_last_error, libinsb200), Cstring, ()))
check(status) = status == 0 ? nothing : error("libinsb200: " * lasterror())

# DirichletBC(nothing) -> no slip, DirichletBC((a, b, c)) -> constants; functions are
# evaluated on the boundary planes by `refresh_bc!` below.
bcconst(bc, α) = 0.0
bcconst(bc::DirichletBC{<:Tuple}, α) = Float64(bc.u[α])
bcconst(bc::DirichletBC{<:Number}, α) = Float64(bc.u)
isfuncbc(bc) = bc isa DirichletBC && bc.u isa Function
hasfuncbc(setup) = any(isfuncbc, Iterators.flatten(setup.boundary_conditions.u))

# Temperature: DirichletBC() (u === nothing) is zero like in the reference (boundary_conditions.jl:237-245);
# function-valued temperature planes have no entry point yet, so they are refused instead of silently zeroed.
tempconst(bc) = 0.0
tempconst(bc::DirichletBC{Nothing}) = 0.0
tempconst(bc::DirichletBC{<:Number}) = Float64(bc.u)
tempconst(bc::DirichletBC{<:Function}) =
    error("libinsb200: function-valued DirichletBC for the temperature is not supported by the B200 backend")

"Handle owning an `insb200_plan_t*` plus the host arrays the plan was created from."
mutable struct B200PSolver{S} <: Function
    plan::Ptr{Cvoid}
    setup::S
    kind::Symbol
    bcplanes::Dict{NTuple{3,Int},Any}       # device planes of function-valued DirichletBC (owned here: the
    bchost::Dict{NTuple{3,Int},Any}         # finalizer releases them with the plan), their host copies,
    bcstate::Tuple{Float64,Bool}            # the (t, dudt) they were evaluated at
    bc_time_dependent::Bool                 # functions depend on t: timestep! composes the step stage by stage
    cache::Dict{Symbol,Any}                 # statestart / k / p of the stage-by-stage composition
    function B200PSolver(plan, setup::S, kind; bc_time_dependent) where {S}
        ps = new{S}(plan, setup, kind, Dict{NTuple{3,Int},Any}(), Dict{NTuple{3,Int},Any}(), (NaN, false),
            bc_time_dependent, Dict{Symbol,Any}())
        finalizer(ps) do x
            x.plan == C_NULL || ccall((:insb200_plan_destroy, libinsb200), Cint, (Ptr{Cvoid},), x.plan)
            x.plan = C_NULL
        end
    end
end

# The stock drivers call `propose_timestep(force!, state, setup, params)` without the psolver
# (src/solver.jl:111): the plan is found through the setup it was made for.  A setup is an immutable NamedTuple
# (it cannot key a WeakKeyDict), so the registry is keyed by the identity of its first grid-spacing array and holds
# the psolver weakly -- a collected psolver leaves a dead reference that the next registration overwrites.
const PLAN_OF_SETUP = Dict{UInt,WeakRef}()
setupkey(setup) = objectid(setup.Δ[1])
function psolver_of(setup)
    w = get(PLAN_OF_SETUP, setupkey(setup), nothing)
    isnothing(w) && return nothing
    ps = w.value
    (ps isa B200PSolver && ps.setup === setup) ? ps : nothing
end

const KINDS = (; spectral = Cint(0), transform = Cint(1), tridiagonal = Cint(2), default = Cint(3), cg = Cint(4))

"""
    psolver_b200(setup; kind = :default, dir = nothing)

B200 replacement of `psolver_spectral` / `psolver_transform` / `psolver_tridiagonal` /
`default_psolver` (`src/pressure.jl:58-78, 215-545`).
"""
function INS.psolver_b200(setup; kind = :default, dir = nothing, rank = 0, world = 1, nccl_id = C_NULL,
        abstol = 0.0, reltol = 0.0, maxiter = 0, bc_time_dependent = nothing)
    (; Δ, Δu, A_stag, N, xlims, boundary_conditions) = setup
    D = length(N)
    T = eltype(Δ[1])
    T in (Float32, Float64) || error("libinsb200 supports Float32 and Float64")
    # 1-D tables are DATA: copy them to the host exactly as Setup computed them
    host = (; Δ = Array.(Δ), Δu = Array.(Δu), A1 = map(a -> Array(a[1]), A_stag), A2 = map(a -> Array(a[2]), A_stag))
    pad3(f, fill) = ntuple(d -> d <= D ? f(d) : fill, 3)
    bcu = boundary_conditions.u
    hastemp = haskey(boundary_conditions, :temp)
    bct = hastemp ? boundary_conditions.temp : bcu
    flat6(f) = ntuple(i -> (d = (i - 1) ÷ 2 + 1; s = (i - 1) % 2 + 1; d <= D ? f(d, s) : zero(f(1, 1))), 6)
    GC.@preserve host begin
        cs = CSetup(
            D,
            pad3(d -> Cint(N[d]), Cint(1)),
            T == Float64 ? Cint(1) : Cint(0),
            flat6((d, s) -> bccode(bcu[d][s])),
            Cint(hastemp),
            flat6((d, s) -> bccode(bct[d][s])),
            pad3(d -> Ptr{Cvoid}(pointer(host.Δ[d])), C_NULL),
            pad3(d -> Ptr{Cvoid}(pointer(host.Δu[d])), C_NULL),
            pad3(d -> Ptr{Cvoid}(pointer(host.A1[d])), C_NULL),
            pad3(d -> Ptr{Cvoid}(pointer(host.A2[d])), C_NULL),
            flat6((d, s) -> Float64(xlims[d][s])),
            ntuple(18) do i
                d = (i - 1) ÷ 6 + 1
                s = ((i - 1) ÷ 3) % 2 + 1
                α = (i - 1) % 3 + 1
                d <= D && α <= D ? bcconst(bcu[d][s], α) : 0.0
            end,
            flat6((d, s) -> hastemp ? tempconst(bct[d][s]) : 0.0),
            KINDS[kind],
            isnothing(dir) ? Cint(-1) : Cint(dir - 1),
            (C_NULL, C_NULL, C_NULL),
            Cint(CUDA.deviceid(CUDA.device())),
            Cint(rank),
            Cint(world),
            nccl_id,
            Float64(abstol),
            Float64(reltol),
            Clonglong(maxiter),
        )
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:insb200_plan_create, libinsb200), Cint, (Ref{CSetup}, Ref{Ptr{Cvoid}}), cs, out))
    end
    resolved = ccall((:insb200_plan_psolver, libinsb200), Cint, (Ptr{Cvoid},), out[])
    resolved == -1 && error("no Poisson solver on the B200 path for this setup")
    # function-valued DirichletBC: time dependent unless the caller says otherwise (a closure cannot be
    # inspected); `bc_time_dependent = false` keeps the fused one-call steppers and the CUDA-graph replay
    td = isnothing(bc_time_dependent) ? hasfuncbc(setup) : bc_time_dependent
    ps = B200PSolver(out[], setup, (:spectral, :transform, :tridiagonal, :default, :cg)[resolved+1]; bc_time_dependent = td)
    PLAN_OF_SETUP[setupkey(setup)] = WeakRef(ps)
    hasfuncbc(setup) && refresh_bc!(ps, 0)      # the planes must exist before the first ghost fill
    ps
end

stream() = Ptr{Cvoid}(CUDA.stream().handle)
devptr(a::CuArray) = Ptr{Cvoid}(UInt(pointer(a)))
devptr(::Nothing) = C_NULL

# poisson!(psolver, p) = psolver(p)   (src/pressure.jl:22)
function (ps::B200PSolver)(p::CuArray)
    check(ccall((:insb200_psolve, libinsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), ps.plan, devptr(p), stream()))
    p
end

# project!(u, setup; psolver, p)   (src/pressure.jl:42-55)
function INS.project!(u::CuArray, setup; psolver::B200PSolver, p::CuArray)
    check(ccall((:insb200_project, libinsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
        psolver.plan, devptr(u), devptr(p), stream()))
    u
end

function cparams(force!, params)
    ν = Float64(params.viscosity)
    if force! === boussinesq!
        CParams(ν, params.conductivity, params.gravity, Cint(params.gdir - 1), Cint(params.dodissipation), Cint(1),
            (0.0, 0.0, 0.0), Cint(0), 0.0)
    else
        bf = get(params, :bodyforce, (0.0, 0.0, 0.0))
        # `eddyviscosity = WALE(C)` etc.: navierstokes! + eddy_viscosity_closure! (examples/ChannelFlow.jl:177-182)
        CParams(ν, 0.0, 0.0, Cint(0), Cint(0), Cint(0), ntuple(i -> i <= length(bf) ? Float64(bf[i]) : 0.0, 3),
            eddycode(get(params, :eddyviscosity, nothing))...)
    end
end

const B200Stepper = NamedTuple{(:setup, :psolver, :state, :t, :n),<:Tuple{Any,<:B200PSolver,Any,Any,Any}}
const BuiltinForce = Union{typeof(navierstokes!),typeof(boussinesq!)}

# Function-valued DirichletBC (src/boundary_conditions.jl:344-357): the closure cannot cross the C
# ABI, so it is evaluated here on the boundary-plane coordinates -- exactly the broadcast of the
# reference, `bcfunc.(α, xI..., t)` over the plane `boundary(β, N, Iu[α], isright)` -- into device
# planes the plan keeps pointers to.  `dudt = true` uses the same centred difference (:348-354).
# A time-independent function is evaluated once; a time-dependent one must be refreshed at every
# stage time, so `timestep!` then uses the per-stage composition (INTEGRATION.md §3).
function refresh_bc!(ps::B200PSolver, t; dudt = false)
    setup = ps.setup
    hasfuncbc(setup) || return nothing
    want = ((ps.bc_time_dependent || dudt) ? Float64(t) : 0.0, dudt)
    ps.bcstate == want && return nothing
    t = want[1]
    (; boundary_conditions, xu, Iu, N) = setup
    D = length(N)
    T = eltype(setup.Δ[1])
    planes = ps.bcplanes
    for β = 1:D, (side, isright) in ((1, false), (2, true))
        bc = boundary_conditions.u[β][side]
        (bc isa DirichletBC && bc.u isa Function) || continue
        for α = 1:D
            i = isright ? last(Iu[α].indices[β]) + 1 : first(Iu[α].indices[β]) - 1
            xI = ntuple(g -> g == β ? Array(xu[α][g])[i:i] : Array(xu[α][g]), D)
            xb = ntuple(g -> reshape(xI[g], ntuple(k -> k == g ? length(xI[g]) : 1, D)), D)
            vals = if dudt
                h = sqrt(eps(T)) / 2
                (bc.u.(α, xb..., T(t) + h) .- bc.u.(α, xb..., T(t) - h)) ./ 2h
            else
                bc.u.(α, xb..., T(t))
            end
            host = T.(dropdims(collect(vals); dims = β))      # (other dims), first one fastest
            key = (β, side, α)
            ps.bchost[key] = host
            if haskey(planes, key)
                copyto!(planes[key], host)
            else
                planes[key] = CuArray(host)
                check(ccall((:insb200_set_bc_u_plane, libinsb200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}),
                    ps.plan, β - 1, side - 1, α - 1, devptr(planes[key])))
            end
        end
    end
    ps.bcstate = want
    nothing
end

# ghost fills through the library (the planes are refreshed first): apply_bc_u!(u, t, setup; dudt)
function b200_apply_bc_u!(ps::B200PSolver, u::CuArray, t; dudt = false)
    refresh_bc!(ps, t; dudt)
    check(ccall((:insb200_apply_bc_u, libinsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid}),
        ps.plan, devptr(u), Float64(t), Cint(dudt), stream()))
    u
end

# ---- stage-by-stage composition (user-defined force!, time-dependent Dirichlet functions) ----------
# Mirrors step_explicit_runge_kutta.jl:18-54 / step_lmwray3.jl:41-90 with one library call per stage for
# copyto! + the axpys + apply_bc_u! + project! + apply_bc_u! (+ apply_bc_temp!): insb200_stage_combine_project.
# The force is the caller's: a built-in one goes to insb200_rhs, anything else runs as Julia code on CuArrays.
function stage_cache(ps::B200PSolver, state, nstage)
    c = ps.cache
    if !haskey(c, :k) || length(c[:k]) < nstage || keys(c[:start]) != keys(state)
        c[:start] = map(similar, state)
        c[:k] = [map(similar, state) for _ = 1:nstage]
        c[:p] = similar(state.u, size(state.u)[1:end-1])
    end
    c[:start], c[:k], c[:p]
end

function b200_force!(force!, k, state, t, ps::B200PSolver, params, force_cache)
    if force! isa BuiltinForce
        cpar = Ref(cparams(force!, params))
        temp = haskey(state, :temp) ? state.temp : nothing
        ktemp = haskey(k, :temp) ? k.temp : nothing
        check(ccall((:insb200_rhs, libinsb200), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{CParams}, Ptr{Cvoid}),
            ps.plan, devptr(state.u), devptr(temp), devptr(k.u), devptr(ktemp), cpar, stream()))
    else
        force!(k, state, t; setup = ps.setup, cache = force_cache, params...)
    end
    k
end

function combine_project!(ps::B200PSolver, state, start, ks, nk, coefs::Vector{Float64}, coefT, p, t)
    refresh_bc!(ps, t)
    ku = Ptr{Cvoid}[devptr(ks[j].u) for j = 1:nk]
    hastemp = haskey(state, :temp)
    kt = hastemp ? Ptr{Cvoid}[devptr(ks[j].temp) for j = 1:nk] : Ptr{Cvoid}[]
    GC.@preserve ku kt coefs begin
        check(ccall((:insb200_stage_combine_project, libinsb200), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}},
             Ptr{Cdouble}, Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
            ps.plan, devptr(state.u), devptr(start.u), hastemp ? devptr(state.temp) : C_NULL,
            hastemp ? devptr(start.temp) : C_NULL, nk, ku, hastemp ? pointer(kt) : Ptr{Ptr{Cvoid}}(C_NULL), coefs, coefT,
            devptr(p), Float64(t), stream()))
    end
    state
end

function timestep_staged!(method::ExplicitRungeKuttaMethod{T}, force!, stepper, Δt; params, force_cache) where {T}
    (; setup, psolver, state, t, n) = stepper
    (; A, b, c) = method
    nstage = length(b)
    start, ks, p = stage_cache(psolver, state, nstage)
    map(copyto!, start, state)
    CT = promote_type(T, typeof(Δt))
    coefT = CT == Float64 ? Cint(1) : Cint(0)
    tstart = t
    for i = 1:nstage
        b200_force!(force!, ks[i], state, t, psolver, params, force_cache)
        t = tstart + c[i] * Δt
        combine_project!(psolver, state, start, ks, i, Float64[Float64(CT(Δt) * CT(A[i, j])) for j = 1:i], coefT, p, t)
    end
    create_stepper(method; setup, psolver, state, t, n = n + 1)
end

function timestep_staged!(method::LMWray3, force!, stepper, Δt; params, force_cache)
    (; setup, psolver, state, t, n) = stepper
    T = eltype(state.u)
    start, ks, p = stage_cache(psolver, state, 1)
    map(copyto!, start, state)
    a = T(8 / 15), T(5 / 12), T(3 / 4)
    b = T(1 / 4), T(0)
    c = T(8 / 15), T(2 / 3), T(1)
    coefT = T == Float64 ? Cint(1) : Cint(0)
    tstart = t
    for i = 1:3
        b200_force!(force!, ks[1], state, t, psolver, params, force_cache)
        t = tstart + c[i] * Δt
        combine_project!(psolver, state, start, ks, 1, Float64[Float64(T(a[i] * T(Δt)))], coefT, p, t)
        if i != 3
            for key in keys(state)
                ncomp = key == :u ? length(setup.N) : 1
                check(ccall((:insb200_axpy, libinsb200), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}),
                    psolver.plan, devptr(start[key]), Float64(T(b[i] * T(Δt))), devptr(ks[1][key]), coefT, Cint(ncomp),
                    stream()))
            end
        end
    end
    create_stepper(method; setup, psolver, state, t = tstart + Δt, n = n + 1)
end

# Any force! that is not one of the built-in ones: stage-by-stage composition (no CPU path, less fusion).
# (one method per stepper type: a Union in the first argument would be ambiguous with the reference's own methods)
INS.timestep!(method::ExplicitRungeKuttaMethod, force!, stepper::B200Stepper, Δt;
        params = nothing, ode_cache = nothing, force_cache = nothing) =
    timestep_staged!(method, force!, stepper, Δt; params, force_cache)
INS.timestep!(method::LMWray3, force!, stepper::B200Stepper, Δt;
        params = nothing, ode_cache = nothing, force_cache = nothing) =
    timestep_staged!(method, force!, stepper, Δt; params, force_cache)

# timestep!(::ExplicitRungeKuttaMethod, ...)   (src/time_steppers/step_explicit_runge_kutta.jl:4-57)
function INS.timestep!(method::ExplicitRungeKuttaMethod{T}, force!::BuiltinForce, stepper::B200Stepper, Δt;
        params = nothing, ode_cache = nothing, force_cache = nothing) where {T}
    (; setup, psolver, state, t, n) = stepper
    psolver.bc_time_dependent && return timestep_staged!(method, force!, stepper, Δt; params, force_cache)
    refresh_bc!(psolver, t)
    (; A, b, c) = method
    s = length(b)
    Arow = Float64.(permutedims(A))[:]      # row-major s*s (already the shifted tableau, methods.jl:80-85)
    cpar = Ref(cparams(force!, params))
    coefT = promote_type(T, typeof(Δt)) == Float64 ? Cint(1) : Cint(0)
    temp = haskey(state, :temp) ? state.temp : nothing
    check(ccall((:insb200_step_rk, libinsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ref{CParams}, Ptr{Cvoid}),
        psolver.plan, Arow, Float64.(c), s, coefT, devptr(state.u), devptr(temp), t, Δt, cpar, stream()))
    create_stepper(method; setup, psolver, state, t = t + Δt, n = n + 1)
end

# timestep!(::LMWray3, ...)   (src/time_steppers/step_lmwray3.jl:3-94)
function INS.timestep!(method::LMWray3, force!::BuiltinForce, stepper::B200Stepper, Δt;
        params = nothing, ode_cache = nothing, force_cache = nothing)
    (; setup, psolver, state, t, n) = stepper
    psolver.bc_time_dependent && return timestep_staged!(method, force!, stepper, Δt; params, force_cache)
    refresh_bc!(psolver, t)
    cpar = Ref(cparams(force!, params))
    temp = haskey(state, :temp) ? state.temp : nothing
    check(ccall((:insb200_step_lmwray3, libinsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ref{CParams}, Ptr{Cvoid}),
        psolver.plan, devptr(state.u), devptr(temp), t, Δt, cpar, stream()))
    create_stepper(method; setup, psolver, state, t = t + Δt, n = n + 1)
end

# propose_timestep(force!, state, setup, params)   (src/solver.jl:157-214); needs the plan, so
# it is reached through the psolver: `propose_timestep(force!, state, setup, params, psolver)`.
function INS.propose_timestep(force!::BuiltinForce, state, setup, params, psolver::B200PSolver)
    out = Ref{Cdouble}(0)
    cpar = Ref(cparams(force!, params))
    check(ccall((:insb200_cfl, libinsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CParams}, Ref{Cdouble}, Ptr{Cvoid}),
        psolver.plan, devptr(state.u), cpar, out, stream()))
    eltype(state.u)(out[])
end

# The 4-argument form is what the stock solve_unsteady calls (src/solver.jl:111).  It is specialised on CuArray
# states only; the plan comes from the setup's registered psolver, and without one the generic method runs.
function INS.propose_timestep(force!::BuiltinForce, state::NamedTuple{K,<:Tuple{CuArray,Vararg}}, setup, params) where {K}
    ps = psolver_of(setup)
    isnothing(ps) && return invoke(INS.propose_timestep, Tuple{Any,Any,Any,Any}, force!, state, setup, params)
    INS.propose_timestep(force!, state, setup, params, ps)
end

# solve_unsteady(; ...)   (src/solver.jl:77-149): with a B200 psolver and a built-in force the
# whole time loop runs inside the library (insb200_solve_unsteady): CFL adaptation, clipping to
# tend, the bad-Δt stop, a CUDA-graph replay of the step for fixed Δt; the processors are
# notified through a C call-back every gcd(nupdate...) steps.
function b200_on_step(user::Ptr{Cvoid}, n::Clonglong, t::Cdouble)::Cint
    f = unsafe_pointer_to_objref(user)::Base.RefValue{Any}
    f[](n, t)
    Cint(0)
end

function INS.solve_unsteady_b200(; setup, tlims, start, force!::BuiltinForce = navierstokes!, docopy = true,
        method = LMWray3(; T = eltype(start.u)), psolver::B200PSolver, Δt = nothing, Δt_min = nothing,
        cfl = eltype(start.u)(0.9), n_adapt_Δt = 1, processors = (;), nupdate = 1, params, use_graph = true)
    tstart, tend = tlims
    T = eltype(start.u)
    if psolver.bc_time_dependent
        # planes change at every stage time: the library loop (and its CUDA graph) would freeze them.  The stock
        # driver runs instead; its timestep! / propose_timestep calls dispatch to this extension.
        return INS.solve_unsteady(; setup, tlims, start, force!, docopy, method, psolver, Δt, Δt_min, cfl, n_adapt_Δt,
            processors, params)
    end
    refresh_bc!(psolver, tstart)
    state = docopy ? deepcopy(start) : start
    stepper = Ref{Any}(create_stepper(method; setup, psolver, state, t = tstart))
    obs = INS.Observable(INS.get_state(stepper[]))
    initialized = (; (k => v.initialize(obs) for (k, v) in pairs(processors))...)
    notify = Ref{Any}((n, t) -> begin
        stepper[] = create_stepper(method; setup, psolver, state, t = typeof(tstart)(t), n = Int(n))
        obs[] = INS.get_state(stepper[])
    end)
    isrk = method isa ExplicitRungeKuttaMethod
    s = isrk ? length(method.b) : 0
    Arow = isrk ? Float64.(permutedims(method.A))[:] : Float64[]
    crow = isrk ? Float64.(method.c) : Float64[]
    Tm = isrk ? typeof(method).parameters[1] : T
    coefT = promote_type(Tm, isnothing(Δt) ? T : typeof(Δt)) == Float64 ? Cint(1) : Cint(0)
    timeT = (isnothing(Δt) ? T : typeof(Δt)) == Float32 ? Cint(0) : Cint(1)
    cpar = Ref(cparams(force!, params))
    tout, nout = Ref{Cdouble}(0), Ref{Clonglong}(0)
    temp = haskey(state, :temp) ? state.temp : nothing
    cb = @cfunction(b200_on_step, Cint, (Ptr{Cvoid}, Clonglong, Cdouble))
    GC.@preserve notify Arow crow begin
        status = ccall((:insb200_solve_unsteady, libinsb200), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble,
             Cdouble, Cdouble, Cint, Cint, Ref{CParams}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Cdouble}, Ref{Clonglong},
             Ptr{Cvoid}),
            psolver.plan, isrk ? Cint(0) : Cint(1), Arow, crow, s, coefT, devptr(state.u), devptr(temp), tstart, tend,
            isnothing(Δt) ? -1.0 : Float64(Δt), isnothing(Δt_min) ? -1.0 : Float64(Δt_min), Float64(cfl), n_adapt_Δt,
            timeT, cpar, isempty(processors) ? Cint(0) : Cint(nupdate), cb, pointer_from_objref(notify), Cint(use_graph),
            tout, nout, stream())
    end
    status == 2 ? (@warn lasterror()) : check(status)   # INSB200_STOPPED_BAD_DT: warn and stop (solver.jl:113-116)
    notify[](nout[], tout[])
    outputs = (; (k => processors[k].finalize(initialized[k], obs) for k in keys(processors))...)
    (; state..., t = typeof(tstart)(tout[])), outputs
end

# random_field(setup, t; profile, totalenergy, rng)   (src/initializers.jl:109-183) on the device: the white noise is
# the caller's (`rng`: `randn!` into view(u, Ip, α), the reference's stream) or generated by the library from `seed`;
# the profile closure is evaluated here on the shells 0:kdiag.
function INS.random_field_b200(setup, psolver::B200PSolver, t = zero(eltype(setup.x[1]));
        profile = INS.orlandi_profile, totalenergy = one(eltype(setup.x[1])), rng = nothing, seed = rand(UInt64), kwargs...)
    (; Np, Ip) = setup
    T = eltype(setup.x[1])
    D = length(Np)
    kdiag = isqrt(sum(α -> (Np[α] ÷ 2)^2, 1:D))
    prof = [T(profile(T(κ); kwargs...)) for κ = 0:kdiag]
    e0 = Float64[Float64(T(totalenergy) * p / sum(prof)) for p in prof]
    u = INS.vectorfield(setup)
    fill = isnothing(rng)
    if !fill
        v = similar(setup.x[1], Np)
        for α = 1:D
            INS.Random.randn!(rng, v)
            copyto!(view(u, Ip, α), v)
        end
    end
    check(ccall((:insb200_random_field, libinsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cint, Culonglong, Cdouble, Ptr{Cvoid}),
        psolver.plan, devptr(u), e0, length(e0), Cint(fill), UInt64(seed), Float64(t), stream()))
    u
end

# energyspectrum(u, setup; stuff)   (src/spectral.jl:75-117): the shells 0:kmax of spectral_stuff, evaluated behind the
# psolver_spectral plan's FFT passes; returns (; κ, ehat) like the reference.
function INS.energyspectrum(u::CuArray, setup, psolver::B200PSolver; stuff = nothing)
    T = eltype(setup.x[1])
    kmax = isnothing(stuff) ? minimum(setup.Np) ÷ 2 : Int(last(stuff.κ))
    ehat = Vector{Float64}(undef, kmax + 1)
    check(ccall((:insb200_energy_spectrum, libinsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cvoid}),
        psolver.plan, devptr(u), Cint(kmax), ehat, stream()))
    (; κ = 0:kmax, ehat = T.(ehat))
end

# turbulence_statistics(u, setup, viscosity)   (src/spectral.jl:144-257): the three spectral sums come from the library
# (spectral_energy_field!, squared_wavenumbers_dimensional and spectralsum fused behind the FFT passes), the derived
# scales are formed here exactly like the reference's two methods.
function INS.turbulence_statistics(u::CuArray, setup, viscosity, psolver::B200PSolver)
    (; Np, xlims) = setup
    T = eltype(setup.x[1])
    D = length(Np)
    m = Vector{Float64}(undef, 3)
    check(ccall((:insb200_spectral_moments, libinsb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cvoid}),
        psolver.plan, devptr(u), m, stream()))
    viscosity = T(viscosity)
    e = T(m[1])
    diss = 2 * viscosity * T(m[2])
    kmax = T(π) * minimum(α -> Np[α] / (xlims[α][2] - xlims[α][1]), 1:D)
    if D == 3
        uavg = sqrt(2 * e / 3)
        l_kol = (viscosity^3 / diss)^T(1 / 4)
        l_tay = sqrt(15 * viscosity / diss) * uavg
        l_int = uavg^3 / diss
        t_int = l_int / uavg
        t_tay = l_tay / uavg
        t_kol = sqrt(viscosity / diss)
        Re_int = l_int * uavg / viscosity
        Re_tay = l_tay * uavg / viscosity
        kmax_eta = kmax * l_kol
        return (; e, uavg, diss, l_int, l_tay, l_kol, t_int, t_tay, t_kol, Re_int, Re_tay, kmax_eta)
    end
    uavg = sqrt(e)
    enstrophy = diss / (2 * viscosity)
    omega_rms = sqrt(2 * enstrophy)
    enstrophy_diss = 2 * viscosity * T(m[3])
    palinstrophy = enstrophy_diss / (2 * viscosity)
    l_kra = (viscosity^3 / enstrophy_diss)^T(1 / 6)
    l_omega = uavg / omega_rms
    t_eddy = 1 / omega_rms
    t_ens = enstrophy_diss^(-T(1 / 3))
    Re = uavg * l_omega / viscosity
    kmax_etaK = kmax * l_kra
    (; e, uavg, diss, enstrophy, palinstrophy, enstrophy_diss, omega_rms, l_kra, l_omega, t_eddy, t_ens, Re, kmax_etaK)
end

# right_hand_side!(dudt, u, params_ref, t)   (src/sciml.jl:35-48)
function INS.right_hand_side!(dudt::CuArray, u::CuArray, params_ref::Base.RefValue{<:Tuple{Any,<:B200PSolver,Any}}, t)
    setup, psolver, viscosity = params_ref[]
    refresh_bc!(psolver, t; dudt = true)      # apply_bc_u!(dudt, t, setup; dudt = true), sciml.jl:44
    cpar = Ref(CParams(viscosity, 0.0, 0.0, Cint(0), Cint(0), Cint(0), (0.0, 0.0, 0.0), Cint(0), 0.0))
    check(ccall((:insb200_right_hand_side, libinsb200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ref{CParams}, Ptr{Cvoid}),
        psolver.plan, devptr(dudt), devptr(u), t, cpar, stream()))
    refresh_bc!(psolver, t)                   # back to the values for the next ghost fill of u
    nothing
end

end
