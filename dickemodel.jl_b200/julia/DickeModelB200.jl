# DickeModelB200.jl -- Julia shim over libdicketwa.so (include/dicketwa.h).
#
# SOURCE ONLY: Julia is not installed in the build image or on the GPU box, so this file has never
# been executed; the Python ctypes mirror (dickemodel.jl_b200/*.py) is the harness the tests and the
# benchmark drive, and it follows this file function for function.  A maintainer of
# saulpila/DickeModel.jl would drop this module next to src/TWA.jl and re-export its names; see
# INTEGRATION.md.
#
# It keeps the reference's names and argument meaning:
#   ClassicalDicke.ClassicalDickeSystem(ω₀=,ω=,γ=)   src/ClassicalDicke.jl:48-52
#   ClassicalLMG.ClassicalLMGSystem(Ω=,ξ=)           src/ClassicalLMG.jl:32-36
#   ClassicalSystems.integrate(system,u₀,t;...)      src/ClassicalSystems.jl:78-140
#   TWA.average / variance / survival_probability / calculate_distribution / monte_carlo_integrate /
#   mcs_for_* / mcs_chain / coherent_Wigner_*        src/TWA.jl
# and replaces the per-trajectory Julia closures by POD structs handed to one fused CUDA kernel.
module DickeModelB200

using SymEngine

const libdicketwa = get(ENV, "LIBDICKETWA", joinpath(@__DIR__, "..", "csrc", "libdicketwa.so"))

# ---- POD mirrors of include/dicketwa.h ---------------------------------------------------------
struct CSystem
    kind::Int32
    reserved::Int32
    params::NTuple{4,Float64}
end
struct CSolver
    alg::Int32
    backwards::Int32
    dt::Float64
    tol::Float64
    dtmin::Float64
    maxiters::Int64
    controller::Int32
    chart::Int32
    beta1::Float64
    beta2::Float64
    qmin::Float64
    qmax::Float64
    gamma::Float64
    qoldinit::Float64
    output::Int32
    reserved::Int32
end
struct CSampler
    kind::Int32
    reserved::Int32
    center::NTuple{4,Float64}
    hbar::Float64
    seed::UInt64
    offset::UInt64
    u0::Ptr{Float64}
end
struct CRange
    start::Float64
    stop::Float64
    len::Int64
end
struct CReducer
    kind::Int32
    nexpr::Int32
    exprs::Ptr{Cstring}
    x_axis::Int32
    y_axis::Int32
    x_expr::Cstring
    y_expr::Cstring
    xs::CRange
    ys::CRange
    animate::Int32
    reserved::Int32
end
struct CReducerOut
    sums::Ptr{Float64}
    sums_len::Int64
    counts::Ptr{UInt64}
    counts_len::Int64
end
mutable struct CResult
    red::Ptr{CReducerOut}
    n_valid::Ptr{UInt64}
    n_failed::UInt64
    n_unfinished::UInt64
    steps_accepted::UInt64
    steps_rejected::UInt64
    lane_steps::UInt64
    resample_rounds::Int32
    launches::Int32
    jit::Int32
    reserved::Int32
    kernel_ms::Float64
    total_ms::Float64
    CResult() = new(C_NULL, C_NULL, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.0, 0.0)
end

const SYS_DICKE, SYS_LMG = Int32(0), Int32(1)
const ALG_RK4, ALG_RK8, ALG_DP5, ALG_DOP853, ALG_TSIT5 = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const CTRL_ODE, CTRL_SCIPY = Int32(0), Int32(1)
const CHART_QP, CHART_BLOCH = Int32(0), Int32(1)
const OUTPUT_STEP, OUTPUT_HERMITE = Int32(0), Int32(1)
const SAMPLER_HOST_ARRAY, SAMPLER_HW, SAMPLER_SU2, SAMPLER_HWXSU2 = Int32(0), Int32(1), Int32(2), Int32(3)
const RED_AVERAGE, RED_HIST, RED_SURVIVAL = Int32(0), Int32(1), Int32(2)
const AXIS_EXPR, AXIS_TIME = Int32(0), Int32(1)

check(rc) = rc == 0 ? nothing :
    error(unsafe_string(ccall((:dtwa_last_error, libdicketwa), Cstring, ())))

# ---- context -----------------------------------------------------------------------------------
mutable struct Context
    handle::Ptr{Cvoid}
    function Context(devices::Vector{Int32}=Int32[0])
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:dtwa_create, libdicketwa), Cint, (Ref{Ptr{Cvoid}}, Ptr{Int32}, Cint, Ptr{Cvoid}),
                    h, devices, length(devices), C_NULL))
        ctx = new(h[])
        finalizer(c -> ccall((:dtwa_destroy, libdicketwa), Cint, (Ptr{Cvoid},), c.handle), ctx)
        return ctx
    end
end
const default_context = Ref{Union{Nothing,Context}}(nothing)
# background specialisation threads must be idle before the process unloads libnvrtc
atexit(() -> ccall((:dtwa_jit_quiesce, libdicketwa), Cint, ()))

function context()
    # all visible GPUs of the box by default: the ensemble is sharded over them inside the library
    if default_context[] === nothing
        n = parse(Int, get(ENV, "DTWA_NDEV", "1"))
        default_context[] = Context(Int32.(0:n-1))
    end
    return default_context[]
end

# ---- systems (src/ClassicalSystems.jl:19-45) ---------------------------------------------------
abstract type ClassicalSystem end
struct ClassicalDickeSystem <: ClassicalSystem
    parameters::Vector{Float64}
    ClassicalDickeSystem(; ω₀::Real, ω::Real, γ::Real) = new(Float64.([ω₀, ω, γ]))
end
struct ClassicalLMGSystem <: ClassicalSystem
    parameters::Vector{Float64}
    ClassicalLMGSystem(; Ω::Real, ξ::Real) = new(Float64.([Ω, ξ]))
end
parameters(s::ClassicalSystem) = s.parameters
varnames(::ClassicalDickeSystem) = [:Q, :q, :P, :p]
varnames(::ClassicalLMGSystem) = [:Q, :P]
kind(::ClassicalDickeSystem) = SYS_DICKE
kind(::ClassicalLMGSystem) = SYS_LMG
out_of_bounds(::ClassicalDickeSystem) = (u, _, t) -> 4 - u[3]^2 - u[1]^2 <= 0
out_of_bounds(::ClassicalLMGSystem) = (u, _, t) -> 4 - u[2]^2 - u[1]^2 <= 0
function csystem(s::ClassicalSystem)
    p = parameters(s)
    CSystem(kind(s), 0, ntuple(i -> i <= length(p) ? p[i] : 0.0, 4))
end

# ---- solver keyword arguments (src/ClassicalSystems.jl:78-88,138) ------------------------------
# integrator_alg may be an OrdinaryDiffEq algorithm object or a Symbol; only the name is used.
# The step controller is OrdinaryDiffEq's (DTWA_CTRL_ODE) with the gains of the NAMED algorithm (alg_utils.jl): TsitPap8 /
# Vern8 (generic order 8: beta1 = 7/80, beta2 = 1/20, qmin 1/5, qmax 10), Tsit5 (7/50, 2/25); DP5 / DP8 use the
# library's built-in defaults (zeros here).  solve()'s own keywords beta1/beta2/qmin/qmax/gamma/qoldinit are honoured;
# controller=:scipy selects SciPy's controller (the one the golden step counts pin).
const warned_substitution = Set{Symbol}()
function csolver(; tol::Real=1e-12, integrator_alg=:TsitPap8, integate_backwards::Bool=false, dt=nothing,
                 adaptive=nothing, dtmin=nothing, maxiters=nothing, controller=:ode, beta1=nothing, beta2=nothing,
                 qmin=nothing, qmax=nothing, gamma=nothing, qoldinit=nothing, chart=:qp, output=:step, ignored...)
    name = Symbol(replace(string(integrator_alg isa Symbol ? integrator_alg : nameof(typeof(integrator_alg))), "()" => ""))
    alg = name in (:DP8, :TsitPap8, :Vern8, :DOP853) ? ALG_DOP853 :
          name == :DP5 ? ALG_DP5 : name == :Tsit5 ? ALG_TSIT5 :
          name == :RK4 ? ALG_RK4 : name == :RK8 ? ALG_RK8 :
          error("integrator_alg=$name is not available in libdicketwa")
    if name in (:TsitPap8, :Vern8) && !(name in warned_substitution)
        push!(warned_substitution, name)
        @warn "integrator_alg=$name(): its tableau is not part of libdicketwa; the same-order pair DOP853 runs instead, under OrdinaryDiffEq's step controller for $name"
    end
    if adaptive === false && alg == ALG_DOP853
        alg = ALG_RK8
    end
    generic = name in (:TsitPap8, :Vern8) ? (7 / 80, 1 / 20, 0.2, 10.0) : name == :Tsit5 ? (7 / 50, 2 / 25, 0.2, 10.0) : (0.0, 0.0, 0.0, 0.0)
    b1, b2 = generic[1], generic[2]
    if beta1 !== nothing || beta2 !== nothing
        d2 = name in (:TsitPap8, :Vern8, :Tsit5) ? generic[2] : alg == ALG_DP5 ? 0.04 : 0.0
        b2 = beta2 === nothing ? d2 : Float64(beta2)
        b1 = beta1 !== nothing ? Float64(beta1) : name in (:TsitPap8, :Vern8, :Tsit5) ? generic[1] : alg == ALG_DP5 ? 0.2 - 3b2 / 4 : 0.125 - b2 / 5
    end
    scipy = controller in (:scipy, :SciPy, "scipy")
    CSolver(alg, integate_backwards ? 1 : 0, dt === nothing ? 0.0 : Float64(dt), Float64(tol),
            dtmin === nothing ? 0.0 : Float64(dtmin), maxiters === nothing ? 0 : Int64(maxiters),
            scipy ? CTRL_SCIPY : CTRL_ODE, chart in (:bloch, "bloch") ? CHART_BLOCH : CHART_QP, scipy ? 0.0 : b1, scipy ? 0.0 : b2,
            qmin === nothing ? generic[3] : Float64(qmin), qmax === nothing ? generic[4] : Float64(qmax),
            gamma === nothing ? 0.0 : Float64(gamma), qoldinit === nothing ? 0.0 : Float64(qoldinit),
            output in (:hermite, :interpolate, "hermite") ? OUTPUT_HERMITE : OUTPUT_STEP, 0)
end

# ---- integrate (src/ClassicalSystems.jl:78-140) -------------------------------------------------
struct Solution
    t::Vector{Float64}
    u::Vector{Vector{Float64}}
    retcode::Symbol
end
function integrate(system::ClassicalSystem, u₀::AbstractVector{<:Real}, t::Real; t₀::Real=0.0, saveat=nothing, kargs...)
    t < t₀ && error("Use integate_backwards=true instead of negative time")
    nv = length(varnames(system))
    length(u₀) != nv && error("u₀ should have length $nv, which is the dimension of the phase space of system")
    out_of_bounds(system)(u₀, nothing, 0) && error("The initial condition u₀=$u₀ is out of bounds")
    ts = saveat === nothing ? [Float64(t₀), Float64(t)] : collect(Float64, saveat)
    rel = ts .- t₀
    out = Array{Float64}(undef, nv, length(ts))      # column-major [nvar][nt] == C [nt][nvar]
    status = Ref{Int32}(0)
    nsteps = Ref{Int64}(0)
    u0 = Float64.(u₀)
    sys = Ref(csystem(system)); sol = Ref(csolver(; kargs...))
    GC.@preserve u0 rel out begin
        check(ccall((:dtwa_integrate, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Ref{CSolver}, Ptr{Float64}, Int64, Ptr{Float64}, Int32, Ptr{Float64}, Ref{Int32}, Ref{Int64}),
                    context().handle, sys, sol, u0, 1, rel, length(ts), out, status, nsteps))
    end
    Solution(ts, [out[:, k] for k in 1:length(ts)], (:Success, :DomainError, :MaxIters, :OutOfBounds)[status[]+1])
end

# ---- hamiltonians (src/ClassicalDicke.jl:80-88, src/ClassicalLMG.jl:63-70): plain Julia closures,
#      exactly the reference's, so that they can also be traced into expressions -------------------
function hamiltonian(system::ClassicalDickeSystem)
    ω₀, ω, γ = parameters(system)
    function H(u)
        Q, q, P, p = u
        r = 1 - (Q^2 + P^2) / 4
        (ω₀ / 2) * (Q^2 + P^2) + (ω / 2) * (q^2 + p^2) + 2 * γ * q * Q * sqrt(r) - ω₀
    end
end
function hamiltonian(system::ClassicalLMGSystem)
    Ω, ξ = parameters(system)
    H(u) = (Q = u[1]; P = u[2]; Ω * (Q^2 + P^2) / 2 + ξ * (Q^2 - P^2 * Q^2 / 4 - Q^4 / 4) - Ω)
end

# ---- distributions (src/TWA.jl:22-26,689-787) ---------------------------------------------------
mutable struct PhaseSpaceDistribution
    kind::Int32
    center::NTuple{4,Float64}
    ħ::Float64
    seed::UInt64
    offset::UInt64
    u0::Union{Nothing,Matrix{Float64}}       # [nvar][N] column-major == C [N][nvar]
end
const seed_counter = Ref{UInt64}(20261019)
next_seed(seed) = seed === nothing ? (seed_counter[] += 1; seed_counter[] - 1) : UInt64(seed)
coherent_Wigner_HW(; q₀::Real, p₀::Real, j::Real=1, ħ::Real=1 / j, seed=nothing) =
    PhaseSpaceDistribution(SAMPLER_HW, (q₀, p₀, 0.0, 0.0), ħ, next_seed(seed), 0, nothing)
coherent_Wigner_SU2(; Q₀::Real, P₀::Real, j::Real=1, ħ::Real=1 / j, seed=nothing) =
    PhaseSpaceDistribution(SAMPLER_SU2, (Q₀, P₀, 0.0, 0.0), ħ, next_seed(seed), 0, nothing)
coherent_Wigner_HWxSU2(; Q₀::Real, q₀::Real, P₀::Real, p₀::Real, j::Real=1, ħ::Real=1 / j, seed=nothing) =
    PhaseSpaceDistribution(SAMPLER_HWXSU2, (Q₀, q₀, P₀, p₀), ħ, next_seed(seed), 0, nothing)
coherent_Wigner_HWxSU2(u₀::AbstractVector{<:Real}; kw...) = coherent_Wigner_HWxSU2(; Q₀=u₀[1], q₀=u₀[2], P₀=u₀[3], p₀=u₀[4], kw...)
coherent_Wigner_SU2(u₀; kw...) = coherent_Wigner_SU2(; Q₀=u₀[1], P₀=u₀[2], kw...)
coherent_Wigner_HW(u₀; kw...) = coherent_Wigner_HW(; q₀=u₀[1], p₀=u₀[2], kw...)
samples_from_array(u0::Matrix{Float64}; j::Real=1, ħ::Real=1 / j) =
    PhaseSpaceDistribution(SAMPLER_HOST_ARRAY, (0.0, 0.0, 0.0, 0.0), ħ, 0, 0, u0)
function csampler(d::PhaseSpaceDistribution, offset)
    ptr = d.u0 === nothing ? Ptr{Float64}(C_NULL) : pointer(d.u0) + 8 * size(d.u0, 1) * offset
    CSampler(d.kind, 0, d.center, d.ħ, d.seed, UInt64(offset), ptr)
end

# ---- observables (src/TWA.jl:277-299): everything becomes an expression string -------------------
function expression_string(system::ClassicalSystem, observable)
    names = varnames(system)
    if observable isa Function
        vars = [SymEngine.symbols(string(n)) for n in names]
        ex = applicable(observable, vars...) ? observable(vars...) :
             applicable(observable, vars) ? observable(vars) :
             error("You should pass a function that takes a n-vector or n arguments, where n is the dimension of the phase space.")
        return SymEngine.toString(SymEngine.Basic(ex))
    end
    b = SymEngine.Basic(observable)
    Symbol.(SymEngine.free_symbols(b)) ⊆ names ||
        error("The expression $(observable) has unknown variables. Write it in terms only of $(names)")
    return SymEngine.toString(b)
end

# ---- Monte-Carlo systems (src/TWA.jl:37-51): reducer descriptions + host finaliser --------------
struct ReducerSpec
    kind::Int32
    exprs::Vector{String}
    x_axis::Int32
    y_axis::Int32
    x_expr::String
    y_expr::String
    xs::CRange
    ys::CRange
    animate::Bool
end
struct MonteCarloSystem
    f_loop::Vector{ReducerSpec}
    f_final
    N::Int
    ts::Vector{Float64}
    distribution::PhaseSpaceDistribution
end
linrange(r) = CRange(Float64(first(r)), Float64(last(r)), length(r))
const norange = CRange(0.0, 0.0, 1)

function mcs_for_averaging(system; observable, ts::AbstractVector{<:Real}=[0.0], N::Integer, distribution)
    onevar = !isa(observable, Array)
    obs = onevar ? [observable] : observable
    spec = ReducerSpec(RED_AVERAGE, [expression_string(system, o) for o in obs], 0, 0, "", "", norange, norange, false)
    f_final = res -> (r = res ./ N; onevar ? [x[1] for x in r] : r)
    MonteCarloSystem([spec], f_final, N, collect(Float64, ts), distribution)
end
function mcs_for_variance(system; observable, N::Integer, ts::AbstractArray{<:Real}=[0.0], return_average::Bool=false, distribution)
    onevar = !isa(observable, Array)
    obs = onevar ? [observable] : observable
    ex = [expression_string(system, o) for o in obs]
    tam = length(ex)
    spec = ReducerSpec(RED_AVERAGE, [ex; ["(" * e * ")^2" for e in ex]], 0, 0, "", "", norange, norange, false)
    function f_final(res)
        r = res ./ N
        vars = [[re[i+tam] - re[i]^2 for i in 1:tam] for re in r]
        proms = [[re[i] for i in 1:tam] for re in r]
        if onevar
            vars = [v[1] for v in vars]; proms = [p[1] for p in proms]
        end
        return_average ? (vars, proms) : vars
    end
    MonteCarloSystem([spec], f_final, N, collect(Float64, ts), distribution)
end
function mcs_for_survival_probability(system; N::Integer, ts::AbstractArray{<:Real}, distribution)
    spec = ReducerSpec(RED_SURVIVAL, String[], 0, 0, "", "", norange, norange, false)
    L = length(varnames(system)) / 2
    MonteCarloSystem([spec], res -> [r[1] for r in res] .* ((2 * π * distribution.ħ)^L / N), N, collect(Float64, ts), distribution)
end
function mcs_for_distributions(system; x, xs=nothing, y=nothing, ts=nothing, animate::Bool=false, ys=nothing, N::Integer, distribution)
    if ts === nothing && y === nothing
        y = :t; ts = 0:0
    end
    ts === nothing && error("You have to pass :ts")
    xt, yt = x == :t, y == :t
    spec = ReducerSpec(RED_HIST, String[], xt ? AXIS_TIME : AXIS_EXPR, yt ? AXIS_TIME : AXIS_EXPR,
                       xt ? "" : expression_string(system, x), yt ? "" : expression_string(system, y),
                       xt ? norange : linrange(xs), yt ? norange : linrange(ys), animate && !(xt || yt))
    MonteCarloSystem([spec], identity, N, collect(Float64, ts), distribution)     # orientation handled in monte_carlo_integrate
end
function mcs_chain(Fs::AbstractVector{MonteCarloSystem})
    any(F.N != Fs[1].N for F in Fs) && error("All systems must have the same N")
    any(F.ts != Fs[1].ts for F in Fs) && error("All systems must have the same ts")
    any(F.distribution !== Fs[1].distribution for F in Fs) && error("All systems must have the same phase space distribution")
    sizes = [length(F.f_loop) for F in Fs]
    f_final = res -> [Fs[i].f_final(res[sum(sizes[1:i-1])+1]) for i in 1:length(Fs)]
    MonteCarloSystem(vcat([F.f_loop for F in Fs]...), f_final, Fs[1].N, Fs[1].ts, Fs[1].distribution)
end
mcs_chain(mcss::MonteCarloSystem...) = mcs_chain(collect(mcss))

# ---- the driver (src/TWA.jl:100-224) -------------------------------------------------------------
function monte_carlo_integrate(system::ClassicalSystem, mcs::MonteCarloSystem; tolerate_errors::Bool=true,
                               show_progress::Bool=true, maxNBatch::Real=Inf, kargs...)
    nt = length(mcs.ts)
    nred = length(mcs.f_loop)
    keep = Any[]                                  # GC roots for everything the C structs point to
    reds = Vector{CReducer}(undef, nred)
    outs = Vector{CReducerOut}(undef, nred)
    raw = Vector{Any}(undef, nred)
    for (i, s) in enumerate(mcs.f_loop)
        cstrs = [Base.unsafe_convert(Cstring, Base.cconvert(Cstring, e)) for e in s.exprs]
        push!(keep, s.exprs, cstrs)
        reds[i] = CReducer(s.kind, length(s.exprs), isempty(cstrs) ? C_NULL : pointer(cstrs), s.x_axis, s.y_axis,
                           Base.unsafe_convert(Cstring, s.x_expr), Base.unsafe_convert(Cstring, s.y_expr), s.xs, s.ys, s.animate ? 1 : 0, 0)
        if s.kind == RED_HIST
            dims = s.x_axis == AXIS_TIME ? (Int(s.ys.len), nt) : s.y_axis == AXIS_TIME ? (Int(s.xs.len), nt) :
                   s.animate ? (Int(s.xs.len), Int(s.ys.len), nt) : (Int(s.xs.len), Int(s.ys.len))
            raw[i] = zeros(UInt64, dims)          # column-major mirror of the C layouts in dicketwa.h
            outs[i] = CReducerOut(C_NULL, 0, pointer(raw[i]), length(raw[i]))
        else
            n = s.kind == RED_SURVIVAL ? 1 : length(s.exprs)
            raw[i] = zeros(Float64, n, nt)
            outs[i] = CReducerOut(pointer(raw[i]), length(raw[i]), C_NULL, 0)
        end
    end
    res = CResult()
    res.red = pointer(outs)
    d = mcs.distribution
    sys = Ref(csystem(system)); sol = Ref(csolver(; kargs...)); smp = Ref(csampler(d, d.offset))
    GC.@preserve keep reds outs raw d begin
        check(ccall((:dtwa_ensemble, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Ref{CSolver}, Ref{CSampler}, Int64, Ptr{Float64}, Int32, Ptr{CReducer}, Int32, Int32, Ref{CResult}),
                    context().handle, sys, sol, smp, mcs.N, mcs.ts, nt, reds, nred, tolerate_errors ? 1 : 0, res))
    end
    d.offset += mcs.N
    final = Vector{Any}(undef, nred)
    for (i, s) in enumerate(mcs.f_loop)
        if s.kind == RED_HIST
            m = Float64.(raw[i]) ./ mcs.N                                   # src/TWA.jl:518
            final[i] = s.x_axis == AXIS_TIME ? m :                          # [iy, i_t]   (:459)
                       s.y_axis == AXIS_TIME ? permutedims(m) :             # [i_t, ix]   (:466)
                       s.animate ? [permutedims(m[:, :, k]) for k in 1:nt] : permutedims(m)   # [iy, ix] (:474,:487)
        else
            final[i] = [raw[i][:, k] for k in 1:nt]
        end
    end
    return mcs.f_final(nred == 1 ? final[1] : final)
end

average(system; observable, N::Integer, ts::AbstractArray{<:Real}=[0.0], distribution, kargs...) =
    monte_carlo_integrate(system, mcs_for_averaging(system; observable=observable, ts=ts, N=N, distribution=distribution); kargs...)
variance(system; observable, N::Integer, ts::AbstractArray{<:Real}=[0.0], distribution, return_average=false, kargs...) =
    monte_carlo_integrate(system, mcs_for_variance(system; observable=observable, ts=ts, N=N, distribution=distribution, return_average=return_average); kargs...)
survival_probability(system; distribution, N::Integer, ts::AbstractArray{<:Real}, kargs...) =
    monte_carlo_integrate(system, mcs_for_survival_probability(system; ts=ts, N=N, distribution=distribution); integate_backwards=true, kargs...)
calculate_distribution(system; x, N::Integer, animate::Bool=false, xs=nothing, y=nothing, ts=nothing, ys=nothing, distribution, kargs...) =
    monte_carlo_integrate(system, mcs_for_distributions(system; x=x, y=y, ts=ts, xs=xs, ys=ys, N=N, distribution=distribution, animate=animate); kargs...)

# ---- SURVEY 8f ranks 2-4: fundamental matrix, Lyapunov spectra, surfaces of section, energy-shell quadrature ----
# integrate(...; get_fundamental_matrix=true) (src/ClassicalSystems.jl:110-128): returns (x(t), Φ(t)) at saveat times
function integrate_with_fundamental_matrix(system::ClassicalSystem, u₀::AbstractVector{<:Real}, t::Real; t₀::Real=0.0,
                                           saveat=nothing, fundamental_matrix_initial=nothing, kargs...)
    nv = length(varnames(system))
    ts = saveat === nothing ? [Float64(t₀), Float64(t)] : collect(Float64, saveat)
    rel = ts .- t₀
    xs = Array{Float64}(undef, nv, length(ts)); Φs = Array{Float64}(undef, nv, nv, length(ts))   # column-major == the C layout
    fm0 = fundamental_matrix_initial === nothing ? C_NULL : pointer(Matrix{Float64}(fundamental_matrix_initial))
    status = Ref{Int32}(0); nsteps = Ref{Int64}(0); u0 = Float64.(u₀)
    sys = Ref(csystem(system)); sol = Ref(csolver(; kargs...))
    GC.@preserve u0 rel xs Φs fundamental_matrix_initial begin
        check(ccall((:dtwa_integrate_fm, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Ref{CSolver}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ref{Int32}, Ref{Int64}),
                    context().handle, sys, sol, u0, fm0, 1, rel, length(ts), xs, Φs, status, nsteps))
    end
    return [(xs[:, k], Φs[:, :, k]) for k in 1:length(ts)]
end

# lyapunov_spectrum (src/ClassicalSystems.jl:183-224); `x` may be a matrix nvar × n: one launch for a whole map
function lyapunov_spectrum(system::ClassicalSystem, x::AbstractVecOrMat{<:Real}, t::Real=100; λtol::Real=1e-4, λreltol::Real=1e-3,
                           maxiters::Integer=100, verbose::Bool=true, kargs...)
    nv = length(varnames(system)); X = Matrix{Float64}(reshape(x, nv, :)); n = size(X, 2)
    λ = Array{Float64}(undef, nv, n); iters = Vector{Int32}(undef, n); status = Vector{Int32}(undef, n)
    sys = Ref(csystem(system)); sol = Ref(csolver(; kargs...))
    GC.@preserve X λ iters status begin
        check(ccall((:dtwa_lyapunov, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Ref{CSolver}, Ptr{Float64}, Int64, Float64, Float64, Float64, Int32,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
                    context().handle, sys, sol, X, n, t, λtol, λreltol, maxiters, λ, C_NULL, iters, status, C_NULL, C_NULL))
    end
    any(status .== 5) && x isa AbstractVector && error("Maximum iterations")
    return x isa AbstractVector ? λ[:, 1] : λ
end
lyapunov_exponent(system::ClassicalSystem, x::AbstractVector{<:Real}, t::Real=100; kargs...) = maximum(lyapunov_spectrum(system, x, t; kargs...))

# integrate(...; callback=ContinuousCallback((x,t,_)->coef⋅x-offset, affect!; save_positions=(false,false)))
# (docs/src/ClassicalDickeExamples.md:82-109): all crossings found on the GPU, `affect!` called afterwards in time order
struct CEvent; coef::NTuple{4,Float64}; offset::Float64; direction::Int32; pad::Int32; end
function section_crossings(system::ClassicalSystem, u₀::AbstractVector{<:Real}, t::Real; coef::AbstractVector{<:Real}, offset::Real=0.0,
                           direction::Integer=0, max_events::Integer=65536, kargs...)
    nv = length(varnames(system)); cf = zeros(4); cf[1:nv] .= coef
    ev = Ref(CEvent(Tuple(cf), offset, direction, 0))
    et = Vector{Float64}(undef, max_events); eu = Array{Float64}(undef, nv, max_events); ed = Vector{Int32}(undef, max_events)
    cnt = Ref{Int32}(0); uend = Vector{Float64}(undef, nv); status = Ref{Int32}(0); u0 = Float64.(u₀)
    sys = Ref(csystem(system)); sol = Ref(csolver(; kargs...))
    GC.@preserve u0 et eu ed uend begin
        check(ccall((:dtwa_integrate_events, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Ref{CSolver}, Ptr{Float64}, Int64, Float64, Ref{CEvent}, Int32, Ptr{Float64}, Ptr{Float64},
                     Ptr{Int32}, Ref{Int32}, Ptr{Float64}, Ref{Int32}, Ptr{Int64}, Ptr{Float64}),
                    context().handle, sys, sol, u0, 1, t, ev, max_events, et, eu, ed, cnt, uend, status, C_NULL, C_NULL))
    end
    k = min(cnt[], max_events)
    return et[1:k], [eu[:, i] for i in 1:k], ed[1:k], uend
end

# Sections of many initial conditions (columns of U₀) in one launch, packed output: (offsets, t, u, direction, count) where
# initial condition i owns the entries offsets[i]+1 : offsets[i+1] (1-based) of t / u[:, k] / direction
function section_crossings_batch(system::ClassicalSystem, U₀::AbstractMatrix{<:Real}, t::Real; coef::AbstractVector{<:Real},
                                 offset::Real=0.0, direction::Integer=0, max_events::Integer=1024, kargs...)
    nv = length(varnames(system)); cf = zeros(4); cf[1:nv] .= coef
    ev = Ref(CEvent(Tuple(cf), offset, direction, 0))
    U = Matrix{Float64}(U₀); n = size(U, 2)
    offsets = Vector{Int64}(undef, n + 1); cnt = Vector{Int32}(undef, n); status = Vector{Int32}(undef, n)
    uend = Matrix{Float64}(undef, nv, n)
    sys = Ref(csystem(system)); sol = Ref(csolver(; kargs...))
    GC.@preserve U offsets cnt status uend begin
        check(ccall((:dtwa_integrate_events_packed, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Ref{CSolver}, Ptr{Float64}, Int64, Float64, Ref{CEvent}, Int32, Ptr{Int64}, Ptr{Int32},
                     Ptr{Float64}, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
                    context().handle, sys, sol, U, n, t, ev, max_events, offsets, cnt, uend, status, C_NULL, C_NULL))
    end
    total = offsets[end]
    et = Vector{Float64}(undef, total); eu = Matrix{Float64}(undef, nv, total); ed = Vector{Int32}(undef, total)
    GC.@preserve et eu ed begin
        check(ccall((:dtwa_integrate_events_packed_fetch, libdicketwa), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    context().handle, total, et, eu, ed))
    end
    return offsets, et, eu, ed, cnt
end

# EnergyShellProjections.∫∫dqdpδϵ / matrix_QP∫∫dqdpδϵ (src/EnergyShellProjections.jl:49-96,159-199) for classical f
function ∫∫dqdpδϵ(system::ClassicalDickeSystem; ϵ::Real, Q, P, f, p_res::Real=0.01, nonvalue=NaN, onlyqroot=nothing)
    QP = Matrix{Float64}(vcat(reshape(collect(Float64, Q), 1, :), reshape(collect(Float64, P), 1, :))); n = size(QP, 2)
    exprs = f isa AbstractVector ? [expression_string(system, g) for g in f] : [expression_string(system, f)]
    out = Array{Float64}(undef, length(exprs), n); valid = Vector{Int32}(undef, n)
    root = onlyqroot === nothing ? 0 : (onlyqroot === (+) ? 1 : -1)
    sys = Ref(csystem(system))
    cptrs = [pointer(e) for e in exprs]        # NUL-terminated UTF-8 (Julia Strings are), kept alive by GC.@preserve exprs
    GC.@preserve QP exprs cptrs out valid begin
        check(ccall((:dtwa_shell_quadrature, libdicketwa), Cint,
                    (Ptr{Cvoid}, Ref{CSystem}, Float64, Ptr{Float64}, Int64, Ptr{Ptr{UInt8}}, Int32, Float64, Int32, Ptr{Float64},
                     Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                    context().handle, sys, ϵ, QP, n, cptrs, length(exprs), p_res, root, out, valid, C_NULL, C_NULL))
    end
    res = [valid[i] == 1 ? (f isa AbstractVector ? out[:, i] : out[1, i]) : nonvalue for i in 1:n]
    return Q isa Real ? res[1] : res
end

end # module
