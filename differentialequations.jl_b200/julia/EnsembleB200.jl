# EnsembleB200.jl -- the reference-side binding of libb200ens.so (UNEXECUTED here: Julia is not
# installed in the build image; kept in lock-step with include/b200ens.h and mirrored by the
# Python host layer differentialequations.jl_b200/api.py, which IS tested).
#
# It adds ONE ensemble algorithm next to EnsembleThreads / EnsembleSerial / EnsembleDistributed
# (/root/reference/test/qa/qa.jl:49-56) and ONE __solve method; nothing in
# DifferentialEquations.jl itself changes (it only re-exports, src/DifferentialEquations.jl:8-9).
module EnsembleB200Backend

using SciMLBase, Symbolics
import SciMLBase: __solve, AbstractEnsembleProblem, EnsembleAlgorithm, EnsembleSolution, ReturnCode

const LIB = get(ENV, "B200ENS_LIB", "libb200ens.so")

struct EnsembleB200 <: EnsembleAlgorithm
    devices::Vector{Int}          # empty = all visible GPUs
    refill_threshold::Int         # 0 = auto
end
EnsembleB200(; devices = Int[], refill_threshold = 0) = EnsembleB200(devices, refill_threshold)

# ---- C structs (field order == include/b200ens.h)
struct ModelDesc
    struct_size::UInt32; n_state::Int32; n_param::Int32; dtype::Int32; alg::Int32; flags::UInt32
    rhs_src::Cstring; jac_src::Cstring; tgrad_src::Cstring; noise_src::Cstring
    condition_src::Cstring; affect_src::Cstring; name::Cstring
    dcondition_src::Cstring; daffect_src::Cstring
end
mutable struct Opts
    struct_size::UInt32; adaptive::Int32
    t0::Float64; t1::Float64; dt::Float64; abstol::Float64; reltol::Float64; dtmin::Float64; dtmax::Float64
    qmin::Float64; qmax::Float64; gamma::Float64; beta1::Float64; beta2::Float64; qoldinit::Float64
    maxiters::Int64; seed::UInt64; traj_offset::UInt64
    noise_injected::Int32; event_terminate::Int32; interp_points::Int32; save_tstops::Int32
    device_mask::UInt32; refill_threshold::Int32; block_threads::Int32; stage_outputs::Int32
    work_order::Int32; save_everystep::Int32
    abstol_vec::Ptr{Float64}; reltol_vec::Ptr{Float64}
    Opts() = new()
end
struct Stats; naccept::Int32; nreject::Int32; nf::Int32; nevents::Int32; end
mutable struct Timing
    h2d_ms::Float64; kernel_ms::Float64; d2h_ms::Float64; total_ms::Float64
    n_devices::Int32; launches::Int32; grid::Int32; block::Int32; smem_bytes::Int32; regs::Int32
    Timing() = new()
end

const ALG_IDS = Dict(:Tsit5 => 1, :Vern7 => 2, :Rosenbrock23 => 3, :Rodas5 => 4, :Rodas5P => 5, :EM => 6, :SOSRA => 7, :Rodas4 => 8, :SRIW1 => 9)
const RETCODES = (ReturnCode.Default, ReturnCode.Success, ReturnCode.Terminated, ReturnCode.MaxIters,
                  ReturnCode.DtLessThanMin, ReturnCode.Unstable, ReturnCode.DtNaN, ReturnCode.Failure)

check(rc) = rc == 0 || error("libb200ens: " * unsafe_string(ccall((:b200ens_last_error, LIB), Cstring, ())))

# ---- codegen: trace f through Symbolics, emit C, wrap as the device functions the ABI documents
function cuda_source(name, outname, exprs)
    body = join(["    $outname[$(i-1)] = " * string(Symbolics.toexpr(Symbolics.build_function(e; target = Symbolics.CTarget(), expression = Val{true})))  * ";" for (i, e) in enumerate(exprs)], "\n")
    "__device__ __forceinline__ void $name(real* $outname, const real* u, const real* p, const real t) {\n$body\n}\n"
end

function model_sources(prob, alg)
    n, m = length(prob.u0), length(prob.p)
    @variables t u[1:n] p[1:m]
    us, ps = collect(u), collect(p)
    du = SciMLBase.isinplace(prob) ? (d = similar(us, Num); prob.f(d, us, ps, t); d) : prob.f(us, ps, t)
    rhs = cuda_source("b2_rhs", "du", du)
    M = prob.f.mass_matrix                              # ODEFunction(f; mass_matrix = M): constant table in the RHS source
    if !(M isa SciMLBase.LinearAlgebra.UniformScaling)
        rhs = "#undef B2_HAS_MASS\n#define B2_HAS_MASS 1\nstatic constexpr double B2_MASS_[$(n*n)] = {" *
              join(string.(Float64.(vec(permutedims(Matrix(M))))), ", ") * "};\n" * rhs
    end
    jac = nameof(typeof(alg)) in (:Rosenbrock23, :Rodas4, :Rodas5, :Rodas5P) ?
          cuda_source("b2_jac", "J", vec(permutedims(Symbolics.jacobian(du, us)))) : nothing
    noise = prob isa SDEProblem ? cuda_source("b2_noise", "g", prob.g(us, ps, t)) : nothing
    (rhs, jac, noise)
end

function __solve(eprob::AbstractEnsembleProblem, alg, ens::EnsembleB200; trajectories, saveat = nothing,
                 dt, abstol = 1e-6, reltol = 1e-3, adaptive = true, maxiters = 100_000, seed = UInt64(0),
                 callback = nothing, kwargs...)
    prob = eprob.prob
    T = eltype(prob.u0); n = length(prob.u0); m = length(prob.p); N = trajectories
    # prob_func on the host, exactly like EnsembleThreads/EnsembleGPUKernel (SURVEY 3.3/3.4); v3 may pass a context
    U0 = Matrix{T}(undef, n, N); P = Matrix{T}(undef, m, N)
    for i in 1:N
        pi = applicable(eprob.prob_func, prob, i, 1) ? eprob.prob_func(prob, i, 1) : eprob.prob_func(prob, SciMLBase.EnsembleContext(i))
        U0[:, i] .= pi.u0; P[:, i] .= pi.p
    end
    ts = saveat === nothing ? T[prob.tspan...] : saveat isa Number ? collect(T, prob.tspan[1]:saveat:prob.tspan[2]) : collect(T, saveat)
    rhs, jac, noise = model_sources(prob, alg)
    model = Ref{Ptr{Cvoid}}(C_NULL); log = Vector{UInt8}(undef, 1 << 16)
    GC.@preserve rhs jac noise begin
        d = ModelDesc(sizeof(ModelDesc), n, m, T == Float64 ? 1 : 0, ALG_IDS[nameof(typeof(alg))], 0,
                      pointer(rhs), jac === nothing ? C_NULL : pointer(jac), C_NULL,
                      noise === nothing ? C_NULL : pointer(noise), C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)
        check(ccall((:b200ens_compile, LIB), Cint, (Ref{ModelDesc}, Ref{Ptr{Cvoid}}, Ptr{UInt8}, Csize_t), d, model, log, length(log)))
    end
    o = Opts(); ccall((:b200ens_opts_init, LIB), Cvoid, (Ref{Opts},), o)
    o.adaptive = adaptive; o.t0, o.t1 = prob.tspan; o.dt = dt
    # scalar tolerances, or one per state component (solve(prob, Rodas5P(); abstol = [1e-8, 1e-14, 1e-6]))
    atolv = abstol isa AbstractVector ? collect(Float64, abstol) : Float64[]
    rtolv = reltol isa AbstractVector ? collect(Float64, reltol) : Float64[]
    o.abstol = isempty(atolv) ? abstol : atolv[1]; o.reltol = isempty(rtolv) ? reltol : rtolv[1]
    o.abstol_vec = isempty(atolv) ? C_NULL : pointer(atolv); o.reltol_vec = isempty(rtolv) ? C_NULL : pointer(rtolv)
    o.maxiters = maxiters; o.seed = seed; o.refill_threshold = ens.refill_threshold
    o.device_mask = isempty(ens.devices) ? 0 : reduce(|, UInt32(1) .<< ens.devices)
    out = Array{T, 3}(undef, n, length(ts), N)          # column-major == [N][n_save][n_state] of the ABI
    rc = Vector{Int32}(undef, N); st = Vector{Stats}(undef, N); tm = Timing()
    elapsed = GC.@preserve atolv rtolv @elapsed check(ccall((:b200ens_solve, LIB), Cint,
        (Ptr{Cvoid}, Ref{Opts}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Int32, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Int32}, Ptr{Stats}, Ref{Timing}),
        model[], o, N, U0, P, ts, length(ts), C_NULL, out, C_NULL, rc, st, tm))
    ccall((:b200ens_free, LIB), Cvoid, (Ptr{Cvoid},), model[])
    sols = map(1:N) do i
        ui = [out[:, k, i] for k in 1:length(ts)]       # zero-copy reinterpret to SVector in a real package
        SciMLBase.build_solution(prob, alg, ts, ui; retcode = RETCODES[rc[i] + 1])
    end
    EnsembleSolution(sols, elapsed, all(==(1), rc))
end

export EnsembleB200
end # module
