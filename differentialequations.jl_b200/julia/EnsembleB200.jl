# EnsembleB200.jl -- the reference-side binding of libb200ens.so.
#
# STATUS: UNEXECUTED.  Julia is not installed in the build image, so this file has never run.  What IS checked
# mechanically (tests/test_julia_binding.py, CPU suite): every `struct` below has exactly the fields -- order, C type,
# count -- of its counterpart in include/b200ens.h, every `ccall` names an exported symbol with the argument count of
# the C prototype, and ALG_IDS / RETCODES agree with the header's enums.  Everything else is written by inspection
# against SciMLBase 3 / OrdinaryDiffEq 7 as the reference pins them (/root/reference/Project.toml:7,10,13,18) and
# mirrors, statement by statement, the Python host layer differentialequations.jl_b200/api.py, which is tested on
# the GPU.  Lines that depend on upstream internals that could not be verified here are marked `UNVERIFIED`.
#
# It adds ONE ensemble algorithm next to EnsembleThreads / EnsembleSerial / EnsembleDistributed
# (/root/reference/test/qa/qa.jl:49-56) and ONE __solve method; nothing in DifferentialEquations.jl itself changes
# (it only re-exports, src/DifferentialEquations.jl:8-9).  Anything the library cannot do is an `error(...)`:
# there is no CPU fallback and nothing is silently ignored.
module EnsembleB200Backend

using SciMLBase, Symbolics
import SciMLBase: __solve, AbstractEnsembleProblem, EnsembleAlgorithm, EnsembleSolution, ReturnCode

const LIB = get(ENV, "B200ENS_LIB", "libb200ens.so")
const ABI_VERSION = 7

struct EnsembleB200 <: EnsembleAlgorithm
    devices::Vector{Int}          # empty = all visible GPUs
    refill_threshold::Int         # 0 = auto
end
EnsembleB200(; devices = Int[], refill_threshold = 0) = EnsembleB200(devices, refill_threshold)

# ---- C structs (field order == include/b200ens.h; checked by tests/test_julia_binding.py)
struct ModelDesc
    struct_size::UInt32; n_state::Int32; n_param::Int32; dtype::Int32; alg::Int32; flags::UInt32
    rhs_src::Cstring; jac_src::Cstring; tgrad_src::Cstring; noise_src::Cstring
    condition_src::Cstring; affect_src::Cstring; name::Cstring
    dcondition_src::Cstring; daffect_src::Cstring
    save_idxs::Ptr{Int32}; n_save_idxs::Int32; reserved0::Int32
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
    noise_stream_len::Int64
    shard_blocks::Int32; n_tstops::Int32
    tstops::Ptr{Float64}
    Opts() = new()
end
struct Stats; naccept::Int32; nreject::Int32; nf::Int32; nevents::Int32; end
mutable struct Timing
    h2d_ms::Float64; kernel_ms::Float64; d2h_ms::Float64; total_ms::Float64
    n_devices::Int32; launches::Int32; grid::Int32; block::Int32; smem_bytes::Int32; regs::Int32
    kernel_ms_min::Float64
    Timing() = new()
end

const ALG_IDS = Dict(:Tsit5 => 1, :Vern7 => 2, :Rosenbrock23 => 3, :Rodas5 => 4, :Rodas5P => 5, :EM => 6, :SOSRA => 7, :Rodas4 => 8, :SRIW1 => 9, :FBDF => 10)
# b200ens_retcode 0..7 -> SciMLBase.ReturnCode, by NAME (the integer values of upstream's enum are not part of the ABI)
const RETCODES = (ReturnCode.Default, ReturnCode.Success, ReturnCode.Terminated, ReturnCode.MaxIters,
                  ReturnCode.DtLessThanMin, ReturnCode.Unstable, ReturnCode.DtNaN, ReturnCode.Failure)

lasterror() = unsafe_string(ccall((:b200ens_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 || error("libb200ens: " * lasterror())

function __init__()
    v = ccall((:b200ens_abi_version, LIB), Cint, ())
    v == ABI_VERSION || error("libb200ens ABI version $v, this binding was written for $ABI_VERSION")
end

# ---- pinned host arrays (b200ens_host_alloc): a plain Julia Array is pageable and costs 3x on the PCIe path
function pinned(::Type{T}, dims::Int...) where {T}
    ptr = ccall((:b200ens_host_alloc, LIB), Ptr{Cvoid}, (Csize_t,), max(1, prod(dims) * sizeof(T)))
    ptr == C_NULL && error("libb200ens: " * lasterror())
    A = unsafe_wrap(Array, Ptr{T}(ptr), dims; own = false)
    finalizer(_ -> ccall((:b200ens_host_free, LIB), Cvoid, (Ptr{Cvoid},), ptr), A)
    A
end

# ---- codegen: trace the problem's functions through Symbolics and emit the device functions the ABI documents.
# One C statement per output, from Symbolics' C target; `real` is float or double on the device.
cexpr(e) = Symbolics.build_function(e; target = Symbolics.CTarget(), expression = Val{true})   # UNVERIFIED: exact form of the C-target output
function vector_fn(name, outname, exprs)
    body = join(["    $outname[$(i - 1)] = " * string(cexpr(e)) * ";" for (i, e) in enumerate(exprs)], "\n")
    "__device__ __forceinline__ void $name(real* __restrict__ $outname, const real* __restrict__ u, const real* __restrict__ p, const real t) {\n$body\n}\n"
end
scalar_fn(name, rettype, e) = "__device__ __forceinline__ $rettype $name(const real* __restrict__ u, const real* __restrict__ p, const real t) {\n    return " * string(cexpr(e)) * ";\n}\n"

# The object affect!(integrator) / condition(u, t, integrator) see while being traced: symbolic u, p, t.
# Only assignments to integrator.u are representable on the device (the kernel's affect takes p as const);
# writes to p or t, or anything else, raise instead of being dropped.
mutable struct TraceIntegrator
    u::Vector{Num}
    p::Tuple            # immutable on purpose: integrator.p[i] = ... is an error, not a silent no-op
    t::Num
    terminated::Bool
end
SciMLBase.terminate!(i::TraceIntegrator, args...) = (i.terminated = true; nothing)   # UNVERIFIED: upstream's method signature
function Base.setproperty!(i::TraceIntegrator, f::Symbol, v)
    f in (:u, :terminated) || error("EnsembleB200: affect! may only modify integrator.u (tried to set integrator.$f)")
    setfield!(i, f, v)
end

function trace_affect(affect!, us, ps, t, args...)
    integ = TraceIntegrator(copy(us), Tuple(ps), t, false)
    affect!(integ, args...)
    integ.u, integ.terminated
end
function affect_fn(name, newu, extra = "")
    body = join(["    const real v$(i - 1) = " * string(cexpr(e)) * ";" for (i, e) in enumerate(newu)], "\n") * "\n" *
           join(["    u[$(i - 1)] = v$(i - 1);" for i in eachindex(newu)], "\n")
    "__device__ __forceinline__ void $name(real* __restrict__ u, const real* __restrict__ p, const real t$extra) {\n$body\n}\n"
end
# bit mask of the state components a symbolic expression reads (B2_COND_MASK: the event search only interpolates those)
cond_mask(exprs, us) = reduce(|, (UInt32(1) << (i - 1) for (i, ui) in enumerate(us) if any(e -> Symbolics.occursin(ui, e), exprs)); init = UInt32(0))   # UNVERIFIED: Symbolics.occursin argument order

flatten_callbacks(::Nothing) = ()
flatten_callbacks(cb::SciMLBase.CallbackSet) = (cb.continuous_callbacks..., cb.discrete_callbacks...)
flatten_callbacks(cb) = (cb,)

"""(condition_src, affect_src, dcondition_src, daffect_src, event_terminate, interp_points) or an error: every callback
is either lowered to device source or rejected -- never dropped."""
function callback_sources(callback, us, ps, t)
    cbs = flatten_callbacks(callback)
    cont = [c for c in cbs if c isa SciMLBase.ContinuousCallback || c isa SciMLBase.VectorContinuousCallback]
    disc = [c for c in cbs if c isa SciMLBase.DiscreteCallback]
    length(cont) + length(disc) == length(cbs) || error("EnsembleB200: unsupported callback type in $(typeof.(cbs))")
    length(disc) <= 1 || error("EnsembleB200: at most one DiscreteCallback per solve")
    csrc = asrc = dcsrc = dasrc = nothing
    term = 0; ip = 10
    integ0 = TraceIntegrator(copy(us), Tuple(ps), t, false)
    if !isempty(cont)
        # direction of every continuous callback (SURVEY A.8): 0 both, +1 upcrossings only (affect_neg! === nothing),
        # -1 downcrossings only (affect! === nothing)
        cbdir(c) = c.affect_neg! === nothing ? 1 : (c.affect! === nothing ? -1 : 0)
        for c in cont
            (c.affect! === nothing && c.affect_neg! === nothing) && error("EnsembleB200: affect! and affect_neg! are both nothing")
            c.save_positions == (false, false) || c.save_positions == (true, true) ||   # with saveat upstream's GPU path requires (false,false); (true,true) is the CPU default and has no effect with saveat
                error("EnsembleB200: save_positions = $(c.save_positions) is not supported")
        end
        ip = cont[1].interp_points
        all(c -> c.interp_points == ip, cont) || error("EnsembleB200: all continuous callbacks must share interp_points")
        if length(cont) == 1 && cont[1] isa SciMLBase.ContinuousCallback
            c1 = cont[1]; d1 = cbdir(c1)
            g = c1.condition(us, t, integ0)
            csrc = (d1 == 0 ? "" : "#undef B2_EVENT_DIR\n#define B2_EVENT_DIR $d1\n") *
                   "#undef B2_COND_MASK\n#define B2_COND_MASK 0x$(string(cond_mask([g], us), base = 16))u\n" * scalar_fn("b2_condition", "real", g)
            # the one affect that can run is b2_affect; a two-sided callback with its own affect_neg! adds b2_affect_neg
            newu, terminated = trace_affect(d1 < 0 ? c1.affect_neg! : c1.affect!, us, ps, t)
            asrc = affect_fn("b2_affect", newu)
            term |= terminated ? 1 : 0
            if d1 == 0 && c1.affect_neg! !== c1.affect!
                newn, termn = trace_affect(c1.affect_neg!, us, ps, t)
                asrc *= "#undef B2_HAS_AFFECT_NEG\n#define B2_HAS_AFFECT_NEG 1\n" * affect_fn("b2_affect_neg", newn)
                term |= termn ? 4 : 0
            end
        else
            # several ContinuousCallbacks and/or VectorContinuousCallbacks: one vector callback on the device
            gs = Num[]; branches = String[]; tmask = UInt32(0)
            dset = cbdir(cont[1])
            all(c -> cbdir(c) == dset && (dset != 0 || c.affect_neg! === c.affect!), cont) ||
                error("EnsembleB200: several continuous callbacks need one common direction (all two-sided with affect_neg! === affect!, all upcrossing-only or all downcrossing-only)")
            for c in cont
                if c isa SciMLBase.VectorContinuousCallback
                    out = Vector{Num}(undef, c.len); c.condition(out, us, t, integ0)
                    for k in 1:c.len
                        newu, terminated = trace_affect(dset < 0 ? c.affect_neg! : c.affect!, us, ps, t, k)
                        push!(gs, out[k]); push!(branches, join(["u[$(i - 1)] = " * string(cexpr(e)) * ";" for (i, e) in enumerate(newu)], " "))
                        terminated && (tmask |= UInt32(1) << (length(gs) - 1))
                    end
                else
                    newu, terminated = trace_affect(dset < 0 ? c.affect_neg! : c.affect!, us, ps, t)
                    push!(gs, c.condition(us, t, integ0)); push!(branches, join(["u[$(i - 1)] = " * string(cexpr(e)) * ";" for (i, e) in enumerate(newu)], " "))
                    terminated && (tmask |= UInt32(1) << (length(gs) - 1))
                end
            end
            length(gs) <= 16 || error("EnsembleB200: at most 16 event functions")
            csrc = (dset == 0 ? "" : "#undef B2_EVENT_DIR\n#define B2_EVENT_DIR $dset\n") *
                   "#undef B2_COND_MASK\n#define B2_COND_MASK 0x$(string(cond_mask(gs, us), base = 16))u\n#define B2_NCOND $(length(gs))\n" *
                   vector_fn("b2_vcondition", "g", gs)
            # the right-hand sides read the PRE-event state: copy it first
            asrc = "#define B2_VTERM_MASK 0x$(string(tmask, base = 16))u\n" *
                   "__device__ __forceinline__ void b2_vaffect(real* __restrict__ u_, const real* __restrict__ p, const real t, const int idx) {\n" *
                   "    real u0_[B2_NSTATE]; for (int i = 0; i < B2_NSTATE; i++) u0_[i] = u_[i];\n    const real* u = u0_; real* const w = u_;\n    switch (idx) {\n" *
                   join(["    case $(k - 1): { " * replace(b, r"\bu\[(\d+)\] = " => s"w[\1] = ") * " } break;" for (k, b) in enumerate(branches)], "\n") *
                   "\n    }\n}\n"
        end
    end
    if !isempty(disc)
        d = disc[1]
        cnd = d.condition(us, t, integ0)                      # a symbolic Bool expression (comparison of Nums)
        newu, terminated = trace_affect(d.affect!, us, ps, t)
        dcsrc = scalar_fn("b2_dcondition", "bool", cnd)
        dasrc = affect_fn("b2_daffect", newu)
        term |= terminated ? 2 : 0
    end
    (csrc, asrc, dcsrc, dasrc, term, ip)
end

function model_sources(prob, alg)
    n, m = length(prob.u0), length(prob.p)
    @variables t u[1:n] p[1:m]
    us, ps = collect(u), collect(p)
    du = SciMLBase.isinplace(prob) ? (d = similar(us, Num); prob.f(d, us, ps, t); d) : prob.f(us, ps, t)
    rhs = vector_fn("b2_rhs", "du", du)
    M = prob.f.mass_matrix                              # ODEFunction(f; mass_matrix = M): constant table in the RHS source
    if !(M isa SciMLBase.LinearAlgebra.UniformScaling)
        nameof(typeof(alg)) in (:Rodas4, :Rodas5, :Rodas5P, :FBDF) || error("EnsembleB200: mass_matrix needs Rodas4 / Rodas5 / Rodas5P / FBDF")
        rhs = "#undef B2_HAS_MASS\n#define B2_HAS_MASS 1\nstatic constexpr double B2_MASS_[$(n * n)] = {" *
              join(string.(Float64.(vec(permutedims(Matrix(M))))), ", ") * "};\n" * rhs
    end
    stiff = nameof(typeof(alg)) in (:Rosenbrock23, :Rodas4, :Rodas5, :Rodas5P, :FBDF)   # analytic Jacobian for W
    jac = stiff ? vector_fn("b2_jac", "J", vec(permutedims(Symbolics.jacobian(du, us)))) : nothing   # row-major
    tgrad = stiff ? vector_fn("b2_tgrad", "dT", Symbolics.derivative.(du, t)) : nothing
    noise = prob isa SciMLBase.SDEProblem ?
            vector_fn("b2_noise", "g", SciMLBase.isinplace(prob) ? (gg = similar(us, Num); prob.g(gg, us, ps, t); gg) : prob.g(us, ps, t)) : nothing
    (rhs, jac, tgrad, noise, us, ps, t)
end

cptr(s) = s === nothing ? Cstring(C_NULL) : Base.unsafe_convert(Cstring, s)

"""One device solve of trajectories lo+1 .. lo+N (1-based, like upstream's batches)."""
function solve_batch_b200(eprob, alg, ens::EnsembleB200, model::Ptr{Cvoid}, lo::Int, N::Int, repeat::Int, ts, term::Int, ip::Int, nout::Int;
                          dt, abstol, reltol, adaptive, maxiters, seed, tstops, dtmin, dtmax)
    prob = eprob.prob
    T = eltype(prob.u0); n = length(prob.u0); m = length(prob.p)
    U0 = pinned(T, n, N); P = pinned(T, max(m, 1), N)
    for i in 1:N
        # prob_func(prob, i, repeat) exactly as EnsembleThreads / EnsembleGPUKernel call it (SURVEY 3.3 / 3.4).  SciMLBase 3
        # also exports an EnsembleContext (qa.jl:48); its constructor is not known here, so the (prob, ctx) form is an error
        # instead of a guess.
        applicable(eprob.prob_func, prob, lo + i, repeat) ||
            error("EnsembleB200: prob_func must accept (prob, i, repeat)")
        pi = eprob.prob_func(prob, lo + i, repeat)
        (pi.f === prob.f && pi.tspan == prob.tspan) || error("EnsembleB200: prob_func may only change u0 and p (one compiled kernel per ensemble)")
        U0[:, i] .= pi.u0
        m > 0 && (P[1:m, i] .= pi.p)
    end
    o = Opts(); ccall((:b200ens_opts_init, LIB), Cvoid, (Ref{Opts},), o)
    o.adaptive = adaptive; o.t0, o.t1 = prob.tspan; o.dt = dt
    # scalar tolerances, or one per state component (solve(prob, Rodas5P(); abstol = [1e-8, 1e-14, 1e-6]))
    atolv = abstol isa AbstractVector ? collect(Float64, abstol) : Float64[]
    rtolv = reltol isa AbstractVector ? collect(Float64, reltol) : Float64[]
    o.abstol = isempty(atolv) ? abstol : atolv[1]; o.reltol = isempty(rtolv) ? reltol : rtolv[1]
    o.abstol_vec = isempty(atolv) ? C_NULL : pointer(atolv); o.reltol_vec = isempty(rtolv) ? C_NULL : pointer(rtolv)
    o.maxiters = maxiters; o.seed = seed; o.traj_offset = lo; o.refill_threshold = ens.refill_threshold
    dtmin === nothing || (o.dtmin = dtmin); dtmax === nothing || (o.dtmax = dtmax)
    tsv = sort(unique(collect(Float64, tstops)))             # solve(...; tstops): times the integrator must hit exactly
    o.tstops = isempty(tsv) ? C_NULL : pointer(tsv); o.n_tstops = length(tsv)
    o.event_terminate = term; o.interp_points = ip
    o.device_mask = isempty(ens.devices) ? 0 : reduce(|, UInt32(1) .<< ens.devices)
    out = pinned(T, nout, length(ts), N)                # column-major == [N][n_save][n_out] of the ABI (n_out = n_state, or length(save_idxs))
    rc = Vector{Int32}(undef, N); st = Vector{Stats}(undef, N); tm = Timing()
    GC.@preserve atolv rtolv tsv check(ccall((:b200ens_solve, LIB), Cint,
        (Ptr{Cvoid}, Ref{Opts}, Int64, Ptr{T}, Ptr{T}, Ptr{T}, Int32, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Int32}, Ptr{Stats}, Ref{Timing}),
        model, o, N, U0, P, ts, length(ts), C_NULL, out, C_NULL, rc, st, tm))
    map(1:N) do i
        ui = [out[:, k, i] for k in 1:length(ts)]       # (a real package would reinterpret to SVector without copying)
        SciMLBase.build_solution(prob, alg, ts, ui; retcode = RETCODES[rc[i] + 1])   # UNVERIFIED: stats keyword of build_solution
    end
end

function __solve(eprob::AbstractEnsembleProblem, alg, ens::EnsembleB200; trajectories, batch_size = trajectories, saveat = nothing,
                 dt = 0.0, abstol = 1e-6, reltol = 1e-3, adaptive = true, maxiters = 100_000, seed = UInt64(0),
                 callback = nothing, tstops = Float64[], dtmin = nothing, dtmax = nothing, save_idxs = nothing, kwargs...)
    isempty(kwargs) || error("EnsembleB200: unsupported solve keyword arguments $(collect(keys(kwargs)))")
    haskey(ALG_IDS, nameof(typeof(alg))) || error("EnsembleB200: algorithm $(nameof(typeof(alg))) is not implemented on the device")
    prob = eprob.prob
    prob.kwargs === nothing || isempty(prob.kwargs) || error("EnsembleB200: problem-level kwargs (e.g. a callback stored in the problem) are not supported; pass callback = ... to solve")
    T = eltype(prob.u0); n = length(prob.u0); m = length(prob.p)
    T in (Float32, Float64) || error("EnsembleB200: eltype(u0) must be Float32 or Float64")
    ts = saveat === nothing ? T[prob.tspan...] : saveat isa Number ? collect(T, prob.tspan[1]:saveat:prob.tspan[2]) : collect(T, saveat)
    rhs, jac, tgrad, noise, us, ps, t = model_sources(prob, alg)
    csrc, asrc, dcsrc, dasrc, term, ip = callback_sources(callback, us, ps, t)     # lowers every callback or errors
    model = Ref{Ptr{Cvoid}}(C_NULL); log = Vector{UInt8}(undef, 1 << 16)
    sidx = save_idxs === nothing ? Int32[] : Int32.(collect(save_idxs) .- 1)      # solve(...; save_idxs): 0-based for the ABI
    nout = isempty(sidx) ? n : length(sidx)
    GC.@preserve rhs jac tgrad noise csrc asrc dcsrc dasrc sidx begin
        d = ModelDesc(sizeof(ModelDesc), n, m, T == Float64 ? 1 : 0, ALG_IDS[nameof(typeof(alg))], 0,
                      cptr(rhs), cptr(jac), cptr(tgrad), cptr(noise), cptr(csrc), cptr(asrc), Cstring(C_NULL), cptr(dcsrc), cptr(dasrc),
                      isempty(sidx) ? Ptr{Int32}(C_NULL) : pointer(sidx), length(sidx), 0)
        check(ccall((:b200ens_compile, LIB), Cint, (Ref{ModelDesc}, Ref{Ptr{Cvoid}}, Ptr{UInt8}, Csize_t), d, model, log, length(log)))
    end
    kw = (; dt = Float64(dt), abstol, reltol, adaptive, maxiters, seed, tstops, dtmin, dtmax)
    elapsed = @elapsed begin
        # the batch / output_func / reduction loop of SciMLBase.__solve for ensembles (qa.jl:56,192; SURVEY 3.3), after the gather
        u = eprob.u_init === nothing ? [] : eprob.u_init          # UNVERIFIED: upstream's default u_init
        converged = false
        N = trajectories
        for lo in 0:batch_size:(N - 1)
            nb = min(batch_size, N - lo)
            sols = solve_batch_b200(eprob, alg, ens, model[], lo, nb, 1, ts, term, ip, nout; kw...)
            data = map(1:nb) do j
                out, rerun = eprob.output_func(sols[j], lo + j)
                rep = 1
                while rerun
                    rep += 1
                    rep > 100 && error("EnsembleB200: output_func keeps asking for a rerun (100 repeats)")
                    one = solve_batch_b200(eprob, alg, ens, model[], lo + j - 1, 1, rep, ts, term, ip, nout; kw..., seed = seed + UInt64(rep - 1) * 0x9E3779B97F4A7C15)
                    out, rerun = eprob.output_func(one[1], lo + j)
                end
                out
            end
            u, converged = eprob.reduction(u, data, (lo + 1):(lo + nb))
            converged && break
        end
    end
    ccall((:b200ens_free, LIB), Cvoid, (Ptr{Cvoid},), model[])
    EnsembleSolution(u, elapsed, converged)
end

export EnsembleB200
end # module
