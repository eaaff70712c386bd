"""Host-side mirror of the reference's problem / ensemble / solve interface for EnsembleB200.

Names follow /root/reference/test/qa/qa.jl (ODEProblem :86, SDEProblem :103, EnsembleProblem
:50, EnsembleSolution :52, ContinuousCallback :26, ReturnCode :213, remake :183, solve :192,
Tsit5 :119, Vern7 :126, Rosenbrock23 :98, Rodas5P :97).  `EnsembleB200` is the new sibling of
EnsembleThreads (:56): same `solve(eprob, alg, ensemblealg; trajectories, saveat, dt, abstol,
reltol, callback, maxiters)` call, but every trajectory is integrated on the GPUs through
libb200ens.so.  Semantics the reference's tests pin (test/core.jl): retcode Success (:15),
first saved value == u0 (:18,:34), saveat=0.1 on (0,1) gives 11 points (:93-95),
ContinuousCallback(condition(u,t,integrator), affect!(integrator)) (:69-72), remake (:84).
"""
import enum
import os
import time

import numpy as np

from . import _lib, codegen
from .codegen import terminate_b  # noqa: F401  (re-export: Julia's terminate!)


class ReturnCode(enum.IntEnum):
    Default = 0
    Success = 1
    Terminated = 2
    MaxIters = 3
    DtLessThanMin = 4
    Unstable = 5
    DtNaN = 6
    Failure = 7


# ---------------------------------------------------------------- algorithms
class _Alg:
    name = ""
    adaptive_default = True
    is_sde = False

    def __repr__(self):
        return f"{self.name}()"


class Tsit5(_Alg):
    name = "Tsit5"


class Vern7(_Alg):
    name = "Vern7"


class Rosenbrock23(_Alg):
    name = "Rosenbrock23"


class Rodas4(_Alg):
    name = "Rodas4"


class Rodas5(_Alg):
    name = "Rodas5"


class Rodas5P(_Alg):
    name = "Rodas5P"


class FBDF(_Alg):
    """FBDF (qa.jl:57): variable-order (1..5) fixed-leading-coefficient BDF with a Newton corrector on the analytic Jacobian."""
    name = "FBDF"


class EM(_Alg):
    name = "EM"
    adaptive_default = False
    is_sde = True


class SOSRA(_Alg):
    name = "SOSRA"
    adaptive_default = False
    is_sde = True


class SRIW1(_Alg):
    """Roessler's SRI W1 (StochasticDiffEq `SRIW1`): strong order 1.5 for diagonal noise, fixed dt."""
    name = "SRIW1"
    adaptive_default = False
    is_sde = True


# ---------------------------------------------------------------- problems
class ODEFunction:
    """ODEFunction(f; mass_matrix, jac, tgrad) (qa.jl: exported by the reference): the wrapper upstream uses to attach a
    constant mass matrix (M u' = f, index-1 DAE) or an analytic Jacobian to f.  The device Jacobian and time gradient are
    always derived symbolically from the traced f; a given `jac` is checked against that derivation when the model is
    built (codegen.check_user_jacobian), `tgrad` is accepted for signature compatibility."""

    def __init__(self, f, mass_matrix=None, jac=None, tgrad=None):
        if isinstance(f, ODEFunction):
            mass_matrix = f.mass_matrix if mass_matrix is None else mass_matrix
            jac = f.jac if jac is None else jac
            f = f.f
        self.f, self.mass_matrix, self.jac, self.tgrad = f, mass_matrix, jac, tgrad

    def __call__(self, *args):
        return self.f(*args)


class SDEFunction(ODEFunction):
    """SDEFunction(f, g): drift and diagonal diffusion (qa.jl: exported by the reference)."""

    def __init__(self, f, g):
        super().__init__(f)
        self.g = g


def successful_retcode(x):
    """successful_retcode(sol) / successful_retcode(retcode) (qa.jl: exported by the reference): Success or Terminated."""
    rc = getattr(x, "retcode", x)
    return int(rc) in (int(ReturnCode.Success), int(ReturnCode.Terminated))


class ODEProblem:
    """ODEProblem(f, u0, tspan, p): f(u,p,t) -> du  or in-place f(du,u,p,t) (test/core.jl:22-30)."""

    is_sde = False

    def __init__(self, f, u0, tspan, p=None, g=None, mass_matrix=None):
        self.jac = None
        if isinstance(f, ODEFunction):   # ODEProblem(ODEFunction(f; mass_matrix = M, jac = J), u0, tspan, p)
            mass_matrix = f.mass_matrix if mass_matrix is None else mass_matrix
            self.jac = f.jac
            f = f.f
        self.f = f
        self.g = g
        # ODEFunction(f; mass_matrix = M): M u' = f(u,p,t) with a constant, possibly singular M (index-1 DAE); supported by
        # the Rodas family (the Robertson DAE example of the DifferentialEquations.jl documentation)
        self.mass_matrix = None if mass_matrix is None else np.asarray(mass_matrix, dtype=np.float64)
        self.scalar = np.ndim(u0) == 0
        self.u0 = np.atleast_1d(np.asarray(u0))
        if self.u0.dtype not in (np.float32, np.float64):
            self.u0 = self.u0.astype(np.float64)
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.p = np.zeros(0, self.u0.dtype) if p is None else np.atleast_1d(np.asarray(p, dtype=self.u0.dtype))
        if self.mass_matrix is not None and self.mass_matrix.shape != (self.u0.shape[0],) * 2:
            raise ValueError("mass_matrix must be n_state x n_state")

    def _replace(self, **kw):
        new = object.__new__(type(self))
        new.__dict__.update(self.__dict__)
        if "u0" in kw:
            u0 = np.atleast_1d(np.asarray(kw["u0"], dtype=self.u0.dtype if np.asarray(kw["u0"]).dtype.kind != "f" else None))
            if u0.dtype not in (np.float32, np.float64):
                u0 = u0.astype(np.float64)
            new.u0 = u0
        if "p" in kw:
            new.p = np.atleast_1d(np.asarray(kw["p"], dtype=new.u0.dtype))
        if "tspan" in kw:
            new.tspan = (float(kw["tspan"][0]), float(kw["tspan"][1]))
        bad = set(kw) - {"u0", "p", "tspan"}
        if bad:
            raise TypeError(f"remake: unsupported fields {sorted(bad)}")
        return new


class SDEProblem(ODEProblem):
    """SDEProblem(f, g, u0, tspan, p) with diagonal noise g(u,p,t) (qa.jl:103)."""

    is_sde = True

    def __init__(self, f, g=None, u0=None, tspan=None, p=None):
        if isinstance(f, SDEFunction):   # SDEProblem(SDEFunction(f, g), u0, tspan, p)
            u0, tspan, p = g, u0, tspan
            f, g = f.f, f.g
        super().__init__(f, u0, tspan, p, g=g)


def remake(prob, **kw):
    """remake(prob; u0, p, tspan) (qa.jl:183, test/core.jl:84)."""
    return prob._replace(**kw)


class _SameAffect:
    """Default of affect_neg!: the downcrossing affect is the upcrossing affect (upstream: affect_neg! = affect!)."""

    def __repr__(self):
        return "affect_neg! = affect!"


SAME_AFFECT = _SameAffect()


class ContinuousCallback:
    """ContinuousCallback(condition, affect!, affect_neg! = affect!; interp_points) with condition(u,t,integrator) and
    affect!(integrator) (test/core.jl:69-72): an upcrossing of the condition (negative -> positive) runs
    affect!, a downcrossing affect_neg!; `None` for either one ignores that direction (upstream's `nothing`).  All
    functions must be symbolically traceable (they are emitted as CUDA C)."""

    def __init__(self, condition, affect, affect_neg=SAME_AFFECT, interp_points=10, save_positions=(False, False)):
        self.condition = condition
        self.affect = affect
        self.affect_neg = affect if affect_neg is SAME_AFFECT else affect_neg
        if self.affect is None and self.affect_neg is None:
            raise ValueError("ContinuousCallback: affect! and affect_neg! cannot both be nothing")
        self.interp_points = interp_points
        if tuple(save_positions) != (False, False):
            raise NotImplementedError("EnsembleB200 needs save_positions=(false,false) (saveat output is fixed-size)")

    @property
    def direction(self):
        """0: both directions, +1: upcrossings only, -1: downcrossings only."""
        return 1 if self.affect_neg is None else (-1 if self.affect is None else 0)


class VectorContinuousCallback:
    """VectorContinuousCallback(condition, affect!, len) (qa.jl:124): condition(out, u, t, integrator) fills `len`
    event functions, the earliest root among those that change sign fires and affect!(integrator, idx) receives its
    1-based index.  Both must be symbolically traceable."""

    def __init__(self, condition, affect, len, affect_neg=SAME_AFFECT, interp_points=10, save_positions=(False, False)):
        self.condition = condition
        self.affect = affect
        # affect_neg!: the same function (default) or None = upcrossings only; affect = None with affect_neg = downcrossings only
        self.affect_neg = affect if affect_neg is SAME_AFFECT else affect_neg
        if self.affect is None and self.affect_neg is None:
            raise ValueError("VectorContinuousCallback: affect! and affect_neg! cannot both be nothing")
        if self.affect is not None and self.affect_neg is not None and self.affect_neg is not self.affect:
            raise NotImplementedError("VectorContinuousCallback with an affect_neg! different from affect! (use nothing for "
                                      "one of them, or ContinuousCallbacks)")
        self.len = int(len)
        self.interp_points = interp_points
        if not 1 <= self.len <= 16:
            raise ValueError("VectorContinuousCallback: len must be in 1..16")
        if tuple(save_positions) != (False, False):
            raise NotImplementedError("EnsembleB200 needs save_positions=(false,false) (saveat output is fixed-size)")

    @property
    def direction(self):
        return 1 if self.affect_neg is None else (-1 if self.affect is None else 0)


class DiscreteCallback:
    """DiscreteCallback(condition, affect!) with condition(u,t,integrator)::Bool tested after every accepted step
    (test/core.jl:76-77).  Both functions must be symbolically traceable."""

    def __init__(self, condition, affect, save_positions=(False, False)):
        self.condition = condition
        self.affect = affect
        if tuple(save_positions) != (False, False):
            raise NotImplementedError("EnsembleB200 needs save_positions=(false,false) (saveat output is fixed-size)")


class CallbackSet:
    """CallbackSet(cb...) (qa.jl:24): any number of ContinuousCallbacks plus at most one DiscreteCallback.  Several
    ContinuousCallbacks are one VectorContinuousCallback to the kernel (same upstream semantics: the earliest root among
    the conditions that change sign fires, and only its affect! runs)."""

    def __init__(self, *cbs):
        cont = [c for c in cbs if isinstance(c, (ContinuousCallback, VectorContinuousCallback))]
        self.discrete = [c for c in cbs if isinstance(c, DiscreteCallback)]
        if len(self.discrete) > 1 or len(cont) + len(self.discrete) != len(cbs):
            raise NotImplementedError("EnsembleB200 supports ContinuousCallbacks plus at most one DiscreteCallback")
        if len(cont) > 1:
            if any(isinstance(c, VectorContinuousCallback) for c in cont):
                raise NotImplementedError("a VectorContinuousCallback cannot be combined with further continuous callbacks")
            scalars = list(cont)
            if any(c.direction != scalars[0].direction or (c.direction == 0 and c.affect_neg is not c.affect) for c in scalars):
                raise NotImplementedError("a CallbackSet of several ContinuousCallbacks needs one common direction "
                                          "(all two-sided with affect_neg! = affect!, all upcrossing-only or all downcrossing-only)")
            sdir = scalars[0].direction

            def condition(out, u, t, integrator, _cbs=scalars):
                for k, c in enumerate(_cbs):
                    out[k] = c.condition(u, t, integrator)

            def affect(integrator, idx, _cbs=scalars, _d=sdir):
                (_cbs[idx - 1].affect_neg if _d < 0 else _cbs[idx - 1].affect)(integrator)

            cont = [VectorContinuousCallback(condition, affect if sdir >= 0 else None, len(scalars),
                                             affect_neg=(SAME_AFFECT if sdir == 0 else (None if sdir > 0 else affect)),
                                             interp_points=max(c.interp_points for c in scalars))]
        self.continuous = cont


def _split_callbacks(callback):
    if callback is None:
        return None, None
    if isinstance(callback, (ContinuousCallback, VectorContinuousCallback)):
        return callback, None
    if isinstance(callback, DiscreteCallback):
        return None, callback
    if isinstance(callback, CallbackSet):
        return (callback.continuous[0] if callback.continuous else None), (callback.discrete[0] if callback.discrete else None)
    raise TypeError("callback must be a ContinuousCallback, DiscreteCallback or CallbackSet")


class EnsembleProblem:
    """EnsembleProblem(prob; prob_func, output_func, reduction) (qa.jl:50).

    prob_func(prob, i, repeat) -> problem for trajectory i (1-based like Julia); it may only
    change u0 and p.  Extension (SURVEY 7.3 'host-side prob_func'): `u0s` / `ps` matrices
    [N, n] can be given directly, skipping N host calls of prob_func."""

    def __init__(self, prob, prob_func=None, output_func=None, reduction=None, u_init=None, u0s=None, ps=None,
                 safetycopy=False):
        self.prob = prob
        self.prob_func = prob_func
        self.output_func = output_func     # (sol, i) -> (out, rerun)
        self.reduction = reduction         # (u, batch_data, I) -> (u, converged)
        self.u_init = u_init               # initial value of the reduction (upstream default: empty vector)
        self.u0s = u0s
        self.ps = ps


class EnsembleContext(int):
    """EnsembleContext (qa.jl:48, SciMLBase 3): what prob_func(prob, ctx) / output_func(sol, ctx) receive per trajectory.
    It IS the 1-based trajectory index (an int, so SciMLBase 2 style callbacks that expect `i` keep working) and carries
    `sim_id`, `repeat` and `rng`: a generator seeded by (solve seed, sim_id), i.e. reproducible for any batch size, rank
    layout or rerun order (get_rng(ctx), qa.jl:152)."""

    def __new__(cls, sim_id, repeat=1, seed=0):
        self = super().__new__(cls, int(sim_id))
        self.sim_id, self.repeat, self._seed, self._rng = int(sim_id), int(repeat), int(seed), None
        return self

    @property
    def rng(self):
        if self._rng is None:
            self._rng = np.random.default_rng([self._seed, self.sim_id, self.repeat])
        return self._rng


def has_rng(ctx):
    return isinstance(ctx, EnsembleContext)


def get_rng(ctx):
    return ctx.rng


def _call_prob_func(prob_func, prob, i, repeat, seed):
    """prob_func(prob, i, repeat) (SciMLBase 2, test usage) or prob_func(prob, ctx) (SciMLBase 3): probed by arity, like
    `applicable` in the Julia glue."""
    if codegen._nparams(prob_func) == 2:
        return prob_func(prob, EnsembleContext(i, repeat, seed))
    return prob_func(prob, i, repeat)


class EnsembleB200:
    """The ensemble algorithm: sibling of EnsembleThreads / EnsembleGPUKernel.

    devices: iterable of CUDA device ids (None = all visible); refill_threshold: idle lanes of a
    warp before it fetches new trajectories (0 = auto); stage_outputs: -1 auto / 0 / 1; work_order: -1 auto / 0
    caller's order / 1 integrate in descending expected-work order (scheduling only, identical results)."""

    def __init__(self, devices=None, refill_threshold=0, stage_outputs=-1, fast_math=False,
                 stage_vectors_in_smem=False, work_order=-1, split=None, shard_blocks=0):
        self.devices = devices
        self.shard_blocks = shard_blocks   # multi-device dealing: 0 auto (8 blocks per device, boustrophedon), 1 contiguous ranges
        self.refill_threshold = refill_threshold
        self.stage_outputs = stage_outputs
        self.work_order = work_order
        self.fast_math = fast_math
        self.stage_vectors_in_smem = stage_vectors_in_smem   # ERK k-vectors in shared memory (large n_state)
        self.split = split   # None auto / True / False: components of one trajectory split over the 4 warps of a CTA


# ---------------------------------------------------------------- solutions
class ODESolution:
    def __init__(self, t, u, retcode, stats, scalar=False, dense=None):
        self.t = t
        self.u = u[:, 0] if scalar else u
        self.retcode = ReturnCode(int(retcode))
        self.stats = None if stats is None else dict(zip(("naccept", "nreject", "nf", "nevents"), map(int, stats)))
        self._scalar = scalar
        self._dense = dense

    def __call__(self, t):
        """sol(t): the continuous (dense-output) solution, as solve(...; dense=true) gives upstream
        (test/core.jl:51-58).  Evaluated by the SAME interpolant the saveat path uses (Tsit5: free 4th-order, Vern7:
        order 6, Rosenbrock23: its own, Rodas: derived order 3-4 dense output, FBDF: cubic Hermite): saveat points do not influence the step sequence, so
        re-running this one trajectory on the device with saveat = t returns exactly the value the dense output of the
        original run has at t -- no per-step storage of stage vectors."""
        if self._dense is None:
            raise ValueError("this solution was computed without dense=True")
        tt = np.atleast_1d(np.asarray(t, dtype=np.float64))
        order = np.argsort(tt, kind="stable")
        vals = self._dense(tt[order])
        out = np.empty_like(vals)
        out[order] = vals
        if self._scalar:
            out = out[:, 0]
        return out[0] if np.ndim(t) == 0 else out

    def __getitem__(self, i):
        return self.u[i]

    def __len__(self):
        return len(self.t)


class EnsembleSolution:
    """EnsembleSolution (qa.jl:52): sol.u[i] / sol[i] is trajectory i's ODESolution.  The raw
    gathered arrays stay available as u_array [N, n_save, n_state], retcodes [N], stats [N,4]."""

    def __init__(self, t, u_array, retcodes, stats, elapsed, timing, scalar=False, dense=None):
        self.t = t
        self.u_array = u_array
        self.retcodes = retcodes
        self.stats = stats
        self.elapsedTime = elapsed
        self.timing = timing
        self.converged = bool(np.all((retcodes == ReturnCode.Success) | (retcodes == ReturnCode.Terminated)))
        self._scalar = scalar
        self._dense = dense   # dense(i, times) -> [len(times), n_state] of trajectory i, or None

    def __len__(self):
        return self.u_array.shape[0]

    def __getitem__(self, i):
        dense = None if self._dense is None else (lambda tt, _i=int(i): self._dense(_i, tt))
        if np.ndim(self.t) == 2:   # save_everystep: per-trajectory times [N, capacity], valid entries = naccept + 1
            k = int(self.stats[i, 0]) + 1
            return ODESolution(self.t[i, :k], self.u_array[i, :k], self.retcodes[i], self.stats[i], self._scalar, dense)
        return ODESolution(self.t, self.u_array[i], self.retcodes[i], None if self.stats is None else self.stats[i],
                           self._scalar, dense)

    @property
    def u(self):
        return _LazySeq(self)


class EnsembleSummary:
    """EnsembleSummary (qa.jl:54): per-time-point mean `u` and sample variance `v` over the trajectories that finished
    with Success, computed ON THE DEVICE (EnsembleAnalysis.timestep_meanvar, qa.jl:211) -- the trajectories never travel
    to the host.  `sum`/`sumsq`/`num_monte` are kept so that partial summaries of several ranks can be merged."""

    def __new__(cls, *args, **kw):
        # upstream's constructor EnsembleSummary(sim, t = sim.t; quantiles = [0.05, 0.95]) on gathered trajectories
        if args and isinstance(args[0], EnsembleSolution):
            from .analysis import HostEnsembleSummary

            return HostEnsembleSummary(*args, **kw)
        return super().__new__(cls)

    def __init__(self, t, s, q, count, retcodes, elapsed, timing):
        self.t = t
        self.sum, self.sumsq, self.num_monte = s, q, int(count)
        self.retcodes = retcodes
        self.elapsedTime = elapsed
        self.timing = timing
        self.converged = bool(np.all((retcodes == ReturnCode.Success) | (retcodes == ReturnCode.Terminated)))

    def _need_members(self):
        if self.num_monte == 0:
            raise ValueError("EnsembleSummary: no trajectory finished with a successful retcode (Success / Terminated); "
                             f"retcodes seen: {sorted(set(int(r) for r in np.unique(self.retcodes)))}")

    @property
    def u(self):
        self._need_members()
        return self.sum / self.num_monte

    @property
    def v(self):
        self._need_members()
        n = self.num_monte
        return (self.sumsq - self.sum * self.sum / max(n, 1)) / max(n - 1, 1)

    def merge(self, other):
        """Combine with the summary of another shard (sums and counts add)."""
        return EnsembleSummary(self.t, self.sum + other.sum, self.sumsq + other.sumsq, self.num_monte + other.num_monte,
                               np.concatenate([self.retcodes, other.retcodes]), max(self.elapsedTime, other.elapsedTime),
                               self.timing)


class _LazySeq:
    def __init__(self, es):
        self._es = es

    def __len__(self):
        return len(self._es)

    def __getitem__(self, i):
        return self._es[i]

    def __iter__(self):
        return (self._es[i] for i in range(len(self._es)))


# ---------------------------------------------------------------- model building (codegen + NVRTC)
_model_cache = {}


def build_model(prob, alg, callback=None, fast_math=False, ksmem=False, split=None, sde_adaptive=False, save_idxs=None):
    """Trace prob.f (and g / callback), emit CUDA C, JIT it for sm_100a.  Cached per function objects."""
    n, m = prob.u0.shape[0], prob.p.shape[0]
    dtype = prob.u0.dtype
    mm = getattr(prob, "mass_matrix", None)
    if save_idxs is not None:
        save_idxs = tuple(int(i) for i in (save_idxs if np.ndim(save_idxs) else [save_idxs]))
        if not save_idxs or len(set(save_idxs)) != len(save_idxs) or min(save_idxs) < 0 or max(save_idxs) >= n:
            raise ValueError(f"save_idxs must be distinct 0-based components of the state (length {n})")
    key = (id(prob.f), id(prob.g), n, m, dtype.str, alg.name, id(callback), fast_math, ksmem, split,
           None if mm is None else mm.tobytes(), sde_adaptive, os.environ.get("B200ENS_DEFINES", ""), save_idxs)
    hit = _model_cache.get(key)
    # the entry keeps STRONG references to f, g and the callback, so none of their ids can be recycled for another object
    # while it is cached; a hit still has to be the very same objects
    if hit is not None and hit[1] is prob.f and hit[2] is prob.g and hit[3] is callback:
        return hit[0]
    exprs, usyms, _, tsym = codegen.trace_vector_fn(prob.f, n, m)
    srcs = {"rhs_src": codegen.emit_rhs(exprs)}
    if mm is not None:
        if alg.name not in ("Rodas4", "Rodas5", "Rodas5P", "FBDF"):
            raise NotImplementedError(f"mass_matrix is supported by Rodas4 / Rodas5 / Rodas5P / FBDF, not by {alg.name}")
        if callback is not None:
            raise NotImplementedError("callbacks on mass-matrix problems (their interpolant needs u', which M u' = f does "
                                      "not give for algebraic components)")
        srcs["rhs_src"] = codegen.emit_mass(mm) + srcs["rhs_src"]
    if alg.name in ("Rosenbrock23", "Rodas4", "Rodas5", "Rodas5P", "FBDF"):
        if getattr(prob, "jac", None) is not None:
            codegen.check_user_jacobian(prob.jac, exprs, usyms, n, m)
        srcs["jac_src"] = codegen.emit_jac(exprs, usyms)
        srcs["tgrad_src"] = codegen.emit_tgrad(exprs, tsym)
    if alg.is_sde:
        if not prob.is_sde:
            raise TypeError(f"{alg.name} needs an SDEProblem")
        gex, _, _, _ = codegen.trace_vector_fn(prob.g, n, m)
        if alg.name == "SOSRA" and any(e.has(*usyms) for e in gex):
            raise ValueError("SOSRA is for additive noise only: g(u,p,t) must not depend on u (SURVEY A.9)")
        srcs["noise_src"] = codegen.emit_noise(gex)
    terminate = 0
    ccb, dcb = _split_callbacks(callback)
    if ccb is not None:
        emit = codegen.emit_vector_callback if isinstance(ccb, VectorContinuousCallback) else codegen.emit_callback
        srcs["condition_src"], srcs["affect_src"], term = emit(ccb, n, m)
        terminate |= int(term) & 5   # bit 0: affect! terminates, bit 2: affect_neg! does (opts.event_terminate)
    if dcb is not None:
        srcs["dcondition_src"], srcs["daffect_src"], term = codegen.emit_discrete_callback(dcb, n, m)
        terminate |= 2 if term else 0
    model = _lib.Model(n, m, dtype, alg.name, name=getattr(prob.f, "__name__", "model"), fast_math=fast_math,
                       ksmem=ksmem, split=split, sde_adaptive=sde_adaptive, save_idxs=save_idxs, **srcs)
    model.sources = srcs
    model.event_terminate = terminate
    _model_cache[key] = (model, prob.f, prob.g, callback)
    return model


def _saveat_array(saveat, tspan, dtype, save_start=None, save_end=None):
    """The save grid of solve(...; saveat, save_start, save_end): a number is a spacing and includes both
    ends of tspan (test/core.jl:93-95); a vector is taken as given; save_start / save_end = True / False force the
    end point of tspan in or out (None: upstream's default, i.e. whatever the grid says)."""
    t0, t1 = tspan
    if saveat is None:
        ts = np.array([t0, t1], dtype=np.float64)
    elif np.ndim(saveat) == 0:
        step = float(saveat)
        k = int(np.floor((t1 - t0) / step * (1 + 1e-12) + 1e-9))
        ts = t0 + step * np.arange(k + 1)
        ts[np.abs(ts - t1) < 1e-12 * max(1.0, abs(t1))] = t1
        if ts[-1] < t1:
            ts = np.append(ts, t1)
    else:
        ts = np.asarray(saveat, dtype=np.float64)
        if ts.size and (np.any(np.diff(ts) <= 0) or ts[0] < t0 or ts[-1] > t1):
            raise ValueError("saveat must be strictly increasing and inside tspan")
    for flag, tend, front in ((save_start, t0, True), (save_end, t1, False)):
        if flag is None:
            continue
        has = ts.size > 0 and (ts[0] if front else ts[-1]) == tend
        if flag and not has:
            ts = np.concatenate(([tend], ts)) if front else np.concatenate((ts, [tend]))
        elif not flag and has:
            ts = ts[1:] if front else ts[:-1]
    return ts.astype(dtype)


def _pack(eprob, N, dtype, lo=0, repeat=1, seed=0):
    """Run prob_func on the host (as EnsembleThreads / EnsembleGPUKernel do) for trajectories lo+1 .. lo+N (1-based
    like Julia) -> u0 [N,n], p [N,m]."""
    prob = eprob.prob
    n, m = prob.u0.shape[0], prob.p.shape[0]
    if eprob.u0s is not None or eprob.ps is not None:
        u0 = np.broadcast_to(prob.u0, (N, n)) if eprob.u0s is None else np.asarray(eprob.u0s).reshape(-1, n)[lo:lo + N]
        p = np.broadcast_to(prob.p, (N, m)) if eprob.ps is None else np.asarray(eprob.ps).reshape(-1, m)[lo:lo + N]
        return u0, p
    u0 = np.empty((N, n), dtype=dtype)
    p = np.empty((N, m), dtype=dtype)
    if eprob.prob_func is None:
        u0[:] = prob.u0
        p[:] = prob.p
        return u0, p
    for i in range(N):
        pi = _call_prob_func(eprob.prob_func, prob, lo + i + 1, repeat, seed)
        if pi.f is not prob.f or pi.tspan != prob.tspan:
            raise ValueError("EnsembleB200: prob_func may only change u0 and p (one compiled kernel per ensemble)")
        u0[i] = pi.u0
        p[i] = pi.p
    return u0, p


class ReducedEnsembleSolution:
    """EnsembleSolution of a run with output_func / reduction / batch_size: `u` is what the reduction built
    (upstream: EnsembleSolution(u, elapsedTime, converged)), `converged` the flag the last reduction call returned."""

    def __init__(self, u, elapsed, converged, batches):
        self.u = u
        self.elapsedTime = elapsed
        self.converged = converged
        self.batches = batches     # number of batches actually solved (early exit when the reduction converged)

    def __len__(self):
        return len(self.u)

    def __getitem__(self, i):
        return self.u[i]


def solve(prob, alg, ensemblealg=None, trajectories=None, batch_size=None, **kw):
    """solve(eprob, alg, EnsembleB200(); trajectories, batch_size=trajectories, saveat, dt, abstol, reltol, ...).

    Mirrors SciMLBase.__solve(::AbstractEnsembleProblem, alg, ensemblealg; trajectories, batch_size) (qa.jl:56,192;
    SURVEY 3.3): the trajectories are solved in batches of `batch_size`; every solution goes through
    `output_func(sol, i) -> (out, rerun)` (rerun: the trajectory is solved again with prob_func(prob, i, repeat + 1)),
    every batch through `u, converged = reduction(u, batch_data, I)` starting from `u_init`, and the loop stops early
    when the reduction reports convergence.  Without output_func / reduction / batch_size the whole ensemble is ONE
    device solve and the raw EnsembleSolution (arrays) is returned.  A plain ODEProblem/SDEProblem is solved as a
    one-trajectory ensemble and returns its ODESolution."""
    eprob = prob if isinstance(prob, EnsembleProblem) else None
    if eprob is None or (eprob.output_func is None and eprob.reduction is None and batch_size is None) or kw.get("summary"):
        return _solve_once(prob, alg, ensemblealg, trajectories, **kw)
    if trajectories is None:
        raise TypeError("solve(::EnsembleProblem, ...) needs `trajectories`")
    N = int(trajectories)
    bs = N if batch_size is None else max(1, int(batch_size))
    output_func = eprob.output_func or (lambda sol, i: (sol, False))
    reduction = eprob.reduction or (lambda u, data, I: (u + list(data), False))
    u = [] if eprob.u_init is None else eprob.u_init
    t_all = time.perf_counter()
    converged, batches = False, 0
    for lo in range(0, N, bs):
        n_b = min(bs, N - lo)
        bsol = _solve_once(eprob, alg, ensemblealg, n_b, _lo=lo, **kw)
        data = []
        for j in range(n_b):
            seed = int(kw.get("seed") or 0)
            out, rerun = output_func(bsol[j], EnsembleContext(lo + j + 1, 1, seed))
            repeat = 1
            while rerun:
                repeat += 1
                if repeat > 100:
                    raise RuntimeError("output_func keeps asking for a rerun (100 repeats)")
                one = _solve_once(eprob, alg, ensemblealg, 1, _lo=lo + j, _repeat=repeat, **kw)
                out, rerun = output_func(one[0], EnsembleContext(lo + j + 1, repeat, seed))
            data.append(out)
        u, converged = reduction(u, data, range(lo + 1, lo + n_b + 1))
        batches += 1
        if converged:
            break
    return ReducedEnsembleSolution(u, time.perf_counter() - t_all, bool(converged), batches)


def _solve_once(prob, alg, ensemblealg=None, trajectories=None, saveat=None, dt=None, abstol=None, reltol=None,
                adaptive=None, callback=None, maxiters=None, save_everystep=None, dense=False, seed=0, dW=None,
                save_tstops=None, summary=False, tstops=None, save_start=None, save_end=None, dtmin=None, dtmax=None,
                qmin=None, qmax=None, gamma=None, beta1=None, beta2=None, qoldinit=None, save_idxs=None, _lo=0, _repeat=1,
                **kwargs):
    """One device solve of trajectories _lo+1 .. _lo+trajectories of the ensemble."""
    if kwargs:
        raise TypeError(f"solve: unsupported keyword arguments {sorted(kwargs)}")
    # save_everystep: upstream's default when no saveat is given.  Here: the default for a single problem without saveat
    # (sol.t / sol.u hold every accepted step, like upstream); ensembles keep the fixed-size saveat output unless asked.
    is_single = not isinstance(prob, EnsembleProblem)
    if save_everystep is None:
        save_everystep = bool(is_single and saveat is None and not alg.is_sde and not summary)
    if save_everystep and (alg.is_sde or summary or dW is not None):
        raise NotImplementedError("save_everystep with SDE steppers / summary: pass the time grid as saveat instead")
    if dense and save_tstops:
        raise ValueError("dense=True needs interpolated saves (save_tstops=False): tstops would change the step sequence")
    if (alg.name == "FBDF" and getattr(prob.prob if isinstance(prob, EnsembleProblem) else prob, "mass_matrix", None) is not None
            and (dense or save_tstops is False or save_tstops == 0 and save_tstops is not None)):
        raise NotImplementedError("FBDF on a mass-matrix problem saves at tstops only (its Hermite dense output takes f for u'); "
                                  "the Rodas family interpolates DAEs through its own dense output")
    if dense and dW is not None:
        raise NotImplementedError("dense=True with injected noise increments")
    single = not isinstance(prob, EnsembleProblem)
    eprob = EnsembleProblem(prob) if single else prob
    if single:
        trajectories = 1
    if ensemblealg is None:
        ensemblealg = EnsembleB200()
    if not isinstance(ensemblealg, EnsembleB200):
        raise TypeError("this back-end only implements EnsembleB200()")
    if trajectories is None:
        raise TypeError("solve(::EnsembleProblem, ...) needs `trajectories`")
    base = eprob.prob
    N = int(trajectories)
    dtype = base.u0.dtype
    if adaptive is None:
        adaptive = alg.adaptive_default
    sde_adaptive = bool(alg.is_sde and adaptive)
    if sde_adaptive:
        if alg.name not in ("SRIW1", "SOSRA"):
            raise NotImplementedError(f"{alg.name} has no embedded error estimate: adaptive stepping needs SRIW1 or SOSRA")
        if dt is None:
            raise ValueError("adaptive SDE stepping needs an initial dt")
        # dW, when given, is the trajectory's stream of STANDARD NORMALS [N, len] (consumed in order by the rejection
        # sampling with memory: 2n per fresh step, bridge draw or rejection) -- the pathwise-parity hook
        abstol = 1e-2 if abstol is None else abstol      # StochasticDiffEq's defaults
        reltol = 1e-2 if reltol is None else reltol
    if dt is None:
        if not adaptive:
            raise ValueError("fixed-step solves need dt")
        dt = 0.0   # automatic per-trajectory initial step on the device
    model = build_model(base, alg, callback, ensemblealg.fast_math,
                        ensemblealg.stage_vectors_in_smem, False if save_everystep else ensemblealg.split,
                        sde_adaptive=sde_adaptive, save_idxs=save_idxs)
    ts = _saveat_array(saveat, base.tspan, dtype, save_start, save_end)
    t_pack = time.perf_counter()
    u0, p = _pack(eprob, N, dtype, _lo, _repeat, int(seed or 0))
    t_pack = time.perf_counter() - t_pack

    o = _lib.default_opts()
    o.adaptive = int(bool(adaptive))
    o.t0, o.t1, o.dt = base.tspan[0], base.tspan[1], float(dt)
    # abstol / reltol: scalars, or one value per state component (solve(prob, Rodas5P(); abstol = [1e-8, 1e-14, 1e-6]))
    tol_keep = []
    for name, val in (("abstol", abstol), ("reltol", reltol)):
        if val is None:
            continue
        if np.ndim(val) == 0:
            setattr(o, name, float(val))
        else:
            vec = np.ascontiguousarray(val, dtype=np.float64)
            if vec.shape != (base.u0.shape[0],):
                raise ValueError(f"{name} must be a scalar or have one entry per state component")
            tol_keep.append(vec)
            setattr(o, name + "_vec", vec.ctypes.data_as(_lib.C.POINTER(_lib.C.c_double)))
    o._tol_keep = tol_keep   # the arrays must outlive every solve that uses these options (dense re-solves too)
    if maxiters is not None:
        o.maxiters = int(maxiters)
    # step-size control keywords of solve: None = upstream's default, chosen by the library
    for name, val in (("dtmin", dtmin), ("dtmax", dtmax), ("qmin", qmin), ("qmax", qmax), ("gamma", gamma), ("beta1", beta1),
                      ("beta2", beta2), ("qoldinit", qoldinit)):
        if val is not None:
            if not float(val) >= 0:
                raise ValueError(f"{name} must be >= 0")
            setattr(o, name, float(val))
    if tstops is not None and len(tstops):
        # solve(...; tstops = [...]): times the integrator must hit exactly (handle_tstop!); with a DiscreteCallback whose
        # condition is t == tstop this is upstream's dosing idiom
        if alg.is_sde:
            raise NotImplementedError("tstops with an SDE stepper")
        tsv = np.ascontiguousarray(np.sort(np.unique(np.asarray(tstops, dtype=np.float64))))
        tol_keep.append(tsv)
        o.tstops = tsv.ctypes.data_as(_lib.C.POINTER(_lib.C.c_double))
        o.n_tstops = int(tsv.shape[0])
    # an output_func rerun (repeat > 1) redraws the noise like upstream does: the Philox key depends on (seed, repeat)
    o.seed = (int(seed) + (int(_repeat) - 1) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    o.traj_offset = int(_lo)    # global trajectory index of this batch's first trajectory (Philox counter base)
    o.noise_injected = 0 if dW is None else 1
    if sde_adaptive and dW is not None:
        dW = np.ascontiguousarray(dW, dtype=dtype).reshape(N, -1)
        o.noise_stream_len = int(dW.shape[1])
    if callback is not None:
        o.event_terminate = int(model.event_terminate)
        ccb, _ = _split_callbacks(callback)
        if ccb is not None:
            o.interp_points = int(ccb.interp_points)
    if save_tstops is not None:
        o.save_tstops = int(save_tstops)
    if dense:
        o.save_tstops = 0   # every save point (now and in sol(t)) goes through the interpolant
    if ensemblealg.devices is not None:
        mask = 0
        for g in ensemblealg.devices:
            mask |= 1 << int(g)
        o.device_mask = mask
    o.refill_threshold = int(ensemblealg.refill_threshold)
    o.stage_outputs = int(ensemblealg.stage_outputs)
    o.work_order = int(ensemblealg.work_order)
    o.shard_blocks = int(ensemblealg.shard_blocks)

    if summary:  # EnsembleAnalysis.timestep_meanvar on the device: returns an EnsembleSummary
        t_solve = time.perf_counter()
        s_, q_, cnt, rc, tm = model.solve_moments(o, u0, p, ts, dW=dW)
        timing = tm.asdict()
        timing["prob_func_s"] = t_pack
        return EnsembleSummary(ts, s_, q_, cnt, rc, time.perf_counter() - t_solve, timing)
    t_solve = time.perf_counter()
    if save_everystep:
        # every accepted step: fixed capacity per trajectory, doubled until the longest trajectory fits
        cap = 256
        while True:
            out, ts, rc, stats, tm = model.solve_everystep(o, u0, p, cap)
            need = int(stats[:, 0].max()) + 1
            if need <= cap:
                break
            cap = 1 << int(np.ceil(np.log2(need)))
        if saveat is not None:
            raise ValueError("save_everystep=True and saveat are mutually exclusive on this back-end")
    else:
        out, rc, stats, tm = model.solve(o, u0, p, ts, dW=dW)
    elapsed = time.perf_counter() - t_solve
    timing = tm.asdict()
    timing["prob_func_s"] = t_pack
    dense_fn = None
    if dense:
        def dense_fn(i, tt, _o=o, _u0=u0, _p=p, _off=int(o.traj_offset)):
            tt = np.asarray(tt, dtype=np.float64)
            if tt.size and (tt.min() < base.tspan[0] or tt.max() > base.tspan[1]):
                raise ValueError("sol(t): t outside tspan")
            o2 = _lib.copy_opts(_o)
            o2.traj_offset = _off + i            # same Philox stream as in the ensemble run (SDE paths)
            o2.work_order = 0
            out1, rc1, _, _ = model.solve(o2, _u0[i:i + 1], _p[i:i + 1], tt.astype(dtype))
            return out1[0].astype(np.float64) if dtype != np.float64 else out1[0]
    sol = EnsembleSolution(ts, out, rc, stats, elapsed, timing, scalar=base.scalar, dense=dense_fn)
    if single:
        return sol[0]
    return sol
