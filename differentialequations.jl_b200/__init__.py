"""b200ens -- host-side mirror of the DifferentialEquations.jl ensemble interface for the
B200-native back-end.

Julia is not available in this image, so the host layer that the north star asks for in
Julia (julia/EnsembleB200.jl, shown in INTEGRATION.md) is mirrored here in Python with the
SAME names, argument meaning and error behaviour as the reference's exports
(/root/reference/test/qa/qa.jl:3-217), so that the parity tests read like the reference's
own tests (/root/reference/test/core.jl):

    prob  = ODEProblem(lorenz, u0, tspan, p)
    eprob = EnsembleProblem(prob, prob_func=lambda prob, i, repeat: remake(prob, p=...))
    sol   = solve(eprob, Tsit5(), EnsembleB200(), trajectories=N, saveat=1.0, dt=0.1,
                  abstol=1e-6, reltol=1e-3)

Everything numerical happens in libb200ens.so (csrc/), reached through the C ABI of
include/b200ens.h; this module only traces the model functions to CUDA C (codegen.py),
packs the per-trajectory u0/p matrices and wraps the outputs.  No CPU fallback exists.
"""
from . import _lib, codegen
from . import analysis as EnsembleAnalysis
from ._lib import B200EnsError, Model, pinned_empty
from .api import (EM, FBDF, SOSRA, SRIW1, CallbackSet, ContinuousCallback, DiscreteCallback, EnsembleB200, EnsembleContext, EnsembleProblem, get_rng, has_rng, EnsembleSolution,
                  EnsembleSummary, ODEFunction, ODEProblem, SDEFunction, successful_retcode,
                  ODESolution, ReturnCode, Rodas4, Rodas5, Rodas5P, Rosenbrock23, SDEProblem, Tsit5, VectorContinuousCallback, Vern7,
                  ReducedEnsembleSolution, build_model, remake, solve, terminate_b)

__all__ = [
    "ODEProblem", "SDEProblem", "ODEFunction", "SDEFunction", "successful_retcode", "EnsembleProblem", "EnsembleB200", "EnsembleContext", "get_rng", "has_rng", "EnsembleSolution", "EnsembleSummary", "ODESolution", "ReturnCode",
    "ContinuousCallback", "VectorContinuousCallback", "DiscreteCallback", "CallbackSet", "remake", "solve", "terminate_b", "Tsit5", "Vern7", "Rosenbrock23", "FBDF", "Rodas4", "Rodas5",
    "Rodas5P", "EM", "SOSRA", "SRIW1", "build_model", "Model", "B200EnsError", "pinned_empty", "EnsembleAnalysis", "ReducedEnsembleSolution",
]
