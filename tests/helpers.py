"""Shared helpers of the parity tests (test infrastructure)."""
import os
import tempfile

import numpy as np

_WORK = os.path.join(tempfile.gettempdir(), "b200ens_test_models")


def oracle_fns(oracle, B, model, f64=True):
    """Function pointers for the oracle built from the SAME emitted model source the GPU kernel was JIT-compiled
    from (g++ -ffp-contract=off), so both sides evaluate one expression tree."""
    src = B.codegen.host_wrapper_source([model.sources.get(k) for k in
                                         ("rhs_src", "jac_src", "tgrad_src", "noise_src", "condition_src", "affect_src",
                                          "dcondition_src", "daffect_src")])
    dll = oracle.compile_host_model(src, "m", _WORK)
    fns = oracle.fns_from_host_model(dll, f64)
    fns["_dll"] = dll
    return fns


def within_tol(a, b, abstol, reltol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = abstol + reltol * np.abs(b)
    ok = (err <= tol) | (np.isnan(a) & np.isnan(b))
    return bool(np.all(ok)), float(np.nanmax(err / tol)) if err.size else 0.0
