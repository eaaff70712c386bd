"""EXPERIMENTAL adaptive SDE stepping on the GPU (SRIW1 / SOSRA + rejection sampling with memory).
The kernel (kernels/b2_sde_adaptive.cuh) was written against the oracle when the round's GPU budget was almost spent: the
last seconds ran the Float64 SRIW1 case below green on a B200 (and the Float32 case up to its flip-rate threshold, 97.9 %
identical step sequences); the SOSRA case has not run yet.  Until all of them have, these parity tests run only with
B200ENS_EXPERIMENTAL_SDE_ADAPTIVE=1, the same switch that opens the feature in the host API.

Device normals differ from glibc's by ulps (log / sin / cos), so an accept/reject decision can flip on a rare path: the
comparison is per trajectory -- same step counts on almost every path, and on those paths agreement to a tolerance."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200ENS_EXPERIMENTAL_SDE_ADAPTIVE") != "1",
                                 reason="adaptive SDE stepping is experimental (set B200ENS_EXPERIMENTAL_SDE_ADAPTIVE=1)")]


def _ens(B, prob, u0, p):
    return B.EnsembleProblem(prob, u0s=u0, ps=p)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_adaptive_sriw1_gbm_matches_oracle(B, gpu_lib, oracle, dtype):
    from b200ens import workloads as W

    N, tol = 4096, 1e-3 if dtype == np.float64 else 1e-2
    u0, p = W.gbm_params(N, dtype=dtype)
    saveat = np.linspace(0, 1, 5)
    sol = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.SRIW1(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.5,
                  adaptive=True, abstol=tol, reltol=tol, seed=21)
    ref, rc, st, _ = oracle.solve("gbm", "SRIW1", u0, p, (0.0, 1.0), saveat, 0.5, dtype=dtype, seed=21, abstol=tol, reltol=tol,
                                  sde_adaptive=True)
    assert np.all(sol.retcodes == 1) and np.all(rc == 1)
    same = np.all(sol.stats[:, :2] == st[:, :2], axis=1)
    # measured on B200 (round 1): 4096 paths, Float64 > 98 % identical step sequences, Float32 97.9 %
    assert same.mean() > (0.98 if dtype == np.float64 else 0.95), same.mean()
    assert st[:, 1].mean() > 1.0                                     # the large first step is rejected: RSwM is exercised
    a, b = sol.u_array[same].astype(np.float64), ref[same].astype(np.float64)
    assert (np.abs(a - b) / np.abs(b)).max() < (1e-7 if dtype == np.float64 else 2e-3)
    # all paths, whatever their step sequence: the law of GBM (mean exp(mu t))
    assert abs(np.mean(sol.u_array[:, -1, 0] / np.exp(p[:, 0].astype(np.float64))) - 1.0) < 0.1


def test_adaptive_sosra_stochastic_lorenz_matches_oracle(B, gpu_lib, oracle):
    from b200ens import workloads as W

    N, tol = 1024, 1e-2
    u0, p = W.lorenz_additive_params(N)
    saveat = np.linspace(0, 2, 9)
    prob = W.lorenz_additive_problem(tspan=(0.0, 2.0))
    sol = B.solve(_ens(B, prob, u0, p), B.SOSRA(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.25, adaptive=True,
                  abstol=tol, reltol=tol, seed=5)
    ref, rc, st, _ = oracle.solve("lorenz_additive", "SOSRA", u0, p, (0.0, 2.0), saveat, 0.25, seed=5, abstol=tol, reltol=tol,
                                  sde_adaptive=True)
    assert np.all(sol.retcodes == 1) and np.all(rc == 1)
    same = np.all(sol.stats[:, :2] == st[:, :2], axis=1)
    assert same.mean() > 0.9, same.mean()                            # chaotic drift: more flips than on GBM
    a, b = sol.u_array[same], ref[same]
    assert np.median(np.abs(a - b) / (1.0 + np.abs(b))) < 1e-8
