"""Adaptive SDE stepping on the GPU (SRIW1 / SOSRA + rejection sampling with memory RSwM1, SURVEY 8f item 3;
SDEProblem is exported at /root/reference/test/qa/qa.jl:103, the adaptive loop lives in StochasticDiffEq).

Two kinds of test, like the fixed-step SDE tests:
  * INJECTED normals: host-generated standard normals, the same array for the kernel and the oracle, consumed in order by
    RSwM (fresh step / bridge draw / rejection).  Every arithmetic operation is then shared, so the accept / reject
    sequence must be IDENTICAL on every path and the saved values agree to rounding of the model functions.
  * DEVICE Philox normals: the device's log / sin / cos differ from glibc's by ulps, so a decision can flip on a rare
    path: >= 98 % identical sequences (measured on B200 in round 1: 98.0-98.4 % Float64, 97.9 % Float32 -- the
    Float32 bar is therefore on the injected test, which is exact), and the law of the solution on all paths.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ens(B, prob, u0, p):
    return B.EnsembleProblem(prob, u0s=u0, ps=p)


def _normals(N, length, dtype, seed):
    return np.random.default_rng(seed).standard_normal((N, length)).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_adaptive_sriw1_gbm_injected_normals_identical_sequences(B, gpu_lib, oracle, dtype):
    from b200ens import workloads as W

    N, tol = 4096, 1e-3 if dtype == np.float64 else 1e-2
    u0, p = W.gbm_params(N, dtype=dtype)
    saveat = np.linspace(0, 1, 5)
    z = _normals(N, 4096, dtype, 77)
    sol = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.SRIW1(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.5,
                  adaptive=True, abstol=tol, reltol=tol, dW=z)
    ref, rc, st, _ = oracle.solve("gbm", "SRIW1", u0, p, (0.0, 1.0), saveat, 0.5, dtype=dtype, abstol=tol, reltol=tol,
                                  sde_adaptive=True, dW=z)
    assert np.all(sol.retcodes == 1) and np.all(rc == 1)
    assert st[:, 1].mean() > 1.0                                     # the large first step is rejected: RSwM is exercised
    assert np.array_equal(sol.stats[:, :2], st[:, :2])               # 100 %: every accept / reject decision
    a, b = sol.u_array.astype(np.float64), ref.astype(np.float64)
    assert (np.abs(a - b) / np.abs(b)).max() < (1e-12 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_adaptive_sosra_stochastic_lorenz_injected_normals_identical_sequences(B, gpu_lib, oracle, dtype):
    from b200ens import workloads as W

    N, tol = 512, 1e-2
    u0, p = W.lorenz_additive_params(N, dtype=dtype)
    saveat = np.linspace(0, 1, 5)
    prob = W.lorenz_additive_problem(dtype, tspan=(0.0, 1.0))
    z = _normals(N, 10 * 4096, dtype, 78)    # up to ~6000 accepted steps + ~9 rejections per path, 6 normals per draw
    sol = B.solve(_ens(B, prob, u0, p), B.SOSRA(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.25, adaptive=True,
                  abstol=tol, reltol=tol, dW=z)
    ref, rc, st, _ = oracle.solve("lorenz_additive", "SOSRA", u0, p, (0.0, 1.0), saveat, 0.25, dtype=dtype, abstol=tol, reltol=tol,
                                  sde_adaptive=True, dW=z)
    assert np.array_equal(sol.retcodes, rc) and np.mean(rc == 1) > 0.95   # (a path that outruns the stream: Failure in both)
    assert np.array_equal(sol.stats[:, :2], st[:, :2])
    ok = rc == 1
    a, b = sol.u_array[ok].astype(np.float64), ref[ok].astype(np.float64)
    assert np.abs(a - b).max() / (1.0 + np.abs(b).max()) < (1e-12 if dtype == np.float64 else 1e-5)


def test_adaptive_injected_stream_too_short_fails_loudly(B, gpu_lib, oracle):
    """A trajectory that runs out of injected normals ends with ReturnCode.Failure (7) in the kernel and in the oracle."""
    from b200ens import workloads as W

    N = 64
    u0, p = W.gbm_params(N)
    saveat = np.linspace(0, 1, 3)
    z = _normals(N, 8, np.float64, 5)     # 2n = 2 normals per draw: four draws, far too few for tol 1e-4
    sol = B.solve(_ens(B, W.gbm_problem(), u0, p), B.SRIW1(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.5,
                  adaptive=True, abstol=1e-4, reltol=1e-4, dW=z)
    ref, rc, st, _ = oracle.solve("gbm", "SRIW1", u0, p, (0.0, 1.0), saveat, 0.5, abstol=1e-4, reltol=1e-4, sde_adaptive=True, dW=z)
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 7)
    assert np.array_equal(sol.stats[:, :2], st[:, :2])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_adaptive_sriw1_gbm_device_philox_matches_oracle(B, gpu_lib, oracle, dtype):
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
    # device vs glibc normals differ by ulps: a rare path flips a decision.  Float64: >= 98 %.  Float32 device normals
    # carry ~1e-6 relative differences; its exact bar is the injected test above, here the flip rate is reported
    print(f"adaptive SRIW1 {np.dtype(dtype).name}: identical accept/reject sequences on {same.mean():.4f} of {N} paths")
    assert same.mean() > (0.98 if dtype == np.float64 else 0.95), same.mean()
    a, b = sol.u_array[same].astype(np.float64), ref[same].astype(np.float64)
    rel = (np.abs(a - b) / np.abs(b)).max(axis=(1, 2))
    # (equal COUNTS do not prove equal sequences: a rare Float32 path takes the same number of steps at different times)
    assert np.quantile(rel, 0.99) < (1e-7 if dtype == np.float64 else 2e-3) and np.median(rel) < (1e-9 if dtype == np.float64 else 1e-5)
    # all paths, whatever their step sequence: the law of GBM (mean exp(mu t))
    assert abs(np.mean(sol.u_array[:, -1, 0] / np.exp(p[:, 0].astype(np.float64))) - 1.0) < 0.1


def test_adaptive_sosra_stochastic_lorenz_device_philox_matches_oracle(B, gpu_lib, oracle):
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
    print(f"adaptive SOSRA float64: identical accept/reject sequences on {same.mean():.4f} of {N} paths")
    assert same.mean() > 0.9, same.mean()                            # chaotic drift: more flips than on GBM
    a, b = sol.u_array[same], ref[same]
    assert np.median(np.abs(a - b) / (1.0 + np.abs(b))) < 1e-8
