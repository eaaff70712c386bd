"""GPU: ensemble moments accumulated INSIDE the solve kernel (b200ens_solve_moments with B200ENS_FUSE_MOMENTS=1, or
automatically for rows of >= 4 MB per trajectory): per-(save point, component) sums over the successful trajectories must equal the statistics of
the full output, for the one-thread kernel and for the split kernel (16-species network with its callback), and a chunk
in which a trajectory fails must fall back to the out_u path so that the failure does not count.
Reference semantics: SciMLBase.EnsembleAnalysis timestep_meanvar / EnsembleSummary (/root/reference/test/qa/qa.jl:211)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _summary(B, eprob, alg, N, env, **kw):
    old = os.environ.get("B200ENS_FUSE_MOMENTS")
    if env is None:
        os.environ.pop("B200ENS_FUSE_MOMENTS", None)
    else:
        os.environ["B200ENS_FUSE_MOMENTS"] = env
    try:
        return B.solve(eprob, alg, B.EnsembleB200(devices=[0]), trajectories=N, summary=True, **kw)
    finally:
        if old is None:
            os.environ.pop("B200ENS_FUSE_MOMENTS", None)
        else:
            os.environ["B200ENS_FUSE_MOMENTS"] = old


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_moments_one_thread_kernel(B, gpu_lib, dtype):
    from b200ens import workloads as W

    N = 20000
    saveat = np.linspace(0.0, 10.0, 401)      # 401 x 3 = 1203 values per trajectory
    u0, p = W.lorenz_params(N, "random", seed=23, dtype=dtype)
    kw = dict(saveat=saveat, dt=0.1, abstol=1e-6, reltol=1e-3)
    full = B.solve(B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), trajectories=N, **kw)
    assert np.all(full.retcodes == 1)
    ref = full.u_array.astype(np.float64)
    fused = _summary(B, B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0, ps=p), B.Tsit5(), N, "1", **kw)
    twopass = _summary(B, B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0, ps=p), B.Tsit5(), N, "0", **kw)
    # the fused run launches no second-pass kernel
    assert fused.timing["launches"] < twopass.timing["launches"]
    for s in (fused, twopass):
        assert s.num_monte == N and np.array_equal(s.retcodes, full.retcodes)
        assert np.allclose(s.u, ref.mean(axis=0), rtol=1e-11, atol=1e-11)
        assert np.allclose(s.v, ref.var(axis=0, ddof=1), rtol=1e-8, atol=1e-11)


def test_fused_moments_fall_back_when_a_trajectory_fails(B, gpu_lib):
    from b200ens import workloads as W

    N = 6000
    saveat = np.linspace(0.0, 10.0, 401)
    u0, p = W.lorenz_params(N, "random", seed=29)
    p[17] = np.nan                                   # DtNaN after the u0 row has already been accumulated
    kw = dict(saveat=saveat, dt=0.1, abstol=1e-6, reltol=1e-3)
    full = B.solve(B.EnsembleProblem(W.lorenz_problem(), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), trajectories=N, **kw)
    ok = full.retcodes == 1
    assert ok.sum() == N - 1
    s = _summary(B, B.EnsembleProblem(W.lorenz_problem(), u0s=u0, ps=p), B.Tsit5(), N, "1", **kw)
    assert s.num_monte == N - 1 and np.array_equal(s.retcodes, full.retcodes)
    assert np.allclose(s.u, full.u_array[ok].mean(axis=0), rtol=1e-11, atol=1e-11)
    assert np.allclose(s.v, full.u_array[ok].var(axis=0, ddof=1), rtol=1e-8, atol=1e-11)


def test_fused_moments_split_kernel_with_callback(B, gpu_lib):
    """Config 5's kernel: 16 species split over the four warps, ContinuousCallback, 101 save points = 1616 values."""
    from b200ens import workloads as W

    N = 3000
    u0, p = W.net16_params(N)
    saveat = np.linspace(0.0, 10.0, 101)
    kw = dict(saveat=saveat, dt=0.01, abstol=1e-8, reltol=1e-8, callback=W.net16_callback())
    full = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(devices=[0], split=True), trajectories=N, **kw)
    assert np.all(full.retcodes == 1)
    fused = _summary(B, B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Vern7(), N, "1", **kw)
    twopass = _summary(B, B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Vern7(), N, "0", **kw)
    assert fused.timing["launches"] < twopass.timing["launches"]
    for s in (fused, twopass):
        assert s.num_monte == N
        assert np.allclose(s.u, full.u_array.mean(axis=0), rtol=1e-11, atol=1e-13)
        assert np.allclose(s.v, full.u_array.var(axis=0, ddof=1), rtol=1e-8, atol=1e-13)


def test_fused_moments_terminating_callback(B, gpu_lib):
    """A trajectory that ends through terminate! counts as successful and fills its remaining save points with the
    terminal state (as the out_u path does)."""
    N = 4000
    rng = np.random.default_rng(5)
    u0 = np.tile(np.array([1.0]), (N, 1))
    p = (0.2 + rng.random((N, 1)))
    prob = B.ODEProblem(lambda u, p, t: [-p[0] * u[0]], np.array([1.0]), (0.0, 10.0), np.array([1.0]))
    cb = B.ContinuousCallback(lambda u, t, integ: u[0] - 0.3, lambda integ: B.terminate_b(integ))
    saveat = np.linspace(0.0, 10.0, 1101)            # 1101 values per trajectory
    kw = dict(saveat=saveat, dt=0.01, abstol=1e-8, reltol=1e-8, callback=cb)
    full = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), trajectories=N, **kw)
    assert np.all(full.retcodes == 2)
    s = _summary(B, B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), N, "1", **kw)
    assert s.num_monte == N
    assert np.allclose(s.u, full.u_array.mean(axis=0), rtol=1e-11, atol=1e-13)
