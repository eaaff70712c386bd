"""solve(...; tstops = [...]) -- times the integrator must hit exactly (OrdinaryDiffEq handle_tstop!) -- and
upstream's dosing idiom built on it: DiscreteCallback(condition = (u, t, integrator) -> t == 4.0, affect!) with
tstops = [4.0].  CPU: the oracle against the closed form; GPU: the kernels bit for bit against the oracle."""
import numpy as np
import pytest

from helpers import oracle_fns


def _dose_problem(B):
    # one-compartment elimination u' = -k u, a dose of p[1] added at t = 4 and t = 8
    prob = B.ODEProblem(lambda u, p, t: [-p[0] * u[0]], np.array([10.0]), (0.0, 12.0), np.array([0.5, 10.0]))
    dcb = B.DiscreteCallback(lambda u, t, integ: (t == 4.0) | (t == 8.0),
                             lambda integ: integ.u.__setitem__(0, integ.u[0] + integ.p[1]))
    return prob, dcb


def _closed_form(u0, k, dose, ts):
    out = np.empty((len(k), len(ts)))
    for j, t in enumerate(ts):
        v = u0 * np.exp(-k * t)
        for td in (4.0, 8.0):
            if t >= td:
                v = v + dose * np.exp(-k * (t - td))
        out[:, j] = v
    return out


def test_oracle_hits_tstops_exactly(oracle):
    """every-step output: the tstops appear among the step times, bit-exact, and only when asked for"""
    from b200ens import workloads as W

    u0, p = W.lorenz_params(8, "random", seed=1)
    cap = 512
    for tst in (None, [2.5, 7.25, 11.0, -1.0]):     # entries outside tspan are ignored
        out, rc, st, times = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), np.zeros(cap), 0.1, save_everystep=1, tstops=tst)
        assert np.all(rc == 1)
        for i in range(8):
            tt = times[i][~np.isnan(times[i])]
            assert (2.5 in tt and 7.25 in tt) == (tst is not None)
            assert tt[-1] == 10.0 and np.all(np.diff(tt) > 0)


def test_oracle_dosing_matches_closed_form(oracle, B):
    prob, dcb = _dose_problem(B)
    model = B.build_model(prob, B.Tsit5(), dcb)
    fns = oracle_fns(oracle, B, model)
    N = 64
    rng = np.random.default_rng(3)
    u0 = np.full((N, 1), 10.0)
    p = np.stack([0.2 + rng.random(N), np.full(N, 10.0)], axis=1)
    saveat = np.array([0.0, 3.0, 4.0, 5.0, 8.0, 9.5, 12.0])
    ref, rc, st = oracle.solve(None, "Tsit5", u0, p, (0.0, 12.0), saveat, 0.1, abstol=1e-10, reltol=1e-10, fns=fns,
                               devent=True, tstops=[4.0, 8.0])
    assert np.all(rc == 1) and np.all(st[:, 3] == 2)          # both doses given, once each
    exact = _closed_form(10.0, p[:, 0], 10.0, saveat)
    # a save point AT the dose time holds the state after the callback (upstream saves after handle_callbacks!)... the
    # saveat point t = 4 is reached exactly at the end of the step, before the affect: the pre-dose value is saved
    pre = exact.copy()
    pre[:, 2] -= 10.0
    pre[:, 4] -= 10.0
    assert np.allclose(ref[:, :, 0], pre, rtol=2e-8, atol=1e-9)
    # without tstops the steps do not land on t = 4 / 8 and the condition t == 4.0 never holds
    ref0, rc0, st0 = oracle.solve(None, "Tsit5", u0, p, (0.0, 12.0), saveat, 0.1, abstol=1e-10, reltol=1e-10, fns=fns, devent=True)
    assert np.all(st0[:, 3] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gpu_dosing_with_tstops_matches_oracle(B, gpu_lib, oracle, dtype):
    prob, dcb = _dose_problem(B)
    prob = B.ODEProblem(prob.f, prob.u0.astype(dtype), prob.tspan, prob.p.astype(dtype))
    N = 3000
    rng = np.random.default_rng(4)
    u0 = np.full((N, 1), 10.0, dtype=dtype)
    p = np.stack([0.2 + rng.random(N), np.full(N, 10.0)], axis=1).astype(dtype)
    saveat = np.linspace(0.0, 12.0, 25)
    tol = 1e-9 if dtype == np.float64 else 1e-5
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), trajectories=N, saveat=saveat, dt=0.1,
                  abstol=tol, reltol=tol, callback=dcb, tstops=[4.0, 8.0])
    model = B.build_model(prob, B.Tsit5(), dcb)
    ref, rc, st = oracle.solve(None, "Tsit5", u0, p, (0.0, 12.0), saveat, 0.1, abstol=tol, reltol=tol, dtype=dtype,
                               fns=oracle_fns(oracle, B, model, dtype == np.float64), devent=True, tstops=[4.0, 8.0])
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st) and np.all(st[:, 3] == 2)
    assert np.array_equal(sol.u_array, ref)


@pytest.mark.gpu
def test_gpu_tstops_lorenz_and_split_kernel(B, gpu_lib, oracle):
    """tstops alone (no callback): same bits as the oracle for the one-thread Tsit5 / Rodas5P kernels and for the split
    Vern7 kernel; the step counts differ from the run without tstops."""
    from b200ens import workloads as W

    N = 2000
    saveat = np.arange(0.0, 10.5, 1.0)
    u0, p = W.lorenz_params(N, "random", seed=9)
    tst = [0.37, 2.5, 7.25]
    for alg in ("Tsit5", "Vern7"):
        sol = B.solve(B.EnsembleProblem(W.lorenz_problem(), u0s=u0, ps=p), getattr(B, alg)(), B.EnsembleB200(devices=[0]), trajectories=N,
                      saveat=saveat, dt=0.1, tstops=tst)
        ref, rc, st = oracle.solve("lorenz", alg, u0, p, (0.0, 10.0), saveat, 0.1, tstops=tst)
        ref0, rc0, st0 = oracle.solve("lorenz", alg, u0, p, (0.0, 10.0), saveat, 0.1)
        assert np.array_equal(sol.retcodes, rc) and np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
        assert not np.array_equal(st, st0)
    u0r, pr = W.robertson_params(500)
    sol = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.Rodas5P(), B.EnsembleB200(devices=[0]), trajectories=500,
                  saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6, tstops=[3.3, 1234.5])
    # bit parity needs the oracle to evaluate the SAME emitted model source (the hand-written C model is another expression tree)
    fns = oracle_fns(oracle, B, B.build_model(W.robertson_problem(), B.Rodas5P()))
    ref, rc, st = oracle.solve(None, "Rodas5P", u0r, pr, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6, fns=fns,
                               tstops=[3.3, 1234.5])
    assert np.array_equal(sol.retcodes, rc) and np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    # split kernel: 16-species network with its ContinuousCallback
    Ns = 400
    u0n, pn = W.net16_params(Ns)
    sv = np.linspace(0.0, 10.0, 21)
    kw = dict(trajectories=Ns, saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8, callback=W.net16_callback(), tstops=[1.111, 6.5])
    a = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0n, ps=pn), B.Vern7(), B.EnsembleB200(devices=[0], split=True), **kw)
    b = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0n, ps=pn), B.Vern7(), B.EnsembleB200(devices=[0], split=False), **kw)
    fns = oracle_fns(oracle, B, B.build_model(W.net16_problem(), B.Vern7(), W.net16_callback()))   # the emitted model: same expression tree
    ref, rc, st = oracle.solve(None, "Vern7", u0n, pn, (0.0, 10.0), sv, 0.01, abstol=1e-8, reltol=1e-8, event=True, fns=fns,
                               tstops=[1.111, 6.5])
    assert np.array_equal(a.u_array, b.u_array) and np.array_equal(a.stats, b.stats)
    assert np.array_equal(a.retcodes, rc) and np.array_equal(a.stats, st)
    assert np.abs(a.u_array - ref).max() <= 1e-13 + 1e-10 * np.abs(ref).max()   # (as test_net16_vern7_callback: libm vs CUDA in the emitted model)


def test_tstops_are_validated(B):
    prob = B.ODEProblem(lambda u, p, t: [-u[0]], np.array([1.0]), (0.0, 1.0), np.array([1.0]))
    sde = B.SDEProblem(lambda u, p, t: [u[0]], lambda u, p, t: [u[0]], np.array([1.0]), (0.0, 1.0), np.array([1.0]))
    with pytest.raises(NotImplementedError):
        B.solve(sde, B.EM(), dt=0.01, tstops=[0.5])
    with pytest.raises(ValueError):
        B.solve(prob, B.Tsit5(), dt=0.1, dtmax=-1.0)
