"""GPU parity for the remaining steppers (BASELINE.json configs 3-5): Vern7, Rosenbrock23,
Rodas4/5/5P, EM, SOSRA, ContinuousCallback -- each against the CPU oracle on identical inputs.

Two oracle flavours are used: the hand-written C models (oracle/models.c, independent of the
sympy->CUDA-C emitter; compared at abstol + reltol*|u|) and the emitter's own source compiled
for the host (same expression tree as the GPU; compared essentially bit-for-bit)."""
import numpy as np
import pytest

from helpers import oracle_fns, within_tol

pytestmark = pytest.mark.gpu


def _ens(B, prob, u0, p):
    return B.EnsembleProblem(prob, u0s=u0, ps=p)


# ---------------------------------------------------------------- Vern7 (config 5's stepper) on Lorenz
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_vern7_lorenz_adaptive(B, gpu_lib, oracle, dtype):
    from b200ens import workloads as W

    N = 3000
    tol = 1e-8 if dtype == np.float64 else 1e-5
    saveat = np.arange(0.0, 10.5, 0.5)
    u0, p = W.lorenz_params(N, "random", seed=7, dtype=dtype)
    sol = B.solve(_ens(B, W.lorenz_problem(dtype), u0, p), B.Vern7(), B.EnsembleB200(), trajectories=N, saveat=saveat,
                  dt=0.05, abstol=tol, reltol=tol)
    ref, rc, st = oracle.solve("lorenz", "Vern7", u0, p, (0.0, 10.0), saveat, 0.05, abstol=tol, reltol=tol, dtype=dtype)
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats[:, :3], st[:, :3])
    ok, worst = within_tol(sol.u_array, ref, tol, tol)
    assert ok, worst


def test_vern7_fixed_dt_f64(B, gpu_lib, oracle):
    from b200ens import workloads as W

    N = 512
    u0, p = W.lorenz_params(N, "ordered")
    sol = B.solve(_ens(B, W.lorenz_problem(), u0, p), B.Vern7(), B.EnsembleB200(), trajectories=N, saveat=[10.0],
                  dt=2e-3, adaptive=False, maxiters=10**6)
    ref, rc, st = oracle.solve("lorenz", "Vern7", u0, p, (0.0, 10.0), [10.0], 2e-3, adaptive=False, maxiters=10**6)
    assert np.all(sol.retcodes == 1)
    rel = np.abs(sol.u_array - ref) / np.maximum(np.abs(ref), 1e-300)
    assert rel.max() <= 1e-12, rel.max()


# ---------------------------------------------------------------- config 3: Robertson, Rosenbrock methods
@pytest.mark.parametrize("alg", ["Rosenbrock23", "Rodas4", "Rodas5", "Rodas5P"])
def test_robertson_rosenbrock(B, gpu_lib, oracle, alg):
    from b200ens import workloads as W

    N = 2000
    abstol, reltol = 1e-8, 1e-6
    u0, p = W.robertson_params(N)
    prob = W.robertson_problem()
    A = getattr(B, alg)()
    sol = B.solve(_ens(B, prob, u0, p), A, B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT, dt=1e-6,
                  abstol=abstol, reltol=reltol)
    assert np.all(sol.retcodes == 1)
    # (a) same expression tree: emitted model compiled for the oracle -> identical step sequences
    model = B.build_model(prob, A)
    ref, rc, st = oracle.solve(None, alg, u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=abstol, reltol=reltol,
                               fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc)
    assert np.array_equal(sol.stats[:, :3], st[:, :3])
    ok, worst = within_tol(sol.u_array, ref, 1e-14, 1e-9)
    assert ok, worst
    # (b) independent hand-written C model: agreement at the solver tolerance
    ref2, rc2, _ = oracle.solve("robertson", alg, u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=abstol, reltol=reltol)
    assert np.array_equal(sol.retcodes, rc2)
    ok, worst = within_tol(sol.u_array, ref2, abstol, reltol)
    assert ok, worst
    # Robertson invariant y1+y2+y3 = 1
    assert np.abs(sol.u_array.sum(axis=2) - 1.0).max() < 1e-9


# ---------------------------------------------------------------- config 4: SDEs with injected increments
def _increments(N, nsteps, nvec, n, dt, dtype, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((N, nsteps, nvec, n)) * np.sqrt(dt)).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_em_gbm_injected_pathwise(B, gpu_lib, oracle, dtype):
    from b200ens import workloads as W

    N, nsteps = 4096, 256
    dt = 1.0 / nsteps
    u0, p = W.gbm_params(N, dtype=dtype)
    dW = _increments(N, nsteps, 1, 1, dt, dtype, 11)
    saveat = np.linspace(0, 1, 5)
    sol = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.EM(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=dt, dW=dW)
    ref, rc, _ = oracle.solve("gbm", "EM", u0, p, (0.0, 1.0), saveat, dt, dtype=dtype, dW=dW, adaptive=False)
    assert np.all(sol.retcodes == 1) and np.all(rc == 1)
    tol = 1e-10 if dtype == np.float64 else 1e-5
    assert np.abs(sol.u_array.astype(np.float64) - ref.astype(np.float64)).max() <= tol * max(1.0, np.abs(ref).max())
    if dtype == np.float64:  # EM converges to the closed form u0 exp((mu - s^2/2) t + s W_t): weak sanity bound
        WT = dW[:, :, 0, 0].sum(axis=1)
        exact = np.exp((p[:, 0] - 0.5 * p[:, 1] ** 2) + p[:, 1] * WT)
        assert np.median(np.abs(sol.u_array[:, -1, 0] - exact) / exact) < 0.05


@pytest.mark.parametrize("alg,nvec", [("EM", 1), ("SOSRA", 2)])
def test_stochastic_lorenz_injected_pathwise(B, gpu_lib, oracle, alg, nvec):
    from b200ens import workloads as W

    N, nsteps = 1024, 512
    dt = 2.0 / nsteps
    u0, p = W.lorenz_additive_params(N)
    dW = _increments(N, nsteps, nvec, 3, dt, np.float64, 13)
    saveat = np.linspace(0, 2, 9)
    prob = W.lorenz_additive_problem(tspan=(0.0, 2.0))
    sol = B.solve(_ens(B, prob, u0, p), getattr(B, alg)(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=dt, dW=dW)
    ref, rc, _ = oracle.solve("lorenz_additive", alg, u0, p, (0.0, 2.0), saveat, dt, dW=dW, adaptive=False)
    assert np.all(sol.retcodes == 1) and np.all(rc == 1)
    assert np.abs(sol.u_array - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_em_philox_device_noise_matches_oracle(B, gpu_lib, oracle, dtype):
    """Device-side Philox4x32-10 + Box-Muller vs the oracle's generator (same counters/keys).  log/sin/cos
    differ by ulps between CUDA and glibc, hence a tolerance instead of bit equality."""
    from b200ens import workloads as W

    N, nsteps = 2048, 64
    dt = 1.0 / nsteps
    u0, p = W.gbm_params(N, dtype=dtype)
    sol = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.EM(), B.EnsembleB200(), trajectories=N, saveat=[1.0], dt=dt, seed=1234)
    ref, rc, _ = oracle.solve("gbm", "EM", u0, p, (0.0, 1.0), [1.0], dt, dtype=dtype, seed=1234, adaptive=False)
    tol = 1e-9 if dtype == np.float64 else 2e-4
    rel = np.abs(sol.u_array.astype(np.float64) - ref.astype(np.float64)) / np.abs(ref.astype(np.float64))
    assert rel.max() < tol, rel.max()
    # the sample mean of GBM is exp(mu t): 2048 paths -> a few percent
    assert abs(np.mean(sol.u_array[:, 0, 0] / np.exp(p[:, 0].astype(np.float64))) - 1.0) < 0.1


# ---------------------------------------------------------------- config 5: Vern7 + ContinuousCallback on the 16-species network
def test_net16_vern7_callback(B, gpu_lib, oracle):
    from b200ens import workloads as W

    N = 1024
    tol = 1e-8
    saveat = np.linspace(0.0, 10.0, 101)
    u0, p = W.net16_params(N)
    prob = W.net16_problem()
    cb = W.net16_callback()
    sol = B.solve(_ens(B, prob, u0, p), B.Vern7(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.01,
                  abstol=tol, reltol=tol, callback=cb)
    assert np.all(sol.retcodes == 1)
    assert sol.stats[:, 3].max() >= 1, "no event fired in the whole ensemble: test is vacuous"
    model = B.build_model(prob, B.Vern7(), cb)
    ref, rc, st = oracle.solve(None, "Vern7", u0, p, (0.0, 10.0), saveat, 0.01, abstol=tol, reltol=tol, event=True,
                               fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc)
    assert np.array_equal(sol.stats, st)                       # same steps, RHS calls and event counts
    ok, worst = within_tol(sol.u_array, ref, 1e-13, 1e-10)
    assert ok, worst
    # independent hand-written network in the oracle: agreement at the solver tolerance (events included)
    ref2, rc2, st2 = oracle.solve("net16", "Vern7", u0, p, (0.0, 10.0), saveat, 0.01, abstol=tol, reltol=tol, event=True)
    assert np.array_equal(st2[:, 3], sol.stats[:, 3])
    ok, worst = within_tol(sol.u_array, ref2, 1e-6, 1e-6)
    assert ok, worst


def test_reference_callback_semantics(B, gpu_lib, oracle):
    """test/core.jl:60-79: condition(u,t,integrator) = t - 0.5 with a no-op affect! -> Success; and terminate!."""
    from b200ens import workloads as W

    prob = B.ODEProblem(W.linear, 0.5, (0.0, 1.0), [1.01])
    cb = B.ContinuousCallback(lambda u, t, integrator: t - 0.5, lambda integrator: None)
    sol = B.solve(prob, B.Tsit5(), callback=cb, dt=0.05, saveat=0.1)
    assert sol.retcode == B.ReturnCode.Success and len(sol.t) == 11
    assert sol.stats["nevents"] == 1
    assert abs(sol.u[-1] - 0.5 * np.exp(1.01)) < 1e-3
    cbt = B.ContinuousCallback(lambda u, t, integrator: u[0] - 0.8, lambda integrator: B.terminate_b(integrator))
    sol2 = B.solve(prob, B.Tsit5(), callback=cbt, dt=0.05, saveat=0.1, abstol=1e-10, reltol=1e-10)
    assert sol2.retcode == B.ReturnCode.Terminated
    assert abs(sol2.u[-1] - 0.8) < 1e-8       # state at the event, held for the remaining save slots


def test_discrete_callback_and_callbackset(B, gpu_lib, oracle):
    """test/core.jl:76-77: DiscreteCallback((u,t,integrator) -> t >= 0.5, affect!).  Here affect! halves u; the
    oracle's hand-written model does the same.  Also a CallbackSet with a ContinuousCallback."""
    from b200ens import workloads as W

    N = 500
    rng = np.random.default_rng(5)
    u0 = 0.5 + rng.random((N, 1))
    p = np.stack([1.01 * (0.5 + rng.random(N)), np.full(N, 0.5)], axis=1)
    prob = B.ODEProblem(W.linear, u0[0], (0.0, 1.0), p[0])
    dcb = B.DiscreteCallback(lambda u, t, integrator: t >= integrator.p[1],
                             lambda integrator: integrator.u.__setitem__(0, integrator.u[0] * 0.5))
    saveat = np.linspace(0, 1, 11)
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.05,
                  callback=dcb, abstol=1e-9, reltol=1e-9)
    ref, rc, st = oracle.solve("linear", "Tsit5", u0, p, (0.0, 1.0), saveat, 0.05, devent=True, abstol=1e-9, reltol=1e-9)
    assert np.all(sol.retcodes == 1) and np.array_equal(sol.retcodes, rc)
    assert np.array_equal(sol.stats, st) and sol.stats[:, 3].min() >= 1
    assert np.abs(sol.u_array - ref).max() <= 1e-14
    # CallbackSet: the continuous crossing of t = 0.5 (no-op affect) plus the discrete halving
    ccb = B.ContinuousCallback(lambda u, t, integrator: t - integrator.p[1], lambda integrator: None)
    sol2 = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.05,
                   callback=B.CallbackSet(ccb, dcb), abstol=1e-9, reltol=1e-9)
    ref2, rc2, st2 = oracle.solve("linear", "Tsit5", u0, p, (0.0, 1.0), saveat, 0.05, event=True, devent=True,
                                  abstol=1e-9, reltol=1e-9)
    assert np.array_equal(sol2.retcodes, rc2) and np.array_equal(sol2.stats, st2)
    assert np.abs(sol2.u_array - ref2).max() <= 1e-14


def test_device_ensemble_summary_matches_host_statistics(B, gpu_lib, oracle):
    """EnsembleAnalysis.timestep_meanvar computed on the device (b200ens_solve_moments) vs numpy statistics of the
    oracle's trajectories; failed trajectories are excluded from the moments and counted out."""
    from b200ens import workloads as W

    N = 30000
    saveat = np.arange(0.0, 10.5, 1.0)
    u0, p = W.lorenz_params(N, "random", seed=17)
    p[5] = np.nan                                               # one trajectory fails (DtNaN)
    eprob = B.EnsembleProblem(W.lorenz_problem(), u0s=u0, ps=p)
    summ = B.solve(eprob, B.Tsit5(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.1, summary=True)
    ref, rc, _ = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), saveat, 0.1)
    ok = rc == 1
    assert summ.num_monte == ok.sum() == N - 1 and np.array_equal(summ.retcodes, rc)
    assert np.allclose(summ.u, ref[ok].mean(axis=0), rtol=1e-12, atol=1e-12)
    assert np.allclose(summ.v, ref[ok].var(axis=0, ddof=1), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("alg", ["Tsit5", "Vern7"])
def test_shared_memory_stage_vectors_bit_identical(B, gpu_lib, alg):
    """B2_KSMEM variant (k-vectors in shared memory as [vector][component][thread]) must reproduce the register
    variant bit for bit, with and without the callback (16-species network, n = 16, f64)."""
    from b200ens import workloads as W

    N = 700
    u0, p = W.net16_params(N)
    prob = W.net16_problem()
    saveat = np.linspace(0.0, 10.0, 21)
    A = getattr(B, alg)()
    for cb in (None, W.net16_callback()):
        kw = dict(trajectories=N, saveat=saveat, dt=0.01, abstol=1e-8, reltol=1e-8, callback=cb)
        a = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), A, B.EnsembleB200(), **kw)
        b = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), A, B.EnsembleB200(stage_vectors_in_smem=True), **kw)
        assert b.timing["smem_bytes"] > 0   # (the default may pick either storage, depending on ptxas spills)
        assert np.array_equal(a.u_array, b.u_array) and np.array_equal(a.stats, b.stats)
        assert np.all(a.retcodes == 1)


def test_reference_dense_solution_interpolation(B, gpu_lib, oracle):
    """test/core.jl:51-58: sol = solve(prob, Tsit5(), dense = true); sol(0.5) isa Number and > 0.5.
    sol(t) must be the value of the run's own dense output: identical to saving at t in the first place."""
    from b200ens import workloads as W

    prob = B.ODEProblem(lambda u, p, t: [1.01 * u[0]], 0.5, (0.0, 1.0))
    sol = B.solve(prob, B.Tsit5(), dense=True)
    assert sol.retcode == B.ReturnCode.Success
    v = sol(0.5)
    assert np.ndim(v) == 0 and v > 0.5
    assert abs(v - 0.5 * np.exp(1.01 * 0.5)) < 1e-5
    tt = np.array([0.9, 0.1, 0.5, 0.0, 1.0])
    ref = B.solve(prob, B.Tsit5(), saveat=np.sort(tt))
    assert np.array_equal(np.sort(sol(tt)), np.asarray(ref.u))     # exponential growth: sorted by time == sorted by value
    with pytest.raises(ValueError):
        B.solve(prob, B.Tsit5())(0.5)
    # ensemble: every trajectory has its own dense output; Vern7 and an event in the step sequence
    N = 64
    u0, p = W.net16_params(N)
    eprob = B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p)
    kw = dict(trajectories=N, dt=0.01, abstol=1e-8, reltol=1e-8, callback=W.net16_callback())
    es = B.solve(eprob, B.Vern7(), B.EnsembleB200(), dense=True, **kw)
    times = np.array([0.37, 3.3, 9.99])
    direct = B.solve(eprob, B.Vern7(), B.EnsembleB200(), saveat=times, **kw)
    for i in (0, 17, 63):
        assert np.array_equal(es[i](times), direct.u_array[i])


def test_vector_continuous_callback(B, gpu_lib, oracle):
    """VectorContinuousCallback(condition!, affect!, len) (qa.jl:124): a ball in a box (gravity and linear drag in y,
    so that the steps stay shorter than an excursion through a wall) -- four walls, the earliest
    crossing fires and affect!(integrator, idx) reflects the matching velocity.  Bit-identical to the oracle (which runs
    the same emitted sources); the elastic x-motion is checked against the closed-form triangle wave."""
    def box(du, u, p, t):
        du[0] = u[1]
        du[1] = 0 * u[0]
        du[2] = u[3]
        du[3] = -p[1] - 0.3 * u[3]

    def condition(out, u, t, integrator):
        out[0] = u[0]
        out[1] = integrator.p[0] - u[0]
        out[2] = u[2]
        out[3] = integrator.p[0] - u[2]

    def affect(integrator, idx):
        if idx <= 2:
            integrator.u[1] = -integrator.u[1]
        else:
            integrator.u[3] = -integrator.u[3]          # elastic: no Zeno accumulation of bounces on the floor

    N = 300
    rng = np.random.default_rng(8)
    u0 = np.stack([0.2 + 0.6 * rng.random(N), 0.5 + 2.0 * rng.random(N), 0.2 + 0.6 * rng.random(N), rng.normal(size=N)], axis=1)
    p = np.stack([np.ones(N), 5.0 + 5.0 * rng.random(N)], axis=1)
    prob = B.ODEProblem(box, u0[0], (0.0, 4.0), p[0])
    cb = B.VectorContinuousCallback(condition, affect, 4, interp_points=20)
    saveat = np.linspace(0.0, 4.0, 41)
    for alg in ("Tsit5", "Vern7", "Rodas5P"):
        A = getattr(B, alg)()
        sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), A, B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.01,
                      abstol=1e-9, reltol=1e-9, callback=cb)
        assert np.all(sol.retcodes == 1)
        model = B.build_model(prob, A, cb)
        ref, rc, st = oracle.solve(None, alg, u0, p, (0.0, 4.0), saveat, 0.01, abstol=1e-9, reltol=1e-9, event=True, ncond=4,
                                   interp_points=20, fns=oracle_fns(oracle, B, model))
        assert np.array_equal(sol.retcodes, rc)
        assert np.array_equal(sol.stats, st)
        ok, worst = within_tol(sol.u_array, ref, 1e-13, 1e-10)
        assert ok, worst
        # elastic walls in x: position = triangle wave of the free flight
        free = u0[:, None, 0] + u0[:, None, 1] * saveat[None, :]
        tri = np.abs(((free + 1.0) % 2.0) - 1.0)
        assert np.max(np.abs(sol.u_array[:, :, 0] - tri)) < 1e-7
        nx = np.floor(u0[:, 0] + u0[:, 1] * 4.0).astype(int)           # wall hits in x up to t = 4
        assert np.all(sol.stats[:, 3] >= nx)                            # plus the bounces in y
        assert np.all((sol.u_array[:, :, 2] > -1e-9) & (sol.u_array[:, :, 2] < 1 + 1e-9))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sriw1_gbm_injected_pathwise_and_strong_accuracy(B, gpu_lib, oracle, dtype):
    """SRIW1 (strong order 1.5, diagonal noise): pathwise agreement with the oracle on injected (dW, dZ) to 1e-10, and
    far closer to the closed form of geometric Brownian motion than Euler-Maruyama on the same Brownian paths."""
    from b200ens import workloads as W

    N, nsteps = 4096, 64
    dt = 1.0 / nsteps
    u0, p = W.gbm_params(N, dtype=dtype)
    dW = _increments(N, nsteps, 2, 1, dt, dtype, 17)
    saveat = np.linspace(0, 1, 5)
    sol = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.SRIW1(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=dt, dW=dW)
    ref, rc, _ = oracle.solve("gbm", "SRIW1", u0, p, (0.0, 1.0), saveat, dt, dtype=dtype, dW=dW, adaptive=False)
    assert np.all(sol.retcodes == 1) and np.all(rc == 1)
    tol = 1e-10 if dtype == np.float64 else 1e-5
    assert np.abs(sol.u_array.astype(np.float64) - ref.astype(np.float64)).max() <= tol * max(1.0, np.abs(ref).max())
    if dtype == np.float64:
        WT = dW[:, :, 0, 0].sum(axis=1)
        exact = np.exp((p[:, 0] - 0.5 * p[:, 1] ** 2) + p[:, 1] * WT)
        em = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.EM(), B.EnsembleB200(), trajectories=N, saveat=[1.0], dt=dt,
                     dW=np.ascontiguousarray(dW[:, :, :1, :]))
        e_sri = np.mean(np.abs(sol.u_array[:, -1, 0] - exact))
        e_em = np.mean(np.abs(em.u_array[:, -1, 0] - exact))
        assert e_sri < e_em / 10, (e_sri, e_em)
    # device Philox noise: same streams as the oracle
    sol2 = B.solve(_ens(B, W.gbm_problem(dtype), u0, p), B.SRIW1(), B.EnsembleB200(), trajectories=N, saveat=[1.0], dt=dt, seed=99)
    ref2, _, _ = oracle.solve("gbm", "SRIW1", u0, p, (0.0, 1.0), [1.0], dt, dtype=dtype, seed=99, adaptive=False)
    rel = np.abs(sol2.u_array.astype(np.float64) - ref2.astype(np.float64)) / np.abs(ref2.astype(np.float64))
    assert rel.max() < (1e-8 if dtype == np.float64 else 5e-4), rel.max()


def test_callbackset_of_continuous_callbacks_and_per_index_terminate(B, gpu_lib, oracle):
    """CallbackSet(cb1, cb2) with two ContinuousCallbacks = one vector callback in the kernel; terminate! in one of the
    affects ends only the trajectories whose FIRST event is that one.  Bit-identical to the oracle."""
    from b200ens import workloads as W

    N = 400
    u0, p = W.lorenz_params(N, "random", seed=12)
    prob = W.lorenz_problem(np.float64, (0.0, 5.0))
    flip = B.ContinuousCallback(lambda u, t, integrator: u[0] + 2.0, lambda integrator: integrator.u.__setitem__(1, -integrator.u[1]))
    stop = B.ContinuousCallback(lambda u, t, integrator: u[2] - 25.0, lambda integrator: B.terminate_b(integrator))
    cs = B.CallbackSet(flip, stop)
    saveat = np.linspace(0.0, 5.0, 11)
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.05,
                  abstol=1e-9, reltol=1e-9, callback=cs)
    model = B.build_model(prob, B.Tsit5(), cs)
    ref, rc, st = oracle.solve(None, "Tsit5", u0, p, (0.0, 5.0), saveat, 0.05, abstol=1e-9, reltol=1e-9, event=True, ncond=2,
                               vterm_mask=0b10, fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.array_equal(sol.stats, st)
    ok, worst = within_tol(sol.u_array, ref, 1e-13, 1e-10)
    assert ok, worst
    term = sol.retcodes == 2
    assert term.any() and (sol.retcodes == 1).any()
    assert np.max(np.abs(sol.u_array[term, -1, 2] - 25.0)) < 1e-6        # held at the terminating event


@pytest.mark.parametrize("alg", ["Rosenbrock23", "Rodas4", "Rodas5P"])
def test_rosenbrock_family_float32(B, gpu_lib, oracle, alg):
    """The Rosenbrock family in Float32 (register LU, derived dense output in single precision): van der Pol with
    mu = 5..50, interpolated saves -- bit-identical to the Float32 oracle, and at the solver tolerance of the Float64 run."""
    def vdp(du, u, p, t):
        du[0] = u[1]
        du[1] = p[0] * ((1 - u[0] ** 2) * u[1] - u[0])

    N = 1024
    rng = np.random.default_rng(21)
    p = (5.0 + 45.0 * rng.random((N, 1))).astype(np.float32)
    u0 = np.tile(np.array([2.0, 0.0], dtype=np.float32), (N, 1))
    saveat = np.linspace(0.0, 10.0, 21).astype(np.float32)
    A = getattr(B, alg)()
    prob = B.ODEProblem(vdp, u0[0], (0.0, 10.0), p[0])
    kw = dict(trajectories=N, saveat=saveat, dt=1e-3, abstol=1e-4, reltol=1e-4)
    sol = B.solve(_ens(B, prob, u0, p), A, B.EnsembleB200(), **kw)
    model = B.build_model(prob, A)
    ref, rc, st = oracle.solve(None, alg, u0, p, (0.0, 10.0), saveat, 1e-3, abstol=1e-4, reltol=1e-4, dtype=np.float32,
                               fns=oracle_fns(oracle, B, model, f64=False))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats[:, :3], st[:, :3]) and np.array_equal(sol.u_array, ref)
    prob64 = B.ODEProblem(vdp, u0[0].astype(np.float64), (0.0, 10.0), p[0].astype(np.float64))
    s64 = B.solve(_ens(B, prob64, u0.astype(np.float64), p.astype(np.float64)), A, B.EnsembleB200(), trajectories=N,
                  saveat=saveat.astype(np.float64), dt=1e-3, abstol=1e-9, reltol=1e-9)
    # the relaxation oscillation amplifies local errors at its fast transitions (a phase error): compare the position away
    # from them in the bulk of the ensemble
    assert np.median(np.abs(sol.u_array[:, :, 0] - s64.u_array[:, :, 0])) < 2e-3
