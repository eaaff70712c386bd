"""The reference's own functional tests (/root/reference/test/core.jl), transcribed testset by testset onto the host
mirror of its interface and run through EnsembleB200 on the GPU.  These are the only behaviours the reference pins
(SURVEY.md section 4); trajectory values are checked against closed forms where one exists."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_basic_ode_solve(B, gpu_lib):                     # core.jl:9-19
    f = lambda u, p, t: [1.01 * u[0]]
    prob = B.ODEProblem(f, 0.5, (0.0, 1.0))
    sol = B.solve(prob, B.Tsit5(), reltol=1.0e-8, abstol=1.0e-8)
    assert sol.retcode == B.ReturnCode.Success
    assert len(sol.t) > 0 and len(sol.u) > 0
    assert sol.u[0] == 0.5
    assert abs(sol.u[-1] - 0.5 * np.exp(1.01)) < 1e-7


def test_in_place_ode_solve(B, gpu_lib):                  # core.jl:21-37
    def lorenz(du, u, p, t):
        sigma, rho, beta = p
        du[0] = sigma * (u[1] - u[0])
        du[1] = u[0] * (rho - u[2]) - u[1]
        du[2] = u[0] * u[1] - beta * u[2]

    u0 = [1.0, 0.0, 0.0]
    prob = B.ODEProblem(lorenz, u0, (0.0, 1.0), (10.0, 28.0, 8 / 3))
    sol = B.solve(prob, B.Tsit5())
    assert sol.retcode == B.ReturnCode.Success
    assert np.array_equal(sol.u[0], u0)
    assert len(sol.u[-1]) == 3


def test_stiff_ode_solve(B, gpu_lib):                     # core.jl:39-49 (the precompile workload, src/DifferentialEquations.jl:14-28)
    def rober(du, u, p, t):
        y1, y2, y3 = u
        k1, k2, k3 = p
        du[0] = -k1 * y1 + k3 * y2 * y3
        du[1] = k1 * y1 - k2 * y2 ** 2 - k3 * y2 * y3
        du[2] = k2 * y2 ** 2

    prob = B.ODEProblem(rober, [1.0, 0.0, 0.0], (0.0, 1.0e5), (0.04, 3.0e7, 1.0e4))
    sol = B.solve(prob, B.Rodas5P())
    assert sol.retcode == B.ReturnCode.Success
    assert abs(np.sum(sol.u[-1]) - 1.0) < 1e-6            # Robertson invariant y1 + y2 + y3 = 1
    assert abs(sol.u[-1][0] - 0.0178) < 5e-4              # y1(1e5) of the classic problem (1.786e-2)


def test_stiff_ode_solve_fbdf(B, gpu_lib):               # qa.jl:57 exports FBDF next to Rodas5P; same problem, same call shape
    from b200ens import workloads as W

    prob = B.ODEProblem(W.robertson, [1.0, 0.0, 0.0], (0.0, 1.0e5), (0.04, 3.0e7, 1.0e4))
    sol = B.solve(prob, B.FBDF())                        # default tolerances, automatic initial dt, every step saved
    assert sol.retcode == B.ReturnCode.Success
    assert abs(sum(sol.u[-1]) - 1.0) < 1e-6 and sol.t[-1] == 1.0e5
    assert 0.015 < sol.u[-1][0] < 0.02                    # y1(1e5) = 0.017866 (scipy Radau, tests/golden/robertson.json)


def test_solution_interpolation(B, gpu_lib):              # core.jl:51-58
    prob = B.ODEProblem(lambda u, p, t: [1.01 * u[0]], 0.5, (0.0, 1.0))
    sol = B.solve(prob, B.Tsit5(), dense=True)
    u_interp = sol(0.5)
    assert np.ndim(u_interp) == 0
    assert u_interp > 0.5


def test_callbacks(B, gpu_lib):                           # core.jl:60-79
    def lorenz(du, u, p, t):
        du[0] = 10.0 * (u[1] - u[0])
        du[1] = u[0] * (28.0 - u[2]) - u[1]
        du[2] = u[0] * u[1] - (8 / 3) * u[2]

    prob = B.ODEProblem(lorenz, [1.0, 0.0, 0.0], (0.0, 1.0))
    condition = lambda u, t, integrator: t - 0.5
    affect = lambda integrator: None
    cb = B.ContinuousCallback(condition, affect)
    sol = B.solve(prob, B.Tsit5(), callback=cb)
    assert sol.retcode == B.ReturnCode.Success
    dcb = B.DiscreteCallback(lambda u, t, integrator: t >= 0.5, affect)
    sol2 = B.solve(prob, B.Tsit5(), callback=dcb)
    assert sol2.retcode == B.ReturnCode.Success
    assert np.array_equal(sol.u[-1], sol2.u[-1]) or np.allclose(sol.u[-1], sol2.u[-1], rtol=1e-2)   # no-op affects


def test_remake(B, gpu_lib):                              # core.jl:81-88
    prob = B.ODEProblem(lambda u, p, t: [p[0] * u[0]], 0.5, (0.0, 1.0), 1.01)
    prob2 = B.remake(prob, u0=1.0)
    sol = B.solve(prob2, B.Tsit5())
    assert sol.retcode == B.ReturnCode.Success
    assert abs(sol.u[0] - 1.0) < 1e-15


def test_saveat(B, gpu_lib):                              # core.jl:90-96
    prob = B.ODEProblem(lambda u, p, t: [1.01 * u[0]], 0.5, (0.0, 1.0))
    sol = B.solve(prob, B.Tsit5(), saveat=0.1)
    assert sol.retcode == B.ReturnCode.Success
    assert len(sol.t) == 11


def test_ensemble_batches_output_func_reduction(B, gpu_lib):
    """EnsembleProblem(prob; prob_func, output_func, reduction, u_init) + solve(...; trajectories, batch_size)
    (qa.jl:50,56): batches, global 1-based indices, rerun with repeat+1, early exit on convergence."""
    from b200ens import workloads as W

    base = W.lorenz_problem(np.float64, (0.0, 1.0))
    calls = []

    def prob_func(prob, i, repeat):
        calls.append((i, repeat))
        return B.remake(prob, p=[10.0, 20.0 + i * 0.01 + (5.0 if repeat > 1 else 0.0), 8 / 3])

    def output_func(sol, i):
        rerun = (i == 3) and (i, 2) not in calls          # ask once for a rerun of trajectory 3
        return (float(sol.u[-1][0]), i), rerun

    seen = []

    def reduction(u, data, I):
        seen.append((list(I)[0], list(I)[-1], len(data)))
        u = u + [d[0] for d in data]
        return u, len(u) >= 200                            # converged after two batches of 100

    eprob = B.EnsembleProblem(base, prob_func=prob_func, output_func=output_func, reduction=reduction, u_init=[])
    sol = B.solve(eprob, B.Tsit5(), B.EnsembleB200(), trajectories=250, batch_size=100, saveat=[1.0], dt=0.01)
    assert sol.converged and sol.batches == 2 and len(sol.u) == 200
    assert seen == [(1, 100, 100), (101, 200, 100)]
    assert (3, 2) in calls and max(i for i, _ in calls) == 200          # batch 3 never ran, trajectory 3 was rerun
    # values: same as one plain solve of the same parameters (trajectory 3 with its rerun parameters)
    ps = np.array([[10.0, 20.0 + i * 0.01 + (5.0 if i == 3 else 0.0), 8 / 3] for i in range(1, 201)])
    ref = B.solve(B.EnsembleProblem(base, ps=ps), B.Tsit5(), B.EnsembleB200(), trajectories=200, saveat=[1.0], dt=0.01)
    assert np.array_equal(np.array(sol.u), ref.u_array[:, -1, 0])
    # default output_func / reduction with batch_size: a list of per-trajectory solutions
    sol2 = B.solve(B.EnsembleProblem(base, ps=ps), B.Tsit5(), B.EnsembleB200(), trajectories=200, batch_size=64, saveat=[1.0], dt=0.01)
    assert len(sol2.u) == 200 and not sol2.converged and sol2.batches == 4
    assert np.array_equal(np.array([s.u[-1] for s in sol2.u]), ref.u_array[:, -1, :])


def test_sde_batches_share_the_global_noise_streams(B, gpu_lib):
    """Philox counters are keyed by the GLOBAL trajectory index: solving in batches gives the same paths."""
    from b200ens import workloads as W

    N = 300
    u0, p = W.gbm_params(N)
    eprob = B.EnsembleProblem(W.gbm_problem(), u0s=u0, ps=p)
    one = B.solve(eprob, B.EM(), B.EnsembleB200(), trajectories=N, saveat=[1.0], dt=1 / 64, seed=11)
    parts = B.solve(eprob, B.EM(), B.EnsembleB200(), trajectories=N, batch_size=128, saveat=[1.0], dt=1 / 64, seed=11)
    assert np.array_equal(np.array([s.u[-1] for s in parts.u]).reshape(N), one.u_array[:, -1, 0])


def test_ensemble_analysis_on_a_device_run(B, gpu_lib):
    """EnsembleAnalysis over a real ensemble; timestep_meanvar agrees with the on-device summary
    (b200ens_solve_moments); timepoint_* off the save grid go through the trajectories' dense output."""
    from b200ens import EnsembleAnalysis as EA
    from b200ens import workloads as W

    N = 2000
    u0, p = W.lorenz_params(N, "random", seed=2)
    eprob = B.EnsembleProblem(W.lorenz_problem(np.float64, (0.0, 2.0)), u0s=u0, ps=p)
    kw = dict(trajectories=N, saveat=0.5, dt=0.05)
    sim = B.solve(eprob, B.Tsit5(), B.EnsembleB200(), dense=True, **kw)
    summ = B.solve(eprob, B.Tsit5(), B.EnsembleB200(), summary=True, **kw)
    m, v = EA.timeseries_steps_meanvar(sim)
    assert np.allclose(m, summ.u, rtol=1e-12, atol=1e-12) and np.allclose(v, summ.v, rtol=1e-9, atol=1e-12)
    direct = B.solve(eprob, B.Tsit5(), B.EnsembleB200(), trajectories=N, saveat=[0.73], dt=0.05)
    sub = B.EnsembleSolution(sim.t, sim.u_array[:50], sim.retcodes[:50], None, 0.0, {}, dense=sim._dense)
    assert np.array_equal(EA.timepoint_mean(sub, 0.73), direct.u_array[:50, 0].mean(axis=0))


def test_per_component_tolerances(B, gpu_lib, oracle):
    """solve(prob, Rodas5P(); reltol = 1e-8, abstol = [1e-8, 1e-14, 1e-6]) -- the standard Robertson call of the
    DifferentialEquations.jl documentation: per-component tolerances, bit-identical step sequence to the oracle."""
    from b200ens import workloads as W
    from helpers import oracle_fns

    N = 200
    u0, p = W.robertson_params(N)
    atol = np.array([1e-8, 1e-14, 1e-6])
    eprob = B.EnsembleProblem(W.robertson_problem(), u0s=u0, ps=p)
    for alg in ("Rodas5P", "Rosenbrock23"):
        sol = B.solve(eprob, getattr(B, alg)(), B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT, dt=1e-6,
                      reltol=1e-8, abstol=atol)
        # the oracle runs the SAME emitted model source (one expression tree on both sides -> identical step sequences)
        model = B.build_model(W.robertson_problem(), getattr(B, alg)())
        ref, rc, st = oracle.solve(None, alg, u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=atol, reltol=1e-8,
                                   fns=oracle_fns(oracle, B, model))
        assert np.all(sol.retcodes == 1) and np.array_equal(sol.retcodes, rc)
        assert np.array_equal(sol.stats[:, :3], st[:, :3])
        assert np.allclose(sol.u_array, ref, rtol=1e-12, atol=1e-300)
        scal = B.solve(eprob, getattr(B, alg)(), B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT, dt=1e-6,
                       reltol=1e-8, abstol=1e-8)
        assert not np.array_equal(scal.stats, sol.stats)            # the tight tolerance on y2 costs steps
    # a vector of equal entries is the scalar case, bit for bit; Tsit5 Float32 and the split kernel take vectors too
    u0l, pl = W.lorenz_params(1000, "random", seed=4, dtype=np.float32)
    el = B.EnsembleProblem(W.lorenz_problem(np.float32), u0s=u0l, ps=pl)
    kw = dict(trajectories=1000, saveat=1.0, dt=0.1)
    a = B.solve(el, B.Tsit5(), B.EnsembleB200(), abstol=1e-6, reltol=1e-3, **kw)
    b = B.solve(el, B.Tsit5(), B.EnsembleB200(), abstol=[1e-6] * 3, reltol=[1e-3] * 3, **kw)
    assert np.array_equal(a.u_array, b.u_array) and np.array_equal(a.stats, b.stats)
    u0n, pn = W.net16_params(100)
    en = B.EnsembleProblem(W.net16_problem(), u0s=u0n, ps=pn)
    av = np.geomspace(1e-10, 1e-6, 16)
    outs = [B.solve(en, B.Vern7(), B.EnsembleB200(split=sp), trajectories=100, saveat=1.0, dt=0.01, abstol=av, reltol=1e-7,
                    callback=W.net16_callback()) for sp in (False, True)]
    assert np.array_equal(outs[0].u_array, outs[1].u_array) and np.array_equal(outs[0].stats, outs[1].stats)
    with pytest.raises(ValueError):
        B.solve(el, B.Tsit5(), B.EnsembleB200(), abstol=[1e-6, 1e-6], **kw)


def test_mass_matrix_dae_rodas(B, gpu_lib, oracle):
    """ODEFunction(f; mass_matrix = M) with a singular M: the Robertson problem as an index-1 DAE (third equation
    y1 + y2 + y3 = 1), the standard Rodas5P DAE example of the DifferentialEquations.jl documentation."""
    from b200ens import workloads as W
    from helpers import oracle_fns

    def rober_dae(du, u, p, t):
        du[0] = -p[0] * u[0] + p[2] * u[1] * u[2]
        du[1] = p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2]
        du[2] = u[0] + u[1] + u[2] - 1.0

    M = np.diag([1.0, 1.0, 0.0])
    prob = B.ODEProblem(rober_dae, [1.0, 0.0, 0.0], (0.0, 1e5), (0.04, 3e7, 1e4), mass_matrix=M)
    sol = B.solve(prob, B.Rodas5P(), reltol=1e-8, abstol=1e-8)            # the documentation's call
    assert sol.retcode == B.ReturnCode.Success
    assert abs(np.sum(sol.u[-1]) - 1.0) < 1e-12
    assert abs(sol.u[-1][0] - 0.017865) < 1e-5
    N = 400
    u0, p = W.robertson_params(N)
    eprob = B.EnsembleProblem(prob, u0s=u0, ps=p)
    ode = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0, ps=p), B.Rodas5P(), B.EnsembleB200(), trajectories=N,
                  saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-10, reltol=1e-8)
    for alg in ("Rodas5P", "Rodas5", "Rodas4"):
        A = getattr(B, alg)()
        dae = B.solve(eprob, A, B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-10, reltol=1e-8)
        assert np.all(dae.retcodes == 1)
        assert np.max(np.abs(dae.u_array.sum(axis=2) - 1.0)) < 1e-13          # the algebraic constraint holds exactly
        # DAE and ODE form take different steps and interpolate: they agree at the solver tolerance
        assert np.max(np.abs(dae.u_array - ode.u_array) / (1e-10 + 1e-8 * np.abs(ode.u_array))) < 10.0
        ref, rc, st = oracle.solve(None, alg, u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-10, reltol=1e-8,
                                   fns=oracle_fns(oracle, B, B.build_model(prob, A)), mass_matrix=M)
        assert np.array_equal(dae.retcodes, rc) and np.array_equal(dae.stats[:, :3], st[:, :3])
        assert np.allclose(dae.u_array, ref, rtol=1e-12, atol=1e-300)
    with pytest.raises(NotImplementedError):
        B.solve(prob, B.Tsit5())
    # dense output on a DAE: the Rodas dense output works on the stage increments (no u' needed); sol(t) keeps the
    # linear algebraic equation and agrees with saving at t in the first place
    dsol = B.solve(prob, B.Rodas5P(), reltol=1e-8, abstol=1e-8, dense=True)
    v = dsol(np.array([0.5, 40.0, 7e3]))
    assert np.max(np.abs(np.sum(v, axis=-1) - 1.0)) < 1e-13
    direct = B.solve(prob, B.Rodas5P(), reltol=1e-8, abstol=1e-8, saveat=[0.5, 40.0, 7e3])
    assert np.array_equal(np.asarray(v), np.asarray(direct.u))
    with pytest.raises(NotImplementedError):
        B.solve(prob, B.FBDF(), dense=True)                # FBDF's Hermite output takes f for u'


def test_save_everystep_default_for_single_solves(B, gpu_lib, oracle):
    """Without saveat upstream saves every accepted step (save_everystep = true): sol.t / sol.u of a single solve hold
    the whole step sequence (test/core.jl:14-18 only pins length > 0 and sol.u[1] == u0).  Bit-identical to the oracle's
    step sequence; the capacity grows until the longest trajectory fits; ensembles opt in."""
    from b200ens import workloads as W
    from helpers import oracle_fns

    prob = W.lorenz_problem(np.float64, (0.0, 10.0))
    sol = B.solve(prob, B.Tsit5(), dt=0.1)                       # no saveat -> every step
    n = sol.stats["naccept"] + 1
    assert len(sol.t) == n and len(sol.u) == n and n > 50
    assert sol.t[0] == 0.0 and sol.t[-1] == 10.0 and np.all(np.diff(sol.t) > 0)
    assert np.array_equal(sol.u[0], prob.u0)
    model = B.build_model(prob, B.Tsit5(), split=False)
    out, rc, st, tt = oracle.solve(None, "Tsit5", prob.u0[None, :], prob.p[None, :], (0.0, 10.0), np.zeros(256), 0.1,
                                   fns=oracle_fns(oracle, B, model), save_everystep=1)
    assert st[0, 0] + 1 == n
    assert np.array_equal(sol.t, tt[0, :n]) and np.array_equal(sol.u, out[0, :n])
    # with saveat the grid wins; the end state is the same either way
    grid = B.solve(prob, B.Tsit5(), dt=0.1, saveat=[0.0, 10.0])
    assert np.array_equal(grid.u[-1], sol.u[-1])
    # ensemble, opt-in: ragged lengths, capacity doubling (tight tolerance -> > 256 steps), callbacks included
    N = 100
    u0, p = W.lorenz_params(N, "random", seed=6)
    es = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(), trajectories=N, dt=0.01, abstol=1e-12,
                 reltol=1e-12, save_everystep=True)
    assert es.u_array.shape[1] >= int(es.stats[:, 0].max()) + 1 > 256
    for i in (0, 50, 99):
        s = es[i]
        assert len(s.t) == es.stats[i, 0] + 1 and s.t[-1] == 10.0 and not np.isnan(s.u).any()
    ref = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(), trajectories=N, dt=0.01, abstol=1e-12,
                  reltol=1e-12, saveat=[10.0])
    assert np.array_equal(np.stack([es[i].u[-1] for i in range(N)]), ref.u_array[:, 0])
    cbt = B.ContinuousCallback(lambda u, t, integrator: u[2] - 20.0, lambda integrator: B.terminate_b(integrator))
    term = B.solve(prob, B.Tsit5(), dt=0.1, callback=cbt)
    assert term.retcode == B.ReturnCode.Terminated and abs(term.u[-1][2] - 20.0) < 1e-6 and term.t[-1] < 10.0
    rob = B.solve(W.robertson_problem(), B.Rodas5P())            # the stiff example keeps its step list too
    assert rob.retcode == B.ReturnCode.Success and len(rob.t) == rob.stats["naccept"] + 1 and rob.t[-1] == 1e5
