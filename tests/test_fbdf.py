"""FBDF (/root/reference/test/qa/qa.jl:57): the variable-order fixed-leading-coefficient BDF.
CPU: the oracle's restatement against independent truths (scipy Radau golden vectors, closed forms) and its behaviour as
a variable-order method; JIT of the kernel for sm_100a.  GPU: the kernel bit for bit against the oracle."""
import json
import os

import numpy as np
import pytest

from helpers import oracle_fns

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return json.load(open(os.path.join(GOLD, name)))


def vdp(du, u, p, t):
    du[0] = u[1]
    du[1] = p[0] * ((1 - u[0] ** 2) * u[1] - u[0])


def decay(du, u, p, t):
    # two time scales: u0 relaxes onto cos(t) at rate p0, u1 is a slow oscillator amplitude
    du[0] = -p[0] * (u[0] - u[1])
    du[1] = -p[1] * u[1]


@pytest.mark.parametrize("tol", [1e-4, 1e-6, 1e-8])
def test_fbdf_robertson_against_radau(oracle, tol):
    """Robertson (test/core.jl:39-46) against scipy Radau at 1e-13 (tests/golden/robertson.json): the global error stays
    within 10 x (abstol + reltol |u|) at every save point (measured: 2.7-4.9 x), with the save points as tstops and through
    the cubic Hermite interpolant (upstream's default dense output for multistep methods); the invariant y1 + y2 + y3 = 1
    holds to the conditioning of the Newton solves (the method is linear)."""
    g = _load("robertson.json")
    t, ref = np.array(g["t"]), np.array(g["u"])
    for save_tstops in (True, False):
        out, rc, st = oracle.solve("robertson", "FBDF", [g["u0"]], [g["p"]], (0.0, 1e5), t, 1e-6, abstol=tol * 1e-2, reltol=tol,
                                   save_tstops=save_tstops)
        assert rc[0] == 1
        assert np.max(np.abs(out[0] - ref) / (tol * 1e-2 + tol * np.abs(ref))) < 10.0
        assert np.abs(out[0].sum(axis=1) - 1.0).max() < 1e-9


def test_fbdf_work_is_that_of_a_variable_order_bdf(oracle):
    """Step counts on Robertson at three tolerances against scipy's BDF (variable order 1..5, quasi-constant step):
    measured 136/295/663 accepted steps here against 153/289/569 there.  A fixed low order would need 10-100x more at the
    tight end (BDF2 at 1e-8: > 10^4), so this also pins that the order is raised."""
    g = _load("robertson.json")
    t = np.array(g["t"])
    scipy_bdf_steps = {1e-4: 153, 1e-6: 289, 1e-8: 569}
    for tol, ref_steps in scipy_bdf_steps.items():
        _, rc, st = oracle.solve("robertson", "FBDF", [g["u0"]], [g["p"]], (0.0, 1e5), t, 1e-6, abstol=tol * 1e-2, reltol=tol,
                                 save_tstops=False)
        assert rc[0] == 1
        assert 0.6 * ref_steps < st[0, 0] < 1.5 * ref_steps, (tol, st[0])
        assert st[0, 1] < 0.05 * st[0, 0]          # few rejections
        assert st[0, 2] < 3.5 * st[0, 0]           # ~2 Newton iterations + f(u_new) per step


def test_fbdf_van_der_pol_converges_with_the_tolerance(oracle, B):
    """van der Pol, mu = 100 (golden vector from scipy Radau).  The relaxation oscillation amplifies local errors ~1e3-fold;
    scipy's own BDF measures 4.0 / 0.087 / 1.0e-3 at these tolerances, this one 3.2 / 0.13 / 5.3e-4."""
    from b200ens import codegen

    g = _load("vdp_mu100.json")
    prob = B.ODEProblem(vdp, np.array(g["u0"]), (0.0, 50.0), np.array(g["p"]))
    model = B.build_model(prob, B.FBDF())
    fns = oracle_fns(oracle, B, model)
    ref = np.array(g["u"])
    errs = []
    for tol in (1e-7, 1e-9):
        out, rc, st = oracle.solve(None, "FBDF", [g["u0"]], [g["p"]], (0.0, 50.0), g["t"], 1e-4, abstol=tol, reltol=tol, fns=fns,
                                   maxiters=10**7, save_tstops=False)
        assert rc[0] == 1
        errs.append(np.max(np.abs(out[0] - ref)))
    assert errs[0] < 0.5 and errs[1] < 2e-3 and errs[1] < 0.05 * errs[0]


def test_fbdf_linear_closed_form_and_fixed_step(oracle):
    g = _load("linear.json")
    t, ref = np.array(g["t"]), np.array(g["u"])
    out, rc, st = oracle.solve("linear", "FBDF", [g["u0"]], [g["p"]], (0.0, 1.0), t, 1e-3, abstol=1e-10, reltol=1e-10)
    assert rc[0] == 1 and np.max(np.abs(out[0] - ref)) < 2e-8
    # adaptive = false: the order selection lives in the controller, so a fixed-step run is backward Euler:
    # u_n = u0 / (1 - lambda dt)^n exactly
    dt = 1.0 / 64
    out, rc, st = oracle.solve("linear", "FBDF", [g["u0"]], [g["p"]], (0.0, 1.0), [1.0], dt, adaptive=False, save_tstops=True)
    assert rc[0] == 1 and st[0, 0] == 64
    assert abs(out[0, 0, 0] - g["u0"][0] / (1 - g["p"][0] * dt) ** 64) < 1e-13


def test_fbdf_kernel_compiles(B):
    for dtype in (np.float64, np.float32):
        prob = B.ODEProblem(vdp, np.array([2.0, 0.0], dtype=dtype), (0.0, 1.0), np.array([5.0], dtype=dtype))
        info = B.build_model(prob, B.FBDF()).info()
        assert info["regs"] > 0 and info["cubin_bytes"] > 0 and info["lmem"] == 0   # the divided-difference table stays in registers


def _sawtooth_callback(B):
    # u1 decays like exp(-p1 t); every time it falls through 0.5 it is kicked back up by 0.4
    return B.ContinuousCallback(lambda u, t, integrator: u[1] - 0.5, None, lambda integrator: integrator.u.__setitem__(1, integrator.u[1] + 0.4))


def test_fbdf_with_a_callback_restarts_its_history(oracle, B):
    """A ContinuousCallback on FBDF: the event is located on the Hermite dense output, affect! modifies u and the multistep
    history restarts at order 1 (upstream: u_modified -> reinitFBDF!).  Closed form of the sawtooth: u1 falls from 1 to 0.5
    in ln(2)/p1, then from 0.9 to 0.5 in ln(1.8)/p1 per tooth."""
    N, T = 16, 5.0
    p = np.stack([np.full(N, 200.0), np.linspace(0.3, 2.0, N)], axis=1)
    u0 = np.tile([0.0, 1.0], (N, 1))
    cb = _sawtooth_callback(B)
    prob = B.ODEProblem(decay, u0[0], (0.0, T), p[0])
    model = B.build_model(prob, B.FBDF(), cb)
    out, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, T), [T], 1e-3, abstol=1e-9, reltol=1e-9, event=True, event_dir=-1,
                               save_tstops=False, fns=oracle_fns(oracle, B, model))
    assert np.all(rc == 1)
    t_first = np.log(2.0) / p[:, 1]
    teeth = np.where(T > t_first, 1 + np.floor((T - t_first) / (np.log(1.8) / p[:, 1])), 0).astype(int)
    assert np.array_equal(st[:, 3], teeth)
    t_last = t_first + (teeth - 1) * np.log(1.8) / p[:, 1]
    exact = np.where(teeth > 0, 0.9 * np.exp(-p[:, 1] * (T - t_last)), np.exp(-p[:, 1] * T))
    assert np.max(np.abs(out[:, 0, 1] - exact)) < 2e-6


# ---------------------------------------------------------------- GPU: kernel against the oracle
@pytest.mark.gpu
@pytest.mark.parametrize("save_tstops", [False, True])
def test_gpu_fbdf_robertson_bit_identical(B, gpu_lib, oracle, save_tstops):
    from b200ens import workloads as W

    N = 2000
    u0, p = W.robertson_params(N)
    prob = W.robertson_problem()
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.FBDF(), B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT,
                  dt=1e-6, abstol=1e-8, reltol=1e-6, save_tstops=save_tstops)
    model = B.build_model(prob, B.FBDF())
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6,
                               save_tstops=save_tstops, fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st)                    # accepted / rejected steps (incl. Newton failures), RHS calls
    assert np.array_equal(sol.u_array, ref)
    assert np.abs(sol.u_array.sum(axis=2) - 1.0).max() < 1e-7     # conserved up to the Newton tolerance (kappa x reltol)
    # and it is a stiff solver: the hand-written model agrees at the solver tolerance
    ref2, rc2, _ = oracle.solve("robertson", "Rodas5P", u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-10, reltol=1e-10)
    assert np.max(np.abs(sol.u_array - ref2) / (1e-8 + 1e-6 * np.abs(ref2))) < 20.0     # measured on one set: 2.7-4.9 (CPU test)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gpu_fbdf_two_time_scales_both_precisions(B, gpu_lib, oracle, dtype):
    N = 1500
    rng = np.random.default_rng(4)
    p = np.stack([10.0 ** rng.uniform(1, 4, N), rng.uniform(0.2, 2.0, N)], axis=1).astype(dtype)
    u0 = np.tile(np.array([0.0, 1.0], dtype=dtype), (N, 1))
    prob = B.ODEProblem(decay, u0[0], (0.0, 5.0), p[0])
    saveat = np.linspace(0.0, 5.0, 11).astype(dtype)
    tol = 1e-7 if dtype == np.float64 else 1e-4
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.FBDF(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=1e-4,
                  abstol=tol, reltol=tol)
    model = B.build_model(prob, B.FBDF())
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, 5.0), saveat, 1e-4, abstol=tol, reltol=tol, dtype=dtype, save_tstops=False,
                               fns=oracle_fns(oracle, B, model, f64=(dtype == np.float64)))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    # closed form of the slow component, and the fast one sits on it up to O(1/p0)
    exact1 = np.exp(-p[:, 1:2].astype(np.float64) * saveat[None, :].astype(np.float64))
    assert np.max(np.abs(sol.u_array[:, :, 1] - exact1)) < (2e-5 if dtype == np.float64 else 5e-3)


@pytest.mark.gpu
def test_gpu_fbdf_fixed_step_is_backward_euler(B, gpu_lib, oracle):
    N = 256
    rng = np.random.default_rng(9)
    p = np.stack([10.0 ** rng.uniform(1, 3, N), rng.uniform(0.2, 2.0, N)], axis=1)
    u0 = np.tile(np.array([0.0, 1.0]), (N, 1))
    prob = B.ODEProblem(decay, u0[0], (0.0, 1.0), p[0])
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.FBDF(), B.EnsembleB200(), trajectories=N, saveat=[0.5, 1.0], dt=1 / 128,
                  adaptive=False)
    model = B.build_model(prob, B.FBDF())
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, 1.0), [0.5, 1.0], 1 / 128, adaptive=False, save_tstops=False,
                               fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1) and np.all(st[:, 0] == 128)
    assert np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    assert np.max(np.abs(sol.u_array[:, 1, 1] - (1 + p[:, 1] / 128) ** -128.0)) < 1e-12


@pytest.mark.gpu
def test_gpu_fbdf_with_callbacks_matches_oracle(B, gpu_lib, oracle):
    """ContinuousCallback (downcrossings only) and a DiscreteCallback on FBDF: events on the Hermite dense output, history
    restart after every affect! -- bit for bit against the oracle."""
    N, T = 1024, 5.0
    rng = np.random.default_rng(12)
    p = np.stack([10.0 ** rng.uniform(1, 3, N), rng.uniform(0.3, 2.0, N)], axis=1)
    u0 = np.tile([0.0, 1.0], (N, 1))
    saveat = np.linspace(0.0, T, 11)
    prob = B.ODEProblem(decay, u0[0], (0.0, T), p[0])
    cb = _sawtooth_callback(B)
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.FBDF(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=1e-3,
                  abstol=1e-8, reltol=1e-8, callback=cb)
    model = B.build_model(prob, B.FBDF(), cb)
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, T), saveat, 1e-3, abstol=1e-8, reltol=1e-8, event=True, event_dir=-1,
                               save_tstops=False, fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    assert st[:, 3].min() >= 1
    dcb = B.DiscreteCallback(lambda u, t, integrator: u[1] < 0.3, lambda integrator: integrator.u.__setitem__(1, integrator.u[1] + 0.5))
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.FBDF(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=1e-3,
                  abstol=1e-8, reltol=1e-8, callback=dcb)
    model = B.build_model(prob, B.FBDF(), dcb)
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, T), saveat, 1e-3, abstol=1e-8, reltol=1e-8, devent=True,
                               save_tstops=False, fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    assert st[:, 3].max() >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("alg", ["Rodas5P", "FBDF"])
def test_gpu_stiff_steppers_on_the_16_species_network(B, gpu_lib, oracle, alg):
    """n = 16: beyond 10 states the LU of the stiff steppers runs with rolled loops on a local-memory matrix (b2_rosenbrock.cuh,
    B2_LU_ROLLED) instead of fully unrolled register code -- the same operations in the same order, so still bit-identical
    to the oracle."""
    from b200ens import workloads as W

    N = 256
    u0, p = W.net16_params(N)
    prob = W.net16_problem()
    saveat = np.linspace(0.0, 10.0, 11)
    A = getattr(B, alg)()
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), A, B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.01, abstol=1e-8,
                  reltol=1e-8)
    model = B.build_model(prob, A)
    assert model.info()["cubin_bytes"] < 2_500_000           # rolled LU: ~1 MB instead of 4 MB of unrolled SASS
    ref, rc, st = oracle.solve(None, alg, u0, p, (0.0, 10.0), saveat, 0.01, abstol=1e-8, reltol=1e-8, fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    # against the explicit high-order solve of the same (non-stiff at these rates) network
    v7 = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.01, abstol=1e-11,
                 reltol=1e-11)
    assert np.max(np.abs(sol.u_array - v7.u_array)) < (2e-6 if alg == "FBDF" else 2e-7)


@pytest.mark.gpu
def test_gpu_fbdf_with_solve_options(B, gpu_lib, oracle):
    """FBDF through the generic kernel entry: user tstops (bit-identical to the oracle), save_idxs (the selected columns of
    the full run) and the on-device ensemble summary (mean of the full output)."""
    from b200ens import workloads as W

    N = 1500
    u0, p = W.robertson_params(N)
    prob = W.robertson_problem((0.0, 1e3))
    sv = np.array([1e-3, 1.0, 40.0, 1e3])
    kw = dict(trajectories=N, saveat=sv, dt=1e-6, abstol=1e-8, reltol=1e-6)
    eprob = B.EnsembleProblem(prob, u0s=u0, ps=p)
    full = B.solve(eprob, B.FBDF(), B.EnsembleB200(), **kw)
    ts = B.solve(eprob, B.FBDF(), B.EnsembleB200(), tstops=[0.5, 40.0, 333.0], **kw)
    model = B.build_model(prob, B.FBDF())
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, 1e3), sv, 1e-6, abstol=1e-8, reltol=1e-6, save_tstops=False,
                               tstops=[0.5, 40.0, 333.0], fns=oracle_fns(oracle, B, model))
    assert np.array_equal(ts.retcodes, rc) and np.all(rc == 1) and np.array_equal(ts.stats, st) and np.array_equal(ts.u_array, ref)
    assert not np.array_equal(ts.stats, full.stats)                       # the stops change the step sequence
    cols = B.solve(eprob, B.FBDF(), B.EnsembleB200(), save_idxs=[2, 0], **kw)
    assert np.array_equal(cols.u_array, full.u_array[:, :, [2, 0]])
    summ = B.solve(eprob, B.FBDF(), B.EnsembleB200(), summary=True, **kw)
    assert summ.num_monte == N
    assert np.max(np.abs(np.asarray(summ.u) - full.u_array.mean(axis=0))) < 1e-12


def rober_dae(du, u, p, t):
    # Robertson as an index-1 DAE: M = diag(1, 1, 0), third equation = the conservation law
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2]
    du[1] = p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2]
    du[2] = u[0] + u[1] + u[2] - 1.0


def test_fbdf_mass_matrix_dae_against_radau(oracle, B):
    """FBDF on M u' = f with a singular constant mass matrix: M (z + tmp) = beta dt f(z), W = M - beta dt J.  The DAE form of
    Robertson gives the accuracy of the ODE form and keeps the algebraic equation to rounding."""
    g = _load("robertson.json")
    prob = B.ODEProblem(rober_dae, np.array(g["u0"]), (0.0, 1e5), np.array(g["p"]), mass_matrix=np.diag([1.0, 1.0, 0.0]))
    model = B.build_model(prob, B.FBDF())
    fns = oracle_fns(oracle, B, model)
    ref = np.array(g["u"])
    for tol in (1e-4, 1e-6, 1e-8):
        out, rc, st = oracle.solve(None, "FBDF", [g["u0"]], [g["p"]], (0.0, 1e5), g["t"], 1e-6, abstol=tol * 1e-2, reltol=tol, fns=fns,
                                   mass_matrix=np.diag([1.0, 1.0, 0.0]))
        assert rc[0] == 1
        assert np.max(np.abs(out[0] - ref) / (tol * 1e-2 + tol * np.abs(ref))) < 10.0
        assert np.abs(out[0].sum(axis=1) - 1.0).max() < 1e-14


@pytest.mark.gpu
def test_gpu_fbdf_mass_matrix_dae_bit_identical(B, gpu_lib, oracle):
    from b200ens import workloads as W

    N = 1024
    u0, p = W.robertson_params(N)
    M = np.diag([1.0, 1.0, 0.0])
    prob = B.ODEProblem(rober_dae, u0[0], (0.0, 1e5), p[0], mass_matrix=M)
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.FBDF(), B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT, dt=1e-6,
                  abstol=1e-8, reltol=1e-6)
    model = B.build_model(prob, B.FBDF())
    ref, rc, st = oracle.solve(None, "FBDF", u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6, mass_matrix=M,
                               fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    assert np.abs(sol.u_array.sum(axis=2) - 1.0).max() < 1e-13
