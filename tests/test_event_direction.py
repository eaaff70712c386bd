"""ContinuousCallback direction (/root/reference/test/core.jl:69-72 pins the constructor): an upcrossing of
the condition runs affect!, a downcrossing affect_neg! (default: the same function); `None` (Julia: nothing) for either
one ignores that direction.  Oracle semantics on the CPU against the closed form of a harmonic oscillator, then the
one-thread and the split kernel bit for bit against the oracle on the GPU."""
import numpy as np
import pytest

from helpers import oracle_fns


def osc(du, u, p, t):
    # x'' = -w^2 x; u[2] is an event counter the affects write to
    du[0] = u[1]
    du[1] = -p[0] * p[0] * u[0]
    du[2] = 0 * u[0]


def _cond(u, t, integrator):
    return u[0] - integrator.p[1]


def _up(integrator):
    integrator.u[2] = integrator.u[2] + 1


def _down(integrator):
    integrator.u[2] = integrator.u[2] + 10


def _down_term(integrator):
    import b200ens as B

    integrator.u[2] = integrator.u[2] + 10
    B.terminate_b(integrator)


def _cases(B):
    return {
        "both_same": (B.ContinuousCallback(_cond, _up), dict()),
        "up_only": (B.ContinuousCallback(_cond, _up, None), dict(event_dir=1)),
        "down_only": (B.ContinuousCallback(_cond, None, _down), dict(event_dir=-1)),
        "distinct": (B.ContinuousCallback(_cond, _up, _down), dict()),
        "distinct_terminate_on_down": (B.ContinuousCallback(_cond, _up, _down_term), dict(terminate_neg=True)),
    }


def _params(N, seed=11):
    rng = np.random.default_rng(seed)
    w = 1.0 + rng.random(N)                  # angular frequency
    thr = 0.5 * rng.random(N) - 0.25         # threshold of the condition x - thr
    u0 = np.zeros((N, 3))
    u0[:, 1] = w                             # x = sin(w t)
    return u0, np.stack([w, thr], axis=1)


def _expected_counts(p, T):
    """Crossings of sin(w t) = thr in (0, T]: up at (asin(thr) + 2 pi k)/w, down at (pi - asin(thr) + 2 pi k)/w."""
    w, thr = p[:, 0], p[:, 1]
    a = np.arcsin(thr)
    ups = np.zeros(len(w), dtype=int)
    downs = np.zeros(len(w), dtype=int)
    for k in range(-1, 8):
        tu = (a + 2 * np.pi * k) / w
        td = (np.pi - a + 2 * np.pi * k) / w
        ups += (tu > 1e-6) & (tu <= T)
        downs += (td > 1e-6) & (td <= T)
    return ups, downs


def _oracle_run(oracle, B, cb, okw, u0, p, T, saveat):
    prob = B.ODEProblem(osc, u0[0], (0.0, T), p[0])
    model = B.build_model(prob, B.Tsit5(), cb)
    return oracle.solve(None, "Tsit5", u0, p, (0.0, T), saveat, 0.01, abstol=1e-9, reltol=1e-9, event=True,
                        terminate=False, fns=oracle_fns(oracle, B, model), **okw)


@pytest.mark.parametrize("case", ["both_same", "up_only", "down_only", "distinct", "distinct_terminate_on_down"])
def test_oracle_event_direction_counts(oracle, B, case):
    N, T = 48, 9.0
    u0, p = _params(N)
    # thresholds below the start value make the first crossing an upcrossing only when thr > 0; keep the start off the threshold
    cb, okw = _cases(B)[case]
    saveat = np.array([T])
    out, rc, st = _oracle_run(oracle, B, cb, okw, u0, p, T, saveat)
    ups, downs = _expected_counts(p, T)
    # x(0) = 0: for thr > 0 the condition starts negative, for thr < 0 positive; either way crossings alternate
    if case == "both_same":
        assert np.all(rc == 1) and np.array_equal(st[:, 3], ups + downs) and np.allclose(out[:, 0, 2], ups + downs)
    elif case == "up_only":
        assert np.all(rc == 1) and np.array_equal(st[:, 3], ups) and np.allclose(out[:, 0, 2], ups)
    elif case == "down_only":
        assert np.all(rc == 1) and np.array_equal(st[:, 3], downs) and np.allclose(out[:, 0, 2], 10 * downs)
    elif case == "distinct":
        assert np.all(rc == 1) and np.array_equal(st[:, 3], ups + downs) and np.allclose(out[:, 0, 2], ups + 10 * downs)
    else:
        # the first downcrossing terminates: exactly one downcrossing, and the upcrossings before it
        assert np.all(rc == 2)
        first_down = (np.pi - np.arcsin(p[:, 1])) / p[:, 0]
        first_down = np.where(first_down > 1e-6, first_down, first_down + 2 * np.pi / p[:, 0])
        ups_before, _ = _expected_counts(p, 0.0)
        a = np.arcsin(p[:, 1])
        ups_before = sum((((a + 2 * np.pi * k) / p[:, 0] > 1e-6) & ((a + 2 * np.pi * k) / p[:, 0] < first_down)).astype(int) for k in range(-1, 8))
        assert np.allclose(out[:, 0, 2], ups_before + 10)
        assert np.allclose(out[:, 0, 0], p[:, 1], atol=1e-7)     # state held at the event: x = thr


def test_direction_sources(B):
    from b200ens import codegen

    c, a, term = codegen.emit_callback(B.ContinuousCallback(_cond, _up, None), 3, 2)
    assert "#define B2_EVENT_DIR 1" in c and "b2_affect_neg" not in a and term == 0
    c, a, term = codegen.emit_callback(B.ContinuousCallback(_cond, None, _down), 3, 2)
    assert "#define B2_EVENT_DIR -1" in c and "b2_affect_neg" not in a and "10" in a
    c, a, term = codegen.emit_callback(B.ContinuousCallback(_cond, _up, _down_term), 3, 2)
    assert "B2_EVENT_DIR" not in c and "#define B2_HAS_AFFECT_NEG 1" in a and "b2_affect_neg" in a and term == 4
    c, a, term = codegen.emit_callback(B.ContinuousCallback(_cond, _up), 3, 2)
    assert "B2_EVENT_DIR" not in c and "b2_affect_neg" not in a
    with pytest.raises(ValueError):
        B.ContinuousCallback(_cond, None, None)
    with pytest.raises(NotImplementedError):
        B.CallbackSet(B.ContinuousCallback(_cond, _up, None), B.ContinuousCallback(_cond, _up))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["both_same", "up_only", "down_only", "distinct", "distinct_terminate_on_down"])
def test_gpu_event_direction_matches_oracle(B, gpu_lib, oracle, case):
    N, T = 2048, 9.0
    u0, p = _params(N, seed=5)
    cb, okw = _cases(B)[case]
    saveat = np.linspace(0.0, T, 10)
    prob = B.ODEProblem(osc, u0[0], (0.0, T), p[0])
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(), trajectories=N, saveat=saveat, dt=0.01,
                  abstol=1e-9, reltol=1e-9, callback=cb)
    ref, rc, st = _oracle_run(oracle, B, cb, okw, u0, p, T, saveat)
    assert np.array_equal(sol.retcodes, rc) and np.array_equal(sol.stats, st)
    assert np.array_equal(sol.u_array, ref)
    assert st[:, 3].min() >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["up_only", "distinct"])
def test_gpu_split_kernel_event_direction(B, gpu_lib, oracle, case):
    """The split kernel (16-species network, Vern7) shares the event search: upcrossing-only bolus, and a two-sided
    callback whose downcrossing affect differs."""
    from b200ens import workloads as W

    N = 512
    u0, p = W.net16_params(N)
    prob = W.net16_problem()

    def cond(u, t, integrator):
        return u[1] - 0.12

    def up(integrator):
        integrator.u[0] = integrator.u[0] + 0.2

    def down(integrator):
        integrator.u[2] = integrator.u[2] + 0.05

    cb = B.ContinuousCallback(cond, up, None) if case == "up_only" else B.ContinuousCallback(cond, up, down)
    okw = dict(event_dir=1) if case == "up_only" else dict()
    saveat = np.linspace(0.0, 10.0, 21)
    kw = dict(trajectories=N, saveat=saveat, dt=0.01, abstol=1e-8, reltol=1e-8, callback=cb)
    sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(split=True), **kw)
    one = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(split=False), **kw)
    model = B.build_model(prob, B.Vern7(), cb)
    ref, rc, st = oracle.solve(None, "Vern7", u0, p, (0.0, 10.0), saveat, 0.01, abstol=1e-8, reltol=1e-8, event=True,
                               fns=oracle_fns(oracle, B, model), **okw)
    assert np.array_equal(sol.retcodes, rc) and np.array_equal(sol.stats, st) and np.array_equal(sol.u_array, ref)
    assert np.array_equal(one.u_array, sol.u_array) and np.array_equal(one.stats, sol.stats)
    assert st[:, 3].max() >= 1


def test_callbackset_of_one_sided_callbacks_lowers_to_one_vector_callback(B):
    """Two upcrossing-only ContinuousCallbacks in a CallbackSet: one vector callback with the common direction."""
    from b200ens import codegen

    cs = B.CallbackSet(B.ContinuousCallback(_cond, _up, None),
                       B.ContinuousCallback(lambda u, t, integrator: u[1] - 0.25, _down, None))
    vcb = cs.continuous[0]
    assert isinstance(vcb, B.VectorContinuousCallback) and vcb.len == 2 and vcb.direction == 1
    c, a, _ = codegen.emit_vector_callback(vcb, 3, 2)
    assert "#define B2_EVENT_DIR 1" in c and "#define B2_NCOND 2" in c and "case 1:" in a and "10" in a
    dn = B.CallbackSet(B.ContinuousCallback(_cond, None, _up), B.ContinuousCallback(lambda u, t, integrator: u[1], None, _down))
    c, a, _ = codegen.emit_vector_callback(dn.continuous[0], 3, 2)
    assert "#define B2_EVENT_DIR -1" in c and "case 0:" in a
