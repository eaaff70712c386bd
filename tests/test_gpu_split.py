"""GPU: the SPLIT kernel (kernels/b2_split.cuh, b2_ode_driver_split.cuh: one trajectory per lane of a 4-warp CTA,
components of the state split over the warps) runs every component through the expression tree of the one-thread
kernel, so saved values, retcodes and statistics must be BIT-IDENTICAL to it -- and through it to the CPU oracle
(test_gpu_parity_algs.py::test_net16_vern7_callback runs the default, which picks the split kernel for config 5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pair(B, prob, alg, u0, p, **kw):
    N = u0.shape[0]
    out = []
    for split in (False, True):
        sol = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), alg, B.EnsembleB200(split=split), trajectories=N, **kw)
        out.append(sol)
    return out


def _same(a, b):
    assert np.array_equal(a.retcodes, b.retcodes)
    assert np.array_equal(a.stats, b.stats)
    assert np.array_equal(a.u_array, b.u_array, equal_nan=True)


@pytest.mark.parametrize("alg", ["Tsit5", "Vern7"])
@pytest.mark.parametrize("event", [False, True])
def test_net16_split_bit_identical(B, gpu_lib, alg, event):
    from b200ens import workloads as W

    N = 700   # not a multiple of 32: the last CTA runs with parked lanes
    u0, p = W.net16_params(N)
    cb = W.net16_callback() if event else None
    a, b = _pair(B, W.net16_problem(), getattr(B, alg)(), u0, p, saveat=np.linspace(0.0, 10.0, 21), dt=0.01,
                 abstol=1e-8, reltol=1e-8, callback=cb)
    assert np.all(a.retcodes == 1)
    if event:
        assert a.stats[:, 3].max() >= 1
    _same(a, b)


def test_split_options(B, gpu_lib):
    """automatic initial dt, fixed dt, saveat points as tstops, expected-work ordering, terminate!"""
    from b200ens import workloads as W

    N = 333
    u0, p = W.net16_params(N)
    prob = W.net16_problem()
    sv = np.linspace(0.0, 10.0, 11)
    _same(*_pair(B, prob, B.Tsit5(), u0, p, saveat=sv, abstol=1e-7, reltol=1e-7))                     # dt=None
    _same(*_pair(B, prob, B.Tsit5(), u0, p, saveat=sv, dt=0.01, adaptive=False))
    _same(*_pair(B, prob, B.Vern7(), u0, p, saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8, save_tstops=True))
    a = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(split=True, work_order=1), trajectories=N,
                saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8)
    b = B.solve(B.EnsembleProblem(prob, u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(split=False, work_order=0), trajectories=N,
                saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8)
    _same(a, b)
    cbt = B.ContinuousCallback(lambda u, t, integrator: u[0] - integrator.p[4], lambda integrator: B.terminate_b(integrator))
    x, y = _pair(B, prob, B.Tsit5(), u0, p, saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8, callback=cbt)
    assert np.any(x.retcodes == 2)
    _same(x, y)


def _ring6(du, u, p, t):
    # six coupled nonlinear relaxations: n_state not divisible by 4 (two padded components in warp 3... and warp 2)
    for i in range(6):
        du[i] = p[0] * (u[(i + 1) % 6] - u[i]) - p[1] * u[i] * u[(i + 5) % 6] + p[2]


def test_split_state_count_not_divisible_by_four(B, gpu_lib):
    N = 257
    rng = np.random.default_rng(11)
    u0 = rng.random((N, 6))
    p = np.stack([1.0 + rng.random(N), 0.5 + rng.random(N), 0.1 * rng.random(N)], axis=1)
    prob = B.ODEProblem(_ring6, np.ones(6), (0.0, 5.0), np.array([1.0, 1.0, 0.1]))
    cb = B.ContinuousCallback(lambda u, t, integrator: u[5] - 0.4, lambda integrator: integrator.u.__setitem__(5, integrator.u[5] + 0.3))
    for alg in (B.Tsit5(), B.Vern7()):
        _same(*_pair(B, prob, alg, u0, p, saveat=np.linspace(0.0, 5.0, 26), dt=0.01, abstol=1e-9, reltol=1e-9, callback=cb))


def test_split_vector_continuous_callback_bit_identical(B, gpu_lib):
    """VectorContinuousCallback (qa.jl:124) in the split kernel: the event search is the one shared with the one-thread
    kernel (kernels/b2_control.cuh), so event counts, per-index affect!, per-index terminate! and every saved value must be
    bit-identical between the two kernels -- on the 16-species network (two thresholds on different species) and on a
    system whose size is not divisible by four."""
    from b200ens import workloads as W

    N = 500
    u0, p = W.net16_params(N)

    def vcond(out, u, t, integrator):
        out[0] = u[0] - integrator.p[4]
        out[1] = u[5] - 0.02

    def vaffect(integrator, idx):
        if idx == 1:
            integrator.u[0] = integrator.u[0] + integrator.p[5]
        else:
            integrator.u[5] = integrator.u[5] * 0.5

    cb = B.VectorContinuousCallback(vcond, vaffect, 2)
    a, b = _pair(B, W.net16_problem(), B.Vern7(), u0, p, saveat=np.linspace(0.0, 10.0, 21), dt=0.01, abstol=1e-8, reltol=1e-8,
                 callback=cb)
    assert np.all(a.retcodes == 1) and a.stats[:, 3].max() >= 2
    _same(a, b)

    def vaffect_term(integrator, idx):
        if idx == 1:
            integrator.u[0] = integrator.u[0] + integrator.p[5]
        else:
            B.terminate_b(integrator)

    cbt = B.VectorContinuousCallback(vcond, vaffect_term, 2)
    x, y = _pair(B, W.net16_problem(), B.Tsit5(), u0, p, saveat=np.linspace(0.0, 10.0, 21), dt=0.01, abstol=1e-8, reltol=1e-8,
                 callback=cbt)
    assert np.any(x.retcodes == 2)
    _same(x, y)
    # a CallbackSet of two ContinuousCallbacks is lowered to one vector callback by the host mirror
    rng = np.random.default_rng(12)
    u6 = rng.random((257, 6))
    p6 = np.stack([1.0 + rng.random(257), 0.5 + rng.random(257), 0.1 * rng.random(257)], axis=1)
    prob6 = B.ODEProblem(_ring6, np.ones(6), (0.0, 5.0), np.array([1.0, 1.0, 0.1]))
    cs = B.CallbackSet(B.ContinuousCallback(lambda u, t, integrator: u[5] - 0.4, lambda integrator: integrator.u.__setitem__(5, integrator.u[5] + 0.3)),
                       B.ContinuousCallback(lambda u, t, integrator: u[1] - 0.6, lambda integrator: integrator.u.__setitem__(1, integrator.u[1] - 0.2)))
    _same(*_pair(B, prob6, B.Vern7(), u6, p6, saveat=np.linspace(0.0, 5.0, 26), dt=0.01, abstol=1e-9, reltol=1e-9, callback=cs))


@pytest.mark.parametrize("ip", [3, 20])
def test_split_event_search_any_interp_points(B, gpu_lib, ip):
    """interp_points other than the default 10: the split kernel equals the one-thread kernel bit for bit.  (Evaluating
    the samples concurrently in the four warps was tried in round 2: 33.9 vs 32.1 ms per 200k, not kept.)"""
    from b200ens import workloads as W

    N = 300
    u0, p = W.net16_params(N)
    cb = B.ContinuousCallback(lambda u, t, integrator: u[0] - integrator.p[4],
                              lambda integrator: integrator.u.__setitem__(0, integrator.u[0] + integrator.p[5]), interp_points=ip)
    a, b = _pair(B, W.net16_problem(), B.Vern7(), u0, p, saveat=np.linspace(0.0, 10.0, 11), dt=0.01, abstol=1e-8, reltol=1e-8, callback=cb)
    assert np.all(a.retcodes == 1) and a.stats[:, 3].max() >= 1
    _same(a, b)
