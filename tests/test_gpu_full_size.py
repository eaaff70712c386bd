"""Parity at the BASELINE sizes: the headline launch shape -- 1M trajectories, expected-work order on,
specialised adaptive entry, device-resident buffers, exactly what bench.py times -- compared BIT FOR BIT with the
oracle over all 1M trajectories (the C oracle integrates 1M Lorenz trajectories in about a second), Float32 and Float64;
and the 10M-trajectory shape of configs[1] through size-independent properties plus a bit-exact comparison of a
strided 1-in-16 sample."""
import numpy as np
import pytest

from helpers import oracle_fns

pytestmark = pytest.mark.gpu

SAVEAT = np.arange(0.0, 10.5, 1.0)


def _device_solve(B, gpu_lib, dtype, u0, p):
    import torch
    from b200ens import workloads as W

    N = u0.shape[0]
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    model = B.build_model(W.lorenz_problem(dtype), B.Tsit5())
    d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
    d_save = torch.from_numpy(SAVEAT.astype(dtype)).cuda()
    d_out = torch.empty((N, 11, 3), dtype=tdt, device="cuda")
    d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
    o = gpu_lib.default_opts()
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
    tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), 11, d_out.data_ptr(),
                            d_rc.data_ptr(), d_st.data_ptr())
    assert tm.launches == 3           # work keys + scatter + the ensemble kernel: the expected-work order is on
    return d_out.cpu().numpy(), d_rc.cpu().numpy(), d_st.cpu().numpy()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["random", "ordered"])
def test_headline_shape_1M_bit_identical_to_the_oracle(B, gpu_lib, oracle, dtype, kind):
    from b200ens import workloads as W

    N = 1_000_000
    u0, p = W.lorenz_params(N, kind, seed=0, dtype=dtype)
    out, rc, st = _device_solve(B, gpu_lib, dtype, u0, p)
    ref, rrc, rst = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), SAVEAT, 0.1, dtype=dtype)
    assert np.array_equal(rc, rrc) and np.all(rc == 1)
    assert np.array_equal(st[:, :3], rst[:, :3])          # naccept, nreject, nf of every trajectory
    assert np.array_equal(out, ref)                         # all 33M saved values, bit for bit


def test_10M_trajectories_properties_and_sampled_bit_parity(B, gpu_lib, oracle):
    from b200ens import workloads as W

    N = 10_000_000
    u0, p = W.lorenz_params(N, "random", seed=1, dtype=np.float32)
    out, rc, st = _device_solve(B, gpu_lib, np.float32, u0, p)
    assert np.all(rc == 1) and np.all(np.isfinite(out))
    assert np.array_equal(out[:, 0, :], u0)                 # first saved value is u0 itself (test/core.jl:34), for every trajectory
    assert st[:, 0].min() >= 10 and np.all(st[:, 2] == 1 + 6 * (st[:, 0] + st[:, 1]))   # nf = 1 + 6 attempts (Tsit5 FSAL)
    sl = slice(0, N, 16)
    ref, rrc, rst = oracle.solve("lorenz", "Tsit5", u0[sl], p[sl], (0.0, 10.0), SAVEAT, 0.1, dtype=np.float32)
    assert np.array_equal(out[sl], ref) and np.array_equal(st[sl, :3], rst[:, :3])


def test_config3_robertson_rodas5p_1M_bit_identical(B, gpu_lib, oracle):
    """BASELINE config 3 at its full size: 1M Robertson trajectories, Rodas5P with the analytic Jacobian, abstol 1e-8,
    reltol 1e-6, t in [0, 1e5], saveat 10^(-5..5) -- every saved value, retcode and step count against the oracle."""
    from b200ens import workloads as W

    N = 1_000_000
    u0, p = W.robertson_params(N)
    eprob = B.EnsembleProblem(W.robertson_problem(), u0s=u0, ps=p)
    sol = B.solve(eprob, B.Rodas5P(), B.EnsembleB200(), trajectories=N, saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6)
    # the oracle integrates the SAME emitted expression tree (tests/helpers.py); its hand-written Robertson model associates
    # p1*u1*u1 differently and is compared at the solver tolerance in test_gpu_parity_algs.py
    model = B.build_model(W.robertson_problem(), B.Rodas5P())
    ref, rc, st = oracle.solve(None, "Rodas5P", u0, p, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6,
                               fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats[:, :3], st[:, :3])
    assert np.array_equal(sol.u_array, ref)
    assert np.abs(sol.u_array.sum(axis=2) - 1.0).max() < 1e-6      # the Robertson invariant y1 + y2 + y3 = 1 on all 11M saved states


def test_config4_gbm_em_10M_device_philox(B, gpu_lib, oracle):
    """BASELINE config 4 at its full size: 10M GBM paths, Euler-Maruyama, dt = 1/256, Philox4x32-10 on the device.  Three
    200k-path windows (same global trajectory indices = same Philox streams) against the oracle to the ulp-level tolerance of
    the Box-Muller functions, and the law of the whole ensemble."""
    from b200ens import workloads as W

    N = 10_000_000
    u0, p = W.gbm_params(N, dtype=np.float32)
    eprob = B.EnsembleProblem(W.gbm_problem(np.float32), u0s=u0, ps=p)
    sol = B.solve(eprob, B.EM(), B.EnsembleB200(), trajectories=N, saveat=[1.0], dt=1 / 256, seed=7)
    assert np.all(sol.retcodes == 1) and np.all(np.isfinite(sol.u_array))
    mu = p[:, 0].astype(np.float64)
    # Euler-Maruyama's own expectation is exact: E[u_256] = (1 + mu dt)^256 (increments are independent and zero-mean); it sits
    # mu^2 / 512 ~ 2e-3 below exp(mu), the weak O(dt) bias.  Standard error of the mean over 10M paths: 3.4e-4.
    assert abs(np.mean(sol.u_array[:, 0, 0] / (1.0 + mu / 256) ** 256) - 1.0) < 1.4e-3
    assert abs(np.mean(sol.u_array[:, 0, 0] / np.exp(mu)) - 1.0) < 4e-3
    # three windows of 200k paths (start, middle, end of the ensemble; the oracle gets the window's global index base, so it
    # draws the same Philox streams)
    worst = 0.0
    for lo in (0, 4_900_000, N - 200_000):
        hi = lo + 200_000
        ref, rc, _ = oracle.solve("gbm", "EM", u0[lo:hi], p[lo:hi], (0.0, 1.0), [1.0], 1 / 256, dtype=np.float32, seed=7, adaptive=False,
                                  traj_offset=lo)
        a, b = sol.u_array[lo:hi, 0, 0].astype(np.float64), ref[:, 0, 0].astype(np.float64)
        worst = max(worst, float((np.abs(a - b) / np.abs(b)).max()))
    assert worst < 5e-4, worst


def test_config5_net16_vern7_event_1M_bit_identical(B, gpu_lib, oracle):
    """BASELINE config 5 at its full trajectory count: 1M trajectories of the 16-species network, Vern7, abstol = reltol =
    1e-8, the bolus ContinuousCallback (~17 events per trajectory), split kernel -- saved values (11 save points to keep the
    comparison arrays at 1.4 GB), retcodes, step / RHS / event counts bit for bit against the oracle."""
    from b200ens import workloads as W

    N = 1_000_000
    u0, p = W.net16_params(N)
    sv = np.linspace(0.0, 10.0, 11)
    eprob = B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p)
    sol = B.solve(eprob, B.Vern7(), B.EnsembleB200(), trajectories=N, saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8, callback=W.net16_callback())
    model = B.build_model(W.net16_problem(), B.Vern7(), W.net16_callback())
    ref, rc, st = oracle.solve(None, "Vern7", u0, p, (0.0, 10.0), sv, 0.01, abstol=1e-8, reltol=1e-8, event=True,
                               fns=oracle_fns(oracle, B, model))
    assert np.array_equal(sol.retcodes, rc) and np.all(rc == 1)
    assert np.array_equal(sol.stats, st)                    # naccept, nreject, nf, nevents
    assert np.array_equal(sol.u_array, ref)
    assert st[:, 3].mean() > 10
