"""Parity at the BASELINE sizes: the headline launch shape -- 1M trajectories, expected-work order on,
specialised adaptive entry, device-resident buffers, exactly what bench.py times -- compared BIT FOR BIT with the
oracle over all 1M trajectories (the C oracle integrates 1M Lorenz trajectories in about a second), Float32 and Float64;
and the 10M-trajectory shape of configs[1] through size-independent properties plus a bit-exact comparison of a
strided 1-in-16 sample."""
import numpy as np
import pytest

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
