"""Edge cases on the GPU path: empty and tiny ensembles, no saveat points, saveat away from the end points,
the largest supported system, parameter-free problems, and fixed-step failure codes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_empty_and_single_trajectory(B, gpu_lib, oracle):
    from b200ens import workloads as W

    prob = W.lorenz_problem()
    u0, p = W.lorenz_params(1, "ordered")
    m = B.build_model(prob, B.Tsit5())
    o = B._lib.default_opts()
    o.t0, o.t1, o.dt = 0.0, 1.0, 0.1
    out, rc, st, tm = m.solve(o, u0[:0], p[:0], [1.0])
    assert out.shape == (0, 1, 3) and rc.shape == (0,)
    out, rc, st, tm = m.solve(o, u0, p, [1.0])
    ref, rc2, _ = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 1.0), [1.0], 0.1)
    assert rc[0] == 1 and np.array_equal(out, ref)


def test_no_save_points_and_interior_saveat(B, gpu_lib, oracle):
    from b200ens import workloads as W

    N = 300
    u0, p = W.lorenz_params(N, "random", seed=21)
    m = B.build_model(W.lorenz_problem(), B.Tsit5())
    o = B._lib.default_opts()
    o.t0, o.t1, o.dt = 0.0, 10.0, 0.1
    out, rc, st, _ = m.solve(o, u0, p, np.zeros(0))
    assert out.shape == (N, 0, 3) and np.all(rc == 1)
    ref_st = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), [10.0], 0.1)[2]
    assert np.array_equal(st[:, :2], ref_st[:, :2])
    saveat = [0.25, 3.3, 9.99]                      # neither end point is saved when it is not in saveat (A.2)
    out, rc, st, _ = m.solve(o, u0, p, saveat)
    ref, rc2, _ = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), saveat, 0.1)
    assert np.array_equal(out, ref) and np.array_equal(rc, rc2)


def test_largest_system_and_no_parameters(B, gpu_lib):
    """n_state = 32 (the ABI maximum), n_param = 0: 32 decoupled decays u_i' = -(i+1)/8 u_i, closed form."""
    n = 32

    def decay(u, p, t):
        return [-(i + 1) / 8.0 * u[i] for i in range(n)]

    prob = B.ODEProblem(decay, np.ones(n), (0.0, 2.0))
    sol = B.solve(B.EnsembleProblem(prob), B.Tsit5(), B.EnsembleB200(), trajectories=257, saveat=[1.0, 2.0], dt=0.01,
                  abstol=1e-10, reltol=1e-10)
    exact = np.exp(-np.outer([1.0, 2.0], (np.arange(n) + 1) / 8.0))
    assert np.all(sol.retcodes == 1)
    assert np.abs(sol.u_array - exact[None]).max() < 1e-8


def test_fixed_step_unstable_and_dt_nan(B, gpu_lib, oracle):
    from b200ens import workloads as W

    prob = B.ODEProblem(W.linear, [1.0], (0.0, 10.0), [400.0])
    sol = B.solve(prob, B.Tsit5(), dt=0.5, adaptive=False, saveat=[10.0])
    ref, rc, _ = oracle.solve("linear", "Tsit5", [[1.0]], [[400.0]], (0.0, 10.0), [10.0], 0.5, adaptive=False)
    assert int(sol.retcode) == rc[0]
    m = B.build_model(W.lorenz_problem(), B.Tsit5())
    o = B._lib.default_opts()
    o.t0, o.t1, o.dt = 0.0, 1.0, 0.1
    with pytest.raises(B.B200EnsError):              # t1 <= t0 is rejected by the library, not silently "solved"
        o2 = B._lib.default_opts()
        o2.t0, o2.t1, o2.dt = 1.0, 0.0, 0.1
        m.solve(o2, np.ones((1, 3)), np.ones((1, 3)), [0.5])


def test_device_resident_entry_matches_host_entry(B, gpu_lib):
    import torch
    from b200ens import workloads as W

    N = 5000
    u0, p = W.lorenz_params(N, "random", seed=8, dtype=np.float32)
    m = B.build_model(W.lorenz_problem(np.float32), B.Tsit5())
    o = B._lib.default_opts()
    o.t0, o.t1, o.dt = 0.0, 10.0, 0.1
    saveat = np.arange(0, 10.5, 1.0, dtype=np.float32)
    out, rc, st, _ = m.solve(o, u0, p, saveat)
    d_u0, d_p, d_s = (torch.from_numpy(x).cuda() for x in (u0, p, saveat))
    d_out = torch.empty((N, 11, 3), dtype=torch.float32, device="cuda")
    d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        m.solve_device(o, 0, s.cuda_stream, N, d_u0.data_ptr(), d_p.data_ptr(), d_s.data_ptr(), 11, d_out.data_ptr(),
                       d_rc.data_ptr(), d_st.data_ptr(), timed=False)
    s.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), out) and np.array_equal(d_rc.cpu().numpy(), rc)
    assert np.array_equal(d_st.cpu().numpy(), st)


@pytest.mark.parametrize("alg", ["Tsit5", "Vern7", "Rodas5P"])
def test_automatic_initial_dt_matches_oracle(B, gpu_lib, oracle, alg):
    """No dt given (test/core.jl:14,32,47 call solve without dt): per-trajectory Hairer-Norsett-Wanner initial step
    on the device, identical to the oracle's."""
    from b200ens import workloads as W

    N = 2000
    u0, p = W.lorenz_params(N, "random", seed=12)
    saveat = np.arange(0.0, 1.05, 0.1)
    sol = B.solve(B.EnsembleProblem(W.lorenz_problem(tspan=(0.0, 1.0)), u0s=u0, ps=p), getattr(B, alg)(), B.EnsembleB200(),
                  trajectories=N, saveat=saveat, abstol=1e-8, reltol=1e-8)
    ref, rc, st = oracle.solve("lorenz", alg, u0, p, (0.0, 1.0), saveat, 0.0, abstol=1e-8, reltol=1e-8)
    assert np.all(sol.retcodes == 1) and np.array_equal(sol.retcodes, rc)
    assert np.array_equal(sol.stats[:, :3], st[:, :3])
    assert np.abs(sol.u_array - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


def test_pinned_and_pageable_host_buffers_agree(B, gpu_lib):
    """b200ens_solve takes any host pointer: pinned arrays go straight to the copy engines, pageable ones are staged
    through pinned bounce buffers with a parallel memcpy.  Both must give the same bytes."""
    from b200ens import workloads as W

    N = 300007
    u0, p = W.lorenz_params(N, "random", seed=31, dtype=np.float32)
    m = B.build_model(W.lorenz_problem(np.float32), B.Tsit5())
    o = B._lib.default_opts()
    o.t0, o.t1, o.dt = 0.0, 10.0, 0.1
    saveat = np.arange(0, 10.5, 1.0)
    out_a, rc_a, st_a, _ = m.solve(o, u0, p, saveat)                       # pageable numpy arrays
    pin = lambda a: (lambda b: (b.__setitem__(slice(None), a), b)[1])(B.pinned_empty(a.shape, a.dtype))
    out_p = B.pinned_empty((N, 11, 3), np.float32)
    rc_p = B.pinned_empty((N,), np.int32)
    st_p = B.pinned_empty((N, 4), np.int32)
    m.solve(o, pin(u0), pin(p), saveat, out=out_p, rc=rc_p, stats=st_p)
    assert np.array_equal(out_a, out_p) and np.array_equal(rc_a, rc_p) and np.array_equal(st_a, st_p)
