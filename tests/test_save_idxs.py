"""solve(...; save_idxs = [...]): only the listed state components are saved (and travel over the host link), in the
listed order.  The kernels compute exactly what they compute without it; the output rows are a column selection of the
full run's rows -- bit for bit, for every kernel family, for the device-side moments and for save_everystep."""
import numpy as np
import pytest


def test_save_idxs_are_validated(B):
    from b200ens import workloads as W

    for bad in ([0, 0], [3], [-1], []):
        with pytest.raises(ValueError):
            B.build_model(W.lorenz_problem(), B.Tsit5(), save_idxs=bad)
    m = B.build_model(W.lorenz_problem(), B.Tsit5(), save_idxs=[2, 0])
    assert m.n_out == 2 and m.info()["cubin_bytes"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lorenz_save_idxs_is_a_column_selection(B, gpu_lib, dtype):
    from b200ens import workloads as W

    N = 40000     # above the work-order threshold: the headline launch shape
    saveat = np.arange(0.0, 10.5, 1.0)
    u0, p = W.lorenz_params(N, "random", seed=31, dtype=dtype)
    kw = dict(trajectories=N, saveat=saveat, dt=0.1, abstol=1e-6, reltol=1e-3)
    full = B.solve(B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), **kw)
    for idxs in ([2], [2, 0], [0, 1, 2]):
        part = B.solve(B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), save_idxs=idxs, **kw)
        assert part.u_array.shape == (N, len(saveat), len(idxs))
        assert np.array_equal(part.u_array, full.u_array[:, :, idxs])
        assert np.array_equal(part.stats, full.stats) and np.array_equal(part.retcodes, full.retcodes)
    # staged outputs and the generic entry (automatic dt) take the same route through the sink
    a = B.solve(B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0[:3000], ps=p[:3000]), B.Tsit5(), B.EnsembleB200(devices=[0], stage_outputs=1),
                trajectories=3000, saveat=saveat, dt=0.1, save_idxs=[1])
    assert np.array_equal(a.u_array, full.u_array[:3000, :, [1]])


@pytest.mark.gpu
def test_save_idxs_other_kernel_families(B, gpu_lib):
    from b200ens import workloads as W

    # split kernel + ContinuousCallback (components owned by different warps, one of them not saved at all)
    N = 600
    u0, p = W.net16_params(N)
    sv = np.linspace(0.0, 10.0, 21)
    kw = dict(trajectories=N, saveat=sv, dt=0.01, abstol=1e-8, reltol=1e-8, callback=W.net16_callback())
    for split in (True, False):
        full = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(devices=[0], split=split), **kw)
        part = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(devices=[0], split=split),
                       save_idxs=[15, 0, 5], **kw)
        assert np.array_equal(part.u_array, full.u_array[:, :, [15, 0, 5]]) and np.array_equal(part.stats, full.stats)
    # Rosenbrock family (saveat points as tstops)
    u0r, pr = W.robertson_params(800)
    kr = dict(trajectories=800, saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6)
    full = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.Rodas5P(), B.EnsembleB200(devices=[0]), **kr)
    part = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.Rodas5P(), B.EnsembleB200(devices=[0]), save_idxs=[1], **kr)
    assert np.array_equal(part.u_array, full.u_array[:, :, [1]])
    # SDE, device Philox
    u0s, ps = W.lorenz_additive_params(2000)
    ks = dict(trajectories=2000, saveat=[5.0, 10.0], dt=1 / 256, seed=11)
    full = B.solve(B.EnsembleProblem(W.lorenz_additive_problem(), u0s=u0s, ps=ps), B.SOSRA(), B.EnsembleB200(devices=[0]), **ks)
    part = B.solve(B.EnsembleProblem(W.lorenz_additive_problem(), u0s=u0s, ps=ps), B.SOSRA(), B.EnsembleB200(devices=[0]), save_idxs=[2, 1], **ks)
    assert np.array_equal(part.u_array, full.u_array[:, :, [2, 1]])


@pytest.mark.gpu
def test_save_idxs_moments_and_everystep(B, gpu_lib):
    from b200ens import workloads as W

    N = 20000
    saveat = np.arange(0.0, 10.5, 1.0)
    u0, p = W.lorenz_params(N, "random", seed=37)
    kw = dict(trajectories=N, saveat=saveat, dt=0.1)
    full = B.solve(B.EnsembleProblem(W.lorenz_problem(), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), **kw)
    summ = B.solve(B.EnsembleProblem(W.lorenz_problem(), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(devices=[0]), summary=True, save_idxs=[2], **kw)
    assert summ.u.shape == (len(saveat), 1)
    assert np.allclose(summ.u[:, 0], full.u_array[:, :, 2].mean(axis=0), rtol=1e-12, atol=1e-12)
    one = B.solve(W.lorenz_problem(), B.Tsit5(), dt=0.1, save_idxs=[0, 2])          # single solve: every accepted step
    ref = B.solve(W.lorenz_problem(), B.Tsit5(), dt=0.1)
    assert np.array_equal(one.t, ref.t)
    assert np.array_equal(np.asarray(one.u), np.asarray(ref.u)[:, [0, 2]])
