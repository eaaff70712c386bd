"""GPU: the expected-work ordering of the trajectory queue (kernels/b2_work.cuh, b200ens_opts.work_order) is a
SCHEDULING decision only: saved values, retcodes and step statistics are bit-identical with the caller's order,
for every adaptive stepper family, and the oracle parity of the ordered run holds like the unordered one's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAVEAT = np.arange(0.0, 10.5, 1.0)


def _lorenz(B, dtype, N, order, alg=None):
    from b200ens import workloads as W

    u0, p = W.lorenz_params(N, "random", seed=3, dtype=dtype)
    eprob = B.EnsembleProblem(W.lorenz_problem(dtype), u0s=u0, ps=p)
    sol = B.solve(eprob, alg or B.Tsit5(), B.EnsembleB200(work_order=order), trajectories=N, saveat=SAVEAT, dt=0.1,
                  abstol=1e-6, reltol=1e-3)
    return u0, p, sol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("N", [1, 33, 4097, 70001])
def test_tsit5_bit_identical_with_and_without_ordering(B, gpu_lib, dtype, N):
    _, _, a = _lorenz(B, dtype, N, 0)
    _, _, b = _lorenz(B, dtype, N, 1)
    assert np.array_equal(a.retcodes, b.retcodes) and np.all(a.retcodes == 1)
    assert np.array_equal(a.stats, b.stats)
    assert np.array_equal(a.u_array, b.u_array)


def test_ordered_run_matches_oracle(B, gpu_lib, oracle):
    N = 40000   # above the auto threshold (32768): the default path of large ensembles
    u0, p, sol = _lorenz(B, np.float64, N, -1)
    ref, rc, st = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), SAVEAT, 0.1, abstol=1e-6, reltol=1e-3)
    assert np.array_equal(sol.retcodes, rc)
    assert np.array_equal(sol.stats[:, :2], st[:, :2])
    err = np.abs(sol.u_array - ref)
    assert np.all(err <= 1e-6 + 1e-3 * np.abs(ref))


def test_vern7_and_rosenbrock_families(B, gpu_lib):
    from b200ens import workloads as W

    _, _, a = _lorenz(B, np.float64, 5000, 0, B.Vern7())
    _, _, b = _lorenz(B, np.float64, 5000, 1, B.Vern7())
    assert np.array_equal(a.u_array, b.u_array) and np.array_equal(a.stats, b.stats)
    N = 3000
    u0, p = W.robertson_params(N)
    sols = []
    for order in (0, 1):
        eprob = B.EnsembleProblem(W.robertson_problem(), u0s=u0, ps=p)
        sols.append(B.solve(eprob, B.Rodas5P(), B.EnsembleB200(work_order=order), trajectories=N,
                            saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6))
    assert np.all(sols[0].retcodes == 1)
    assert np.array_equal(sols[0].u_array, sols[1].u_array) and np.array_equal(sols[0].stats, sols[1].stats)


def test_nan_and_failing_trajectories_keep_their_slots(B, gpu_lib):
    """Trajectories whose proxy is NaN/inf sort first; their outputs must still land at their own index."""
    from b200ens import workloads as W

    N = 2048
    u0, p = W.lorenz_params(N, "random", seed=5, dtype=np.float64)
    u0[7, 0] = np.nan
    p[100, 1] = np.inf
    outs = []
    for order in (0, 1):
        eprob = B.EnsembleProblem(W.lorenz_problem(np.float64), u0s=u0, ps=p)
        outs.append(B.solve(eprob, B.Tsit5(), B.EnsembleB200(work_order=order), trajectories=N, saveat=SAVEAT, dt=0.1))
    assert np.array_equal(outs[0].retcodes, outs[1].retcodes)
    assert outs[1].retcodes[7] != 1 and outs[1].retcodes[100] != 1
    assert np.array_equal(outs[0].u_array, outs[1].u_array, equal_nan=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_windowed_ordering_is_bit_identical(B, gpu_lib, dtype):
    """The order is established inside windows of consecutive trajectories (b2_work.cuh; keeps the output rows in flight
    within L2's reach).  work_order > 1 = explicit window: several windows, a ragged last one, and one tile per window
    must give the bits of the caller's order."""
    N = 70001
    _, _, ref = _lorenz(B, dtype, N, 0)
    for window in (4096, 8192, 65536):
        _, _, w = _lorenz(B, dtype, N, window)
        assert np.array_equal(ref.retcodes, w.retcodes) and np.all(w.retcodes == 1), window
        assert np.array_equal(ref.stats, w.stats), window
        assert np.array_equal(ref.u_array, w.u_array), window


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_generic_entry_gives_the_specialised_entrys_bits(B, gpu_lib, dtype):
    """Solves with tstops, fused moments, save_everystep or an automatic dt run through the generic kernel entry; the
    headline shape runs through the compile-time specialised one.  Same template, same bits (B200ENS_GENERIC_ENTRY=1
    forces the generic entry)."""
    import os

    from b200ens import workloads as W

    N = 40000
    _, _, spec = _lorenz(B, dtype, N, -1)
    u0r, pr = W.robertson_params(2000)
    kr = dict(trajectories=2000, saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6)
    spec_r = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.Rodas5P(), B.EnsembleB200(), **kr) if dtype == np.float64 else None
    os.environ["B200ENS_GENERIC_ENTRY"] = "1"
    try:
        _, _, gen = _lorenz(B, dtype, N, -1)
        gen_r = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.Rodas5P(), B.EnsembleB200(), **kr) if dtype == np.float64 else None
    finally:
        del os.environ["B200ENS_GENERIC_ENTRY"]
    assert np.array_equal(spec.u_array, gen.u_array) and np.array_equal(spec.stats, gen.stats)
    if spec_r is not None:
        assert np.array_equal(spec_r.u_array, gen_r.u_array) and np.array_equal(spec_r.stats, gen_r.stats)
