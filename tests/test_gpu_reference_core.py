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
