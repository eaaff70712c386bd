"""Synthetic inputs of the five BASELINE.json configs, seeded so that CPU
oracle and GPU runs see byte-identical u0 / p matrices.  Model functions are written the way
a user of the reference writes them (test/core.jl:22-27, :39-45) and are traced to CUDA C by
codegen.py; the oracle has its own hand-written C versions (oracle/models.c)."""
import numpy as np

from .api import ContinuousCallback, ODEProblem, SDEProblem


# ---- Lorenz (/root/reference/test/core.jl:22-30)
def lorenz(du, u, p, t):
    du[0] = p[0] * (u[1] - u[0])
    du[1] = u[0] * (p[1] - u[2]) - u[1]
    du[2] = u[0] * u[1] - p[2] * u[2]


def lorenz_problem(dtype=np.float64, tspan=(0.0, 10.0)):
    return ODEProblem(lorenz, np.array([1.0, 0.0, 0.0], dtype=dtype), tspan, np.array([10.0, 28.0, 8.0 / 3.0], dtype=dtype))


def lorenz_params(N, kind="random", seed=0, dtype=np.float64):
    """kind='random': p_i = r_i * (10,28,8/3), r ~ U(0,1)^3 (DiffEqGPU-style sweep);
    kind='ordered': p_i = (10, rho_i, 8/3), rho_i = 56 (i-1)/(N-1)."""
    if kind == "random":
        r = np.random.default_rng(seed).random((N, 3))
        p = r * np.array([10.0, 28.0, 8.0 / 3.0])
    else:
        p = np.empty((N, 3))
        p[:, 0] = 10.0
        p[:, 1] = 56.0 * np.arange(N) / max(N - 1, 1)
        p[:, 2] = 8.0 / 3.0
    u0 = np.tile(np.array([1.0, 0.0, 0.0]), (N, 1))
    return u0.astype(dtype), p.astype(dtype)


# ---- Robertson (/root/reference/test/core.jl:39-46, src/DifferentialEquations.jl:14-23)
def robertson(du, u, p, t):
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2]
    du[1] = p[0] * u[0] - p[1] * u[1] * u[1] - p[2] * u[1] * u[2]
    du[2] = p[1] * u[1] * u[1]


def robertson_problem(tspan=(0.0, 1e5)):
    return ODEProblem(robertson, np.array([1.0, 0.0, 0.0]), tspan, np.array([0.04, 3e7, 1e4]))


def robertson_params(N, seed=1):
    r = np.random.default_rng(seed).random((N, 1))
    p = np.array([0.04, 3e7, 1e4]) * (0.5 + r)
    u0 = np.tile(np.array([1.0, 0.0, 0.0]), (N, 1))
    return u0, p


ROBERTSON_SAVEAT = 10.0 ** np.arange(-5, 6)


# ---- scalar linear (/root/reference/test/core.jl:10-13)
def linear(u, p, t):
    return [p[0] * u[0]]


# ---- GBM and stochastic Lorenz (cfg 4)
def gbm_f(u, p, t):
    return [p[0] * u[0]]


def gbm_g(u, p, t):
    return [p[1] * u[0]]


def gbm_problem(dtype=np.float64):
    return SDEProblem(gbm_f, gbm_g, np.array([1.0], dtype=dtype), (0.0, 1.0), np.array([1.01, 0.87], dtype=dtype))


def gbm_params(N, seed=3, dtype=np.float64):
    r = np.random.default_rng(seed).random((N, 1))
    p = np.array([1.01, 0.87]) * (0.5 + r)
    return np.ones((N, 1), dtype=dtype), p.astype(dtype)


def lorenz_add_f(du, u, p, t):
    lorenz(du, u, p, t)


def lorenz_add_g(u, p, t):
    return [p[3], p[3], p[3]]


def lorenz_additive_problem(dtype=np.float64, tspan=(0.0, 10.0)):
    return SDEProblem(lorenz_add_f, lorenz_add_g, np.array([1.0, 0.0, 0.0], dtype=dtype), tspan,
                      np.array([10.0, 28.0, 8.0 / 3.0, 3.0], dtype=dtype))


def lorenz_additive_params(N, seed=0, dtype=np.float64):
    u0, p3 = lorenz_params(N, "random", seed, np.float64)
    p = np.concatenate([p3, np.full((N, 1), 3.0)], axis=1)
    return u0.astype(dtype), p.astype(dtype)


# ---- 16-species mass-action network with a bolus event (cfg 5; mirrors oracle/models.c net16)
NET16_W = [1.3, 0.42, 6.1, 0.17, 2.9, 0.88, 4.4, 0.23, 7.7, 1.9, 0.35, 3.3, 0.61, 5.2, 1.1]
NET16_V = [0.7, 2.4, 0.19, 3.8, 0.52, 1.6, 0.11, 8.3, 0.93, 0.27, 4.9, 0.44, 2.2, 0.15, 6.6]
NET16_Z = [0.9, 0.31, 2.7, 0.14, 1.8, 0.66, 3.9, 0.21, 5.5, 0.48, 1.2, 0.12, 2.1, 0.77]


def net16(du, u, p, t):
    acc = [0] * 16
    for i in range(15):
        fl = p[0] * NET16_W[i] * u[i] - p[1] * NET16_V[i] * u[i + 1]
        acc[i] = acc[i] - fl
        acc[i + 1] = acc[i + 1] + fl
    for i in range(14):
        r = p[2] * NET16_Z[i] * u[i] * u[i + 1]
        acc[i] = acc[i] - r
        acc[i + 1] = acc[i + 1] - r
        acc[i + 2] = acc[i + 2] + r
    acc[0] = acc[0] - p[3] * u[0]
    for i in range(16):
        du[i] = acc[i]


def net16_problem(tspan=(0.0, 10.0)):
    u0 = np.zeros(16)
    u0[0] = 1.0
    return ODEProblem(net16, u0, tspan, np.array([1.0, 0.5, 0.8, 0.3, 0.25, 0.5]))


def net16_params(N, seed=2):
    rng = np.random.default_rng(seed)
    p = np.tile(np.array([1.0, 0.5, 0.8, 0.3, 0.25, 0.5]), (N, 1))
    p[:, 0] = 10.0 ** rng.uniform(-0.5, 0.5, N)   # sweep the forward and the coupling rate
    p[:, 2] = 10.0 ** rng.uniform(-0.5, 0.5, N)
    u0 = np.zeros((N, 16))
    u0[:, 0] = 1.0
    return u0, p


def net16_callback():
    """condition X1 - theta (p[4]); affect!: X1 += bolus (p[5])."""
    def condition(u, t, integrator):
        return u[0] - integrator.p[4]

    def affect(integrator):
        integrator.u[0] = integrator.u[0] + integrator.p[5]

    return ContinuousCallback(condition, affect)
