"""Default-usage probe: the calls a DifferentialEquations.jl user would type first, through the host mirror on the GPU,
checked against scipy at tight tolerance.  Prints one line per scenario."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import workloads as W
from scipy.integrate import solve_ivp

def lorenz(du, u, p, t):
    du[0] = p[0] * (u[1] - u[0]); du[1] = u[0] * (p[1] - u[2]) - u[1]; du[2] = u[0] * u[1] - p[2] * u[2]
def rober(du, u, p, t):
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2]; du[1] = p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2]; du[2] = p[1] * u[1] ** 2
def vdp(du, u, p, t):
    du[0] = u[1]; du[1] = p[0] * ((1 - u[0] ** 2) * u[1] - u[0])
def ref(f, u0, tspan, p, method="DOP853"):
    def g(t, y):
        d = [0.0] * len(y); f(d, y, p, t); return d
    return solve_ivp(g, tspan, u0, method=method, rtol=1e-11, atol=1e-13).y[:, -1]

def show(name, sol, truth=None):
    last = np.asarray(sol.u[-1], dtype=np.float64)
    err = None if truth is None else float(np.max(np.abs(last - truth) / (1e-6 + np.abs(truth))))
    print(f"{name:55s} retcode={sol.retcode.name:12s} stats={sol.stats} relerr={err}", flush=True)

P = (10.0, 28.0, 8 / 3)
t_l = ref(lorenz, [1.0, 0.0, 0.0], (0.0, 1.0), P)
for alg in (B.Tsit5(), B.Vern7(), B.Rosenbrock23(), B.Rodas4(), B.Rodas5(), B.Rodas5P()):
    show(f"lorenz (0,1) {alg.name} defaults", B.solve(B.ODEProblem(lorenz, [1.0, 0.0, 0.0], (0.0, 1.0), P), alg), t_l)
show("lorenz (0,1) Tsit5 Float32 defaults", B.solve(B.ODEProblem(lorenz, np.array([1.0, 0.0, 0.0], dtype=np.float32), (0.0, 1.0), np.array(P, dtype=np.float32)), B.Tsit5()), t_l)
show("lorenz (0,100) Tsit5 defaults (chaotic: retcode only)", B.solve(B.ODEProblem(lorenz, [1.0, 0.0, 0.0], (0.0, 100.0), P), B.Tsit5()))
t_r = ref(rober, [1.0, 0.0, 0.0], (0.0, 1e5), (0.04, 3e7, 1e4), "Radau")
for alg in (B.Rosenbrock23(), B.Rodas4(), B.Rodas5(), B.Rodas5P()):
    show(f"robertson (0,1e5) {alg.name} defaults", B.solve(B.ODEProblem(rober, [1.0, 0.0, 0.0], (0.0, 1e5), (0.04, 3e7, 1e4)), alg), t_r)
    show(f"robertson (0,1e5) {alg.name} tol 1e-8", B.solve(B.ODEProblem(rober, [1.0, 0.0, 0.0], (0.0, 1e5), (0.04, 3e7, 1e4)), alg, abstol=1e-10, reltol=1e-8), t_r)
t_v = ref(vdp, [2.0, 0.0], (0.0, 50.0), (100.0,), "Radau")
for alg in (B.Rodas5P(), B.Rosenbrock23(), B.Tsit5()):
    show(f"van der Pol mu=100 (0,50) {alg.name} tol 1e-8", B.solve(B.ODEProblem(vdp, [2.0, 0.0], (0.0, 50.0), (100.0,)), alg, abstol=1e-8, reltol=1e-8, maxiters=10**7), t_v)
show("no-parameter problem u' = -u, Tsit5", B.solve(B.ODEProblem(lambda u, p, t: [-u[0]], 1.0, (0.0, 2.0)), B.Tsit5(), abstol=1e-10, reltol=1e-10), np.array([np.exp(-2.0)]))
show("maxiters=5 -> MaxIters", B.solve(B.ODEProblem(lorenz, [1.0, 0.0, 0.0], (0.0, 10.0), P), B.Tsit5(), maxiters=5))
show("gbm EM dt=1/256 (one path)", B.solve(W.gbm_problem(), B.EM(), dt=1 / 256, seed=3))
show("additive lorenz SOSRA dt=1/256", B.solve(W.lorenz_additive_problem(np.float64), B.SOSRA(), dt=1 / 256, seed=3))
