"""Small solves for compute-sanitizer (memcheck / racecheck): split kernel with events, work-ordered Tsit5, Rodas5P."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import workloads as W
N = 300
u0, p = W.net16_params(N)
s = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Vern7(), B.EnsembleB200(split=True, work_order=1), trajectories=N,
            saveat=np.linspace(0, 10, 11), dt=0.01, abstol=1e-6, reltol=1e-6, callback=W.net16_callback())
print("split vern7 event", (s.retcodes == 1).all(), s.stats[:, 3].mean())
s = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(split=True), trajectories=N,
            saveat=np.linspace(0, 10, 11), abstol=1e-6, reltol=1e-6)
print("split tsit5 autodt", (s.retcodes == 1).all())
N = 5000
u0, p = W.lorenz_params(N, "random", 0, np.float32)
s = B.solve(B.EnsembleProblem(W.lorenz_problem(np.float32), u0s=u0, ps=p), B.Tsit5(), B.EnsembleB200(work_order=1), trajectories=N,
            saveat=np.arange(0, 10.5, 1.0), dt=0.1)
print("tsit5 work order", (s.retcodes == 1).all())
# session-2 additions: every-step output, vector callback, mass matrix, SRIW1, per-component tolerances
s = B.solve(W.lorenz_problem(np.float64, (0.0, 2.0)), B.Tsit5(), dt=0.1)
print("every step", s.retcode.name, len(s.t))
def _cond(out, u, t, integ):
    out[0] = u[0] - 5.0
    out[1] = u[2] - 20.0
def _aff(integ, idx):
    integ.u[1] = -integ.u[1] if idx == 1 else integ.u[1] * 0.5
s = B.solve(B.EnsembleProblem(W.lorenz_problem(np.float64, (0.0, 2.0)), u0s=u0.astype(np.float64)[:500], ps=p.astype(np.float64)[:500]), B.Tsit5(),
            B.EnsembleB200(), trajectories=500, saveat=0.5, dt=0.05, callback=B.VectorContinuousCallback(_cond, _aff, 2))
print("vector callback", (s.retcodes == 1).all(), s.stats[:, 3].sum())
def _rd(du, u, p, t):
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2]; du[1] = p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2]; du[2] = u[0] + u[1] + u[2] - 1.0
s = B.solve(B.ODEProblem(_rd, [1.0, 0.0, 0.0], (0.0, 1e3), (0.04, 3e7, 1e4), mass_matrix=np.diag([1.0, 1.0, 0.0])), B.Rodas5P(),
            abstol=[1e-8, 1e-12, 1e-8], reltol=1e-6)
print("dae + vector tolerances", s.retcode.name, len(s.t))
u0g, pg = W.gbm_params(2000)
s = B.solve(B.EnsembleProblem(W.gbm_problem(), u0s=u0g, ps=pg), B.SRIW1(), B.EnsembleB200(), trajectories=2000, saveat=[1.0], dt=1 / 64, seed=5)
print("sriw1", (s.retcodes == 1).all())
