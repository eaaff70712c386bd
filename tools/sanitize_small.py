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
# round-2 additions: tstops + DiscreteCallback, save_idxs (one-thread + split), windowed work order, fused moments
# (+ the fallback through out_u), grouped SDE loops (EM / SOSRA), adaptive SDE
import os
prob = B.ODEProblem(lambda u, p, t: [-p[0] * u[0]], np.array([10.0]), (0.0, 12.0), np.array([0.5, 10.0]))
dcb = B.DiscreteCallback(lambda u, t, integ: (t == 4.0) | (t == 8.0), lambda integ: integ.u.__setitem__(0, integ.u[0] + integ.p[1]))
pp = np.stack([0.2 + np.random.default_rng(0).random(700), np.full(700, 10.0)], axis=1)
s = B.solve(B.EnsembleProblem(prob, u0s=np.full((700, 1), 10.0), ps=pp), B.Tsit5(), B.EnsembleB200(), trajectories=700, saveat=1.0, dt=0.1,
            callback=dcb, tstops=[4.0, 8.0])
print("tstops dosing", (s.retcodes == 1).all(), s.stats[:, 3].mean())
u0n, pn = W.net16_params(200)
s = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0n, ps=pn), B.Vern7(), B.EnsembleB200(split=True), trajectories=200,
            saveat=np.linspace(0, 10, 11), dt=0.01, abstol=1e-6, reltol=1e-6, callback=W.net16_callback(), save_idxs=[15, 0, 5], tstops=[3.3])
print("split save_idxs tstops", (s.retcodes == 1).all(), s.u_array.shape)
u0l, pl = W.lorenz_params(9000, "random", 0, np.float32)
s = B.solve(B.EnsembleProblem(W.lorenz_problem(np.float32), u0s=u0l, ps=pl), B.Tsit5(), B.EnsembleB200(work_order=4096), trajectories=9000,
            saveat=np.arange(0, 10.5, 1.0), dt=0.1, save_idxs=[2])
print("windowed order + save_idxs", (s.retcodes == 1).all(), s.u_array.shape)
os.environ["B200ENS_FUSE_MOMENTS"] = "1"
pl64 = pl.astype(np.float64); pl64[7] = np.nan
s = B.solve(B.EnsembleProblem(W.lorenz_problem(), u0s=u0l.astype(np.float64), ps=pl64), B.Tsit5(), B.EnsembleB200(), trajectories=9000,
            saveat=np.arange(0, 10.5, 1.0), dt=0.1, summary=True)
print("fused moments with a failing trajectory", s.num_monte)
s = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0n, ps=pn), B.Vern7(), B.EnsembleB200(split=True), trajectories=200,
            saveat=np.linspace(0, 10, 11), dt=0.01, abstol=1e-6, reltol=1e-6, callback=W.net16_callback(), summary=True)
print("fused moments split", s.num_monte)
del os.environ["B200ENS_FUSE_MOMENTS"]
u0a, pa = W.lorenz_additive_params(1500)
for alg in (B.EM(), B.SOSRA()):
    s = B.solve(B.EnsembleProblem(W.lorenz_additive_problem(), u0s=u0a, ps=pa), alg, B.EnsembleB200(), trajectories=1500, saveat=[0.3, 1.0],
                dt=1 / 256, seed=3)
    print("sde grouped loop", alg.name, (s.retcodes == 1).all())
pr1 = B.SDEProblem(W.lorenz_add_f, W.lorenz_add_g, np.array([1.0, 0.0, 0.0]), (0.0, 1.0), np.array([10.0, 28.0, 8.0 / 3.0, 3.0]))
s = B.solve(B.EnsembleProblem(pr1, u0s=u0a, ps=pa), B.SOSRA(), B.EnsembleB200(), trajectories=1500, saveat=[1.0], dt=0.01, adaptive=True, seed=3)
print("sde adaptive", (s.retcodes == 1).all())
# session-3 additions: FBDF (plain, with a downcrossing-only ContinuousCallback, fixed step), affect_neg! in the one-thread
# and the split kernel
u0r, pr = W.robertson_params(600)
s = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.FBDF(), B.EnsembleB200(), trajectories=600, saveat=W.ROBERTSON_SAVEAT,
            dt=1e-6, abstol=1e-8, reltol=1e-6)
print("fbdf robertson", (s.retcodes == 1).all(), s.stats[:, 0].mean())
def _decay(du, u, p, t):
    du[0] = -p[0] * (u[0] - u[1]); du[1] = -p[1] * u[1]
pd = np.stack([np.full(400, 200.0), np.linspace(0.3, 2.0, 400)], axis=1)
cbn = B.ContinuousCallback(lambda u, t, integ: u[1] - 0.5, None, lambda integ: integ.u.__setitem__(1, integ.u[1] + 0.4))
s = B.solve(B.EnsembleProblem(B.ODEProblem(_decay, np.array([0.0, 1.0]), (0.0, 5.0), pd[0]), u0s=np.tile([0.0, 1.0], (400, 1)), ps=pd), B.FBDF(),
            B.EnsembleB200(), trajectories=400, saveat=1.0, dt=1e-3, abstol=1e-8, reltol=1e-8, callback=cbn)
print("fbdf downcrossing callback", (s.retcodes == 1).all(), s.stats[:, 3].mean())
s = B.solve(B.EnsembleProblem(B.ODEProblem(_decay, np.array([0.0, 1.0]), (0.0, 1.0), pd[0]), u0s=np.tile([0.0, 1.0], (400, 1)), ps=pd), B.FBDF(),
            B.EnsembleB200(), trajectories=400, saveat=[1.0], dt=1 / 64, adaptive=False)
print("fbdf fixed step", (s.retcodes == 1).all())
cb2 = B.ContinuousCallback(lambda u, t, integ: u[1] - 0.12, lambda integ: integ.u.__setitem__(0, integ.u[0] + 0.2),
                           lambda integ: integ.u.__setitem__(2, integ.u[2] + 0.05))
for split in (True, False):
    s = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0n, ps=pn), B.Vern7(), B.EnsembleB200(split=split), trajectories=200,
                saveat=np.linspace(0, 10, 11), dt=0.01, abstol=1e-6, reltol=1e-6, callback=cb2)
    print("affect_neg split" if split else "affect_neg one-thread", (s.retcodes == 1).all(), s.stats[:, 3].mean())
# end of session 3: Rodas dense output (interpolated saves, DAE), FBDF with a mass matrix, rolled LU (n = 16)
def _rd2(du, u, p, t):
    du[0] = -p[0] * u[0] + p[2] * u[1] * u[2]; du[1] = p[0] * u[0] - p[1] * u[1] ** 2 - p[2] * u[1] * u[2]; du[2] = u[0] + u[1] + u[2] - 1.0
pdae = B.ODEProblem(_rd2, [1.0, 0.0, 0.0], (0.0, 1e3), (0.04, 3e7, 1e4), mass_matrix=np.diag([1.0, 1.0, 0.0]))
for alg in (B.Rodas5P(), B.Rodas4(), B.FBDF()):
    s = B.solve(B.EnsembleProblem(pdae, u0s=u0r[:300], ps=pr[:300]), alg, B.EnsembleB200(), trajectories=300, saveat=[1e-3, 1.0, 40.0, 1e3],
                dt=1e-6, abstol=1e-8, reltol=1e-6)
    print("dae", alg.name, (s.retcodes == 1).all(), float(np.abs(s.u_array.sum(axis=2) - 1).max()))
for alg in (B.Rodas5P(), B.FBDF()):
    s = B.solve(B.EnsembleProblem(W.net16_problem(), u0s=u0n[:64], ps=pn[:64]), alg, B.EnsembleB200(), trajectories=64,
                saveat=np.linspace(0, 10, 5), dt=0.01, abstol=1e-6, reltol=1e-6)
    print("rolled LU n=16", alg.name, (s.retcodes == 1).all())
