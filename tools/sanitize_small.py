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
