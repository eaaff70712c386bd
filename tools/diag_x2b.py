import sys
if len(sys.argv) > 1 and sys.argv[1] == "torch":
    import torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import b200ens as B, oracle_py as oracle
from b200ens import workloads as W
SAVEAT = np.arange(0.0, 10.5, 1.0)
N = 20011
u0, p = W.lorenz_params(N, "random", seed=9, dtype=np.float32)
eprob = B.EnsembleProblem(W.lorenz_problem(np.float32), u0s=u0, ps=p)
ref, rc, st = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), SAVEAT, 0.1, dtype=np.float32)
sol = B.solve(eprob, B.Tsit5(), B.EnsembleB200(packed_x2=True), trajectories=N, saveat=SAVEAT, dt=0.1, abstol=1e-6, reltol=1e-3)
print(sys.argv[1:], "u mismatches", int((sol.u_array != ref).any((1, 2)).sum()), "regs", sol.timing["regs"])
print([l.split()[-1] for l in open("/proc/self/maps") if "nvrtc" in l and "r-xp" in l])
