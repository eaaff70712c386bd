import sys, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import b200ens as B, oracle_py as oracle
from b200ens import workloads as W
SAVEAT = np.arange(0.0, 10.5, 1.0)
N = 20011
u0, p = W.lorenz_params(N, "random", seed=9, dtype=np.float32)
eprob = B.EnsembleProblem(W.lorenz_problem(np.float32), u0s=u0, ps=p)
ref, rc, st = oracle.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), SAVEAT, 0.1, dtype=np.float32)
for rep in range(3):
    for x2 in (True, False):
        sol = B.solve(eprob, B.Tsit5(), B.EnsembleB200(packed_x2=x2), trajectories=N, saveat=SAVEAT, dt=0.1, abstol=1e-6, reltol=1e-3)
        bad = np.where((sol.stats[:, :3] != st[:, :3]).any(1))[0]
        badu = np.where((sol.u_array != ref).any((1, 2)))[0]
        print("x2", x2, "rep", rep, "stat mismatches", len(bad), bad[:10], "u mismatches", len(badu), badu[:10])
        for i in bad[:5]:
            print("   ", i, sol.stats[i], st[i], "rc", sol.retcodes[i], rc[i])
sol = B.solve(eprob, B.Tsit5(), B.EnsembleB200(packed_x2=True), trajectories=N, saveat=SAVEAT, dt=0.1, abstol=1e-6, reltol=1e-3)
d = (sol.u_array != ref)
print("mismatch per save index", d.any(2).sum(0))
print("mismatch per component", d.any(1).sum(0))
okst = (sol.stats[:, :3] == st[:, :3]).all(1)
print("max abs diff (same step counts)", np.abs(sol.u_array[okst] - ref[okst]).max(), "rel", (np.abs(sol.u_array[okst] - ref[okst]) / (1e-6 + np.abs(ref[okst]))).max())
i = 0
print(sol.u_array[i], ref[i], sol.stats[i], st[i])
print("timing", sol.timing)
