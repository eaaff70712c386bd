import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import b200ens as B
from b200ens import workloads as W
import oracle_py as oracle
u0r, pr = W.robertson_params(500)
for tst in (None, [3.3, 1234.5], [3.3], [1234.5]):
    sol = B.solve(B.EnsembleProblem(W.robertson_problem(), u0s=u0r, ps=pr), B.Rodas5P(), B.EnsembleB200(devices=[0]), trajectories=500,
                  saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6, tstops=tst)
    ref, rc, st = oracle.solve("robertson", "Rodas5P", u0r, pr, (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6, tstops=tst)
    bad = np.where((sol.stats != st).any(axis=1))[0]
    print("tstops", tst, "generic", os.environ.get("B200ENS_GENERIC_ENTRY"), "mismatching", len(bad), bad[:10], "maxdiff", np.abs(sol.u_array - ref).max())
    for i in bad[:3]:
        print("  traj", i, "gpu", sol.stats[i], "oracle", st[i])
        # every-step times from the oracle for this trajectory, with and without
