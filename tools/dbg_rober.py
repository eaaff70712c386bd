import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import b200ens as B
from b200ens import workloads as W
import oracle_py as oracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = 20000
u0, p = W.robertson_params(N)
ref, rc, st = oracle.solve("robertson", "Rodas5P", u0[:M], p[:M], (0.0, 1e5), W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6)
print("oracle first", st[:4].tolist())
for wo in (-1, 0):
    for n in (N, M):
        eprob = B.EnsembleProblem(W.robertson_problem(), u0s=u0[:n], ps=p[:n])
        sol = B.solve(eprob, B.Rodas5P(), B.EnsembleB200(work_order=wo), trajectories=n, saveat=W.ROBERTSON_SAVEAT, dt=1e-6, abstol=1e-8, reltol=1e-6)
        bad = np.nonzero((sol.stats[:M, :3] != st[:, :3]).any(axis=1))[0]
        print("wo", wo, "n", n, "mismatch", bad.size, bad[:10].tolist(), sol.stats[bad[:3]].tolist(), st[bad[:3]].tolist(),
              "values equal on matching:", np.array_equal(np.delete(sol.u_array[:M], bad, axis=0), np.delete(ref, bad, axis=0)))
