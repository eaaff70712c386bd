"""How well does the initial-step proxy of kernels/b2_work.cuh predict the work of a trajectory?  (CPU, oracle.)
Prints the Spearman rank correlation between the proxy max(d1, d2) and the attempted step count of the random
Lorenz sweep, and the list-scheduling makespan (lane model, no SIMT effects) for a few queue orders."""
import heapq, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py
from b200ens import workloads as W
from scipy.stats import spearmanr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
u0, p = W.lorenz_params(N, "random", 0, np.float32)
_, rc, st = oracle_py.solve("lorenz", "Tsit5", u0, p, (0.0, 10.0), np.arange(0, 10.5, 1.0), 0.1, abstol=1e-6, reltol=1e-3, dtype=np.float32)
steps = st[:, 0] + st[:, 1]
def f(u, p):
    return np.stack([p[:, 0] * (u[:, 1] - u[:, 0]), u[:, 0] * (p[:, 1] - u[:, 2]) - u[:, 1], u[:, 0] * u[:, 1] - p[:, 2] * u[:, 2]], 1)
u, pp = u0.astype(np.float64), p.astype(np.float64)
f0 = f(u, pp); sk = 1e-6 + 1e-3 * np.abs(u)
d0 = np.sqrt(((u / sk) ** 2).mean(1)); d1 = np.sqrt(((f0 / sk) ** 2).mean(1))
dt0 = 0.01 * d0 / d1
d2 = np.sqrt((((f(u + dt0[:, None] * f0, pp) - f0) / sk) ** 2).mean(1)) / dt0
proxy = np.maximum(d1, d2)
print("steps mean/min/max", steps.mean(), steps.min(), steps.max(), " spearman(proxy, steps) = %.3f" % spearmanr(proxy, steps)[0])
L = max(1, N * 132608 // 1000000)   # resident lanes scaled to N (148 SMs x 7 CTAs x 128 threads per 1M trajectories)
def makespan(order):
    h = [0] * L
    for s in steps[order]:
        heapq.heappush(h, heapq.heappop(h) + int(s))
    return max(h)
print("ideal", steps.sum() / L)
for name, order in [("caller's order", np.arange(N)), ("proxy descending", np.argsort(-proxy)), ("true LPT", np.argsort(-steps))]:
    print(name, makespan(order))
