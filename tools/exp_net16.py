"""cfg5 decomposition experiments: which part of Vern7 + event + dense saveat costs what (device-resident)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import workloads as W
sys.argv = [sys.argv[0], "none"]
import importlib.util
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)
N = 200000
u0, p = W.net16_params(N)
prob = W.net16_problem()
for dt in (np.float64, np.float32):
    pr = B.ODEProblem(W.net16, prob.u0.astype(dt), prob.tspan, prob.p.astype(dt))
    tol = 1e-8 if dt == np.float64 else 1e-5
    for alg in (B.Vern7(), B.Tsit5()):
        bc.run(f"{alg.name} {np.dtype(dt).name} no event, save end only", pr, alg, u0, p, [10.0], 0.01, abstol=tol, reltol=tol)
    bc.run(f"Vern7 {np.dtype(dt).name} no event, saveat101", pr, B.Vern7(), u0, p, np.linspace(0, 10, 101), 0.01, abstol=tol, reltol=tol)
    bc.run(f"Vern7 {np.dtype(dt).name} event, save end only", pr, B.Vern7(), u0, p, [10.0], 0.01, abstol=tol, reltol=tol, callback=W.net16_callback())
    bc.run(f"Vern7 {np.dtype(dt).name} event interp_points=1, save end only", pr, B.Vern7(), u0, p, [10.0], 0.01, abstol=tol, reltol=tol, callback=W.net16_callback(), interp_points=1)
