"""Stiff steppers on the 16-species network (n = 16): python tools/exp_stiff16.py [N]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "none"] + sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)
import b200ens as B
from b200ens import workloads as W
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
u0, p = W.net16_params(N)
for alg in (B.Vern7(), B.Rodas5P(), B.FBDF()):
    bc.run(f"net16 no event {alg.name}", W.net16_problem(), alg, u0, p, np.linspace(0, 10, 11), 0.01, abstol=1e-8, reltol=1e-8, reps=2)
