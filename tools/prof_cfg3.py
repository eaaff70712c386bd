"""Device-resident Robertson / Rodas5P run for ncu captures (config 3): python tools/prof_cfg3.py [N] [alg]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import workloads as W
sys.argv = [sys.argv[0], "none"] + sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
alg = getattr(B, sys.argv[3] if len(sys.argv) > 3 else "Rodas5P")()
u0, p = W.robertson_params(N)
bc.run(f"cfg3 robertson {alg.name}", W.robertson_problem(), alg, u0, p, W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6, reps=3)
