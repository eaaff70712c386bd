"""Split-kernel experiments on the 16-species network (device-resident): env B200ENS_SPLIT / B200ENS_MINBLOCKS select the variant."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import workloads as W
sys.argv = [sys.argv[0], "none"]
import importlib.util
spec = importlib.util.spec_from_file_location("bc", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)
N = int(os.environ.get("EXP_N", "200000"))
u0, p = W.net16_params(N)
prob = W.net16_problem()
tag = f"split={os.environ.get('B200ENS_SPLIT','auto')} mb={os.environ.get('B200ENS_MINBLOCKS','auto')}"
which = os.environ.get("EXP_WHICH", "abc")
if "a" in which:
    bc.run(f"[{tag}] Vern7 event saveat101", prob, B.Vern7(), u0, p, np.linspace(0, 10, 101), 0.01, abstol=1e-8, reltol=1e-8, callback=W.net16_callback(), reps=2)
if "b" in which:
    bc.run(f"[{tag}] Vern7 no event, save end", prob, B.Vern7(), u0, p, [10.0], 0.01, abstol=1e-8, reltol=1e-8, reps=2)
if "c" in which:
    bc.run(f"[{tag}] Tsit5 no event, save end", prob, B.Tsit5(), u0, p, [10.0], 0.01, abstol=1e-8, reltol=1e-8, reps=2)
