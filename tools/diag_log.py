import sys, numpy as np
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1: import torch
import b200ens as B
from b200ens import workloads as W
print(B._lib.lib().b200ens_nvrtc_info())
m = B.build_model(W.lorenz_problem(np.float64), B.Tsit5())
print(m.info()); print(str(m.log)[:3000])
