"""Minimal device-resident run of config 4 (GBM, Euler-Maruyama, Philox on the device) for ncu captures / A-B timing:
python tools/prof_sde.py [f32|f64] [N] [gbm|lorenz] [EM|SOSRA]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens as B
from b200ens import _lib, workloads as W

dt = np.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else np.float64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
which = sys.argv[3] if len(sys.argv) > 3 else "gbm"
alg = getattr(B, sys.argv[4])() if len(sys.argv) > 4 else B.EM()
if which == "gbm":
    prob, (u0, p), t1 = W.gbm_problem(dt), W.gbm_params(N, dtype=dt), 1.0
else:
    prob, (u0, p), t1 = W.lorenz_additive_problem(dt), W.lorenz_additive_params(N, dtype=dt), 10.0
n = u0.shape[1]
model = B.build_model(prob, alg)
tdt = torch.float32 if dt == np.float32 else torch.float64
d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
d_save = torch.tensor([t1], dtype=tdt, device="cuda")
d_out = torch.empty((N, 1, n), dtype=tdt, device="cuda")
d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
o = _lib.default_opts()
o.adaptive, o.t0, o.t1, o.dt, o.seed, o.maxiters = 0, 0.0, t1, 1 / 256, 7, 10**6
for i in range(3):
    tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), 1, d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
    print("kernel_ms", round(tm.kernel_ms, 3), "regs", tm.regs, "grid", tm.grid, flush=True)
steps = float(d_st[:, 0].double().sum())
print("steps/s", steps / tm.kernel_ms * 1e3, "mean", float(d_out.double().mean()), "ok", float((d_rc == 1).float().mean()))
