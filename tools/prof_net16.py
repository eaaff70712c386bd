"""Minimal device-resident run of config 5 (16-species network, Vern7 + ContinuousCallback) for ncu captures / A-B timing:
python tools/prof_net16.py [N] [n_save]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens as B
from b200ens import _lib, workloads as W

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 101
u0, p = W.net16_params(N)
prob = W.net16_problem()
model = B.build_model(prob, B.Vern7(), W.net16_callback())
saveat = np.linspace(0, 10, ns)
d_u0, d_p, d_save = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(saveat).cuda()
d_out = torch.empty((N, ns, 16), dtype=torch.float64, device="cuda")
d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
o = _lib.default_opts()
o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol, o.interp_points = 1, 0.0, 10.0, 0.01, 1e-8, 1e-8, 10
for i in range(3):
    tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), ns, d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
    print("kernel_ms", round(tm.kernel_ms, 3), "regs", tm.regs, "grid", tm.grid, "info", model.info(), flush=True)
st = d_st.cpu().numpy()
print("steps/traj", st[:, :2].sum() / N, "events/traj", st[:, 3].mean(), "ok", float((d_rc == 1).float().mean()))
