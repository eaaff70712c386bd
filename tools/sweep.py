"""Device-resident timing sweep over scheduling knobs (refill threshold, output staging).
Usage (on the GPU box): python tools/sweep.py [f32|f64] [random|ordered] [N]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens
from b200ens import _lib, workloads as W

dt = sys.argv[1] if len(sys.argv) > 1 else "f32"
sweep = sys.argv[2] if len(sys.argv) > 2 else "random"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
npdt = np.float32 if dt == "f32" else np.float64
tdt = torch.float32 if dt == "f32" else torch.float64
SAVEAT = np.arange(0.0, 10.5, 1.0)
u0, p = W.lorenz_params(N, sweep, 0, npdt)
model = b200ens.build_model(W.lorenz_problem(npdt), b200ens.Tsit5())
if os.environ.get('SWEEP_DIAG'):
    print('DIAG', _lib.lib().b200ens_nvrtc_info(), model.info(), repr(str(model.log)[:300]), flush=True)
d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
d_save = torch.from_numpy(SAVEAT.astype(npdt)).cuda()
d_out = torch.empty((N, 11, 3), dtype=tdt, device="cuda")
d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
REFILLS = [int(x) for x in os.environ.get('SWEEP_REFILL', '1,2,4,8,16,32').split(',')]
STAGES = [int(x) for x in os.environ.get('SWEEP_STAGE', '0,1').split(',')]
for refill in REFILLS:
    for stage in STAGES:
        o = _lib.default_opts()
        o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
        o.refill_threshold, o.stage_outputs = refill, stage
        ms = []
        for i in range(6):
            tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), 11,
                                    d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
            ms.append(tm.kernel_ms)
        st = d_st.cpu().numpy()
        steps = int(st[:, :2].sum())
        best = min(ms[1:])
        print(json.dumps({"dtype": dt, "sweep": sweep, "refill": refill, "stage": stage, "ms": round(best, 4),
                          "traj_per_s": N / best * 1e3, "steps_per_s": steps / best * 1e3, "grid": tm.grid,
                          "smem": tm.smem_bytes, "regs": tm.regs, "minblocks": os.environ.get("B200ENS_MINBLOCKS", "1")}), flush=True)
