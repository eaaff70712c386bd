"""A/B timing of kernel variants of the headline workload (Lorenz Tsit5, saveat 0:1:10), device-resident.
Usage (GPU box): python tools/ab_headline.py [N] -- variants are B200ENS_DEFINES strings (tools only, see b200ens.cpp)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens
import b200ens.api as api
from b200ens import _lib, workloads as W

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
VARIANTS = os.environ.get("AB_VARIANTS", "|B2_PACK2=0|B2_NORM_DIV=1|B2_PACK2=0,B2_NORM_DIV=1").split("|")
SAVEAT = np.arange(0.0, 10.5, 1.0)
for dt in os.environ.get("AB_DTYPES", "f32,f64").split(","):
    npdt = np.float32 if dt == "f32" else np.float64
    tdt = torch.float32 if dt == "f32" else torch.float64
    for sweep in os.environ.get("AB_SWEEPS", "random").split(","):
        u0, p = W.lorenz_params(N, sweep, 0, npdt)
        d_u0, d_p = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda()
        d_save = torch.from_numpy(SAVEAT.astype(npdt)).cuda()
        d_out = torch.empty((N, 11, 3), dtype=tdt, device="cuda")
        d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
        d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
        ref = None
        for var in VARIANTS:
            if var:
                os.environ["B200ENS_DEFINES"] = var
            else:
                os.environ.pop("B200ENS_DEFINES", None)
            api._model_cache.clear()
            model = b200ens.build_model(W.lorenz_problem(npdt), b200ens.Tsit5())
            o = _lib.default_opts()
            o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
            ms = []
            for i in range(8):
                tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), 11,
                                        d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
                ms.append(tm.kernel_ms)
            st = d_st.cpu().numpy()
            out = d_out.cpu().numpy()
            if ref is None:
                ref = out
            print(json.dumps({"dtype": dt, "sweep": sweep, "defines": var, "ms_best": round(min(ms[2:]), 4),
                              "ms_med": round(float(np.median(ms[2:])), 4), "traj_per_s": N / min(ms[2:]) * 1e3,
                              "steps": int(st[:, :2].sum()), "regs": tm.regs, "grid": tm.grid,
                              "same_bits_as_first": bool(np.array_equal(out, ref, equal_nan=True))}), flush=True)
