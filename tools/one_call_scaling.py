"""The ONE-CALL multi-GPU path: b200ens_solve with device_mask = the first G GPUs, host (pinned) buffers, the whole
ensemble in one blocking call (what a Julia user of EnsembleB200() gets).  Strong scaling: N trajectories total.
python tools/one_call_scaling.py [N] [f32|f64] [random|ordered] [n_save: 11 = saveat 0:1:10 (copy-bound), 1 = end point only
(compute-bound: shows what the dealing of an ordered sweep buys)]  -> JSON lines for G = 1, 2, 4, 8."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import _lib, workloads as W

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dt = np.float32 if (len(sys.argv) < 3 or sys.argv[2] == "f32") else np.float64
sweep = sys.argv[3] if len(sys.argv) > 3 else "random"
NS = int(sys.argv[4]) if len(sys.argv) > 4 else 11
SAVEAT = np.arange(0.0, 10.5, 1.0) if NS == 11 else np.array([10.0])
ndev = _lib.lib().b200ens_device_count()
u0, p = W.lorenz_params(N, sweep, 0, dt)
u0p, pp = _lib.pinned_empty(u0.shape, dt), _lib.pinned_empty(p.shape, dt)
u0p[:], pp[:] = u0, p
out = _lib.pinned_empty((N, len(SAVEAT), 3), dt)
rc = _lib.pinned_empty((N,), np.int32)
model = B.build_model(W.lorenz_problem(dt), B.Tsit5())
ref = None
for G in [g for g in (1, 2, 4, 8) if g <= ndev]:
    for blocks in ((0,) if G == 1 else (0, 1)):
        o = _lib.default_opts()
        o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
        o.device_mask = (1 << G) - 1
        o.shard_blocks = blocks
        best, tm = 1e9, None
        for rep in range(4):
            t = time.perf_counter()
            _, _, _, tm = model.solve(o, u0p, pp, SAVEAT, out=out, rc=rc, want_stats=False)
            el = time.perf_counter() - t
            if rep:
                best = min(best, el)
        if ref is None:
            ref = out.copy()
        print(json.dumps({"gpus": G, "N": N, "dtype": np.dtype(dt).name, "sweep": sweep, "n_save": len(SAVEAT), "shard_blocks": blocks, "wall_ms": round(best * 1e3, 2),
                          "traj_per_s": N / best, "kernel_ms_max": round(tm.kernel_ms, 2), "kernel_ms_min": round(tm.kernel_ms_min, 2),
                          "d2h_gbs": round(out.nbytes / best / 1e9, 1), "bit_identical_to_1gpu": bool(np.array_equal(out, ref, equal_nan=True)),
                          "all_success": bool((rc == 1).all())}), flush=True)
