"""e2e pipeline experiment: number of slots x chunk size (pinned buffers, 1M f32 Lorenz)."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens
from b200ens import _lib, workloads as W
N = 1_000_000
dt = np.float32
u0, p = W.lorenz_params(N, "random", 0, dt)
model = b200ens.build_model(W.lorenz_problem(dt), b200ens.Tsit5())
o = _lib.default_opts()
o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
SAVEAT = np.arange(0.0, 10.5, 1.0)
u0p = _lib.pinned_empty(u0.shape, dt); u0p[:] = u0
pp = _lib.pinned_empty(p.shape, dt); pp[:] = p
outp = _lib.pinned_empty((N, 11, 3), dt)
rcp = _lib.pinned_empty((N,), np.int32); stp = _lib.pinned_empty((N, 4), np.int32)
for frac in (16, 24, 32, 48):
  os.environ['B200ENS_FIRST_FRAC'] = str(frac)
  for slots in (4,):
    os.environ["B200ENS_SLOTS"] = str(slots)
    for chunk in (0,):
        if chunk:
            os.environ["B200ENS_CHUNK"] = str(chunk)
        else:
            os.environ.pop("B200ENS_CHUNK", None)   # library default: geometric schedule
        best = 1e9
        for i in range(6):
            t = time.perf_counter(); _, _, _, tm = model.solve(o, u0p, pp, SAVEAT, out=outp, rc=rcp, stats=stp); best = min(best, (time.perf_counter() - t) * 1e3)
        print(json.dumps({"frac": frac, "slots": slots, "chunk": chunk, "wall_ms": round(best, 3), "h2d": round(tm.h2d_ms, 3), "kern": round(tm.kernel_ms, 3), "d2h": round(tm.d2h_ms, 3), "launches": tm.launches}), flush=True)
