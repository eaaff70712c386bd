"""e2e timing of the device-side ensemble summary (b200ens_solve_moments) vs the full-output host path."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens
from b200ens import _lib, workloads as W
N = 1_000_000
for dt in (np.float32, np.float64):
    u0, p = W.lorenz_params(N, "random", 0, dt)
    model = b200ens.build_model(W.lorenz_problem(dt), b200ens.Tsit5())
    o = _lib.default_opts()
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
    SAVEAT = np.arange(0.0, 10.5, 1.0)
    u0p = _lib.pinned_empty(u0.shape, dt); u0p[:] = u0
    pp = _lib.pinned_empty(p.shape, dt); pp[:] = p
    best = 1e9
    for i in range(6):
        t = time.perf_counter(); s, q, cnt, rc, tm = model.solve_moments(o, u0p, pp, SAVEAT); el = (time.perf_counter() - t) * 1e3
        best = min(best, el)
    print(json.dumps({"dtype": np.dtype(dt).name, "summary_wall_ms": round(best, 3), "traj_per_s": N / best * 1e3, "count": cnt,
                      "mean_t10": (s[-1] / cnt).tolist(), **{k: round(v, 3) if isinstance(v, float) else v for k, v in tm.asdict().items()}}))
