"""Ensemble summary (b200ens_solve_moments) with the moments accumulated INSIDE the solve kernel vs the second pass over
out_u: wall time, kernel time and agreement.  python tools/summary_fused_probe.py [cfg5_101|cfg5_1001|lorenz401] [N]"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ens as B
from b200ens import _lib, workloads as W

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5_101"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
o = _lib.default_opts()
if which.startswith("cfg5"):
    ns = int(which.split("_")[1])
    u0, p = W.net16_params(N)
    model = B.build_model(W.net16_problem(), B.Vern7(), W.net16_callback())
    saveat = np.linspace(0, 10, ns)
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol, o.interp_points = 1, 0.0, 10.0, 0.01, 1e-8, 1e-8, 10
    dt = np.float64
else:
    dt = np.float32
    u0, p = W.lorenz_params(N, "random", 0, dt)
    model = B.build_model(W.lorenz_problem(dt), B.Tsit5())
    saveat = np.linspace(0, 10, 401)
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
u0p = _lib.pinned_empty(u0.shape, dt); u0p[:] = u0
pp = _lib.pinned_empty(p.shape, dt); pp[:] = p
res = {}
for mode in ("1", "0"):
    os.environ["B200ENS_FUSE_MOMENTS"] = mode
    best, bk = 1e9, 1e9
    for i in range(3):
        t = time.perf_counter(); s, q, cnt, rc, tm = model.solve_moments(o, u0p, pp, saveat); el = (time.perf_counter() - t) * 1e3
        best = min(best, el); bk = min(bk, tm.kernel_ms)
    res[mode] = (s / cnt, q, cnt)
    print(json.dumps({"config": which, "N": N, "fused": int(mode), "wall_ms": round(best, 2), "kernel_ms_incl_second_pass": round(bk, 2), "count": cnt,
                      "launches": tm.launches, "out_u_bytes_avoided": int(N * saveat.size * u0.shape[1] * np.dtype(dt).itemsize) if mode == "1" else 0}), flush=True)
print(json.dumps({"max_rel_mean_diff": float(np.max(np.abs(res["1"][0] - res["0"][0]) / (np.abs(res["0"][0]) + 1e-300)))}))
