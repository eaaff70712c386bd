"""Device-resident throughput of the non-headline BASELINE configs (3, 4, 5) -- numbers for profiles/README.md.
These are parity-test cases, not bench lines (bench.py reports the headline config only)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens as B
from b200ens import _lib, workloads as W


def run(name, prob, alg, u0, p, saveat, dt, adaptive=True, abstol=1e-6, reltol=1e-3, callback=None, reps=3, **optkw):
    npdt = prob.u0.dtype
    tdt = torch.float64 if npdt == np.float64 else torch.float32
    N, n = u0.shape
    t = time.time()
    model = B.build_model(prob, alg, callback)
    tc = time.time() - t
    saveat = np.asarray(saveat, dtype=npdt)
    d_u0, d_p = torch.from_numpy(np.ascontiguousarray(u0, dtype=npdt)).cuda(), torch.from_numpy(np.ascontiguousarray(p, dtype=npdt)).cuda()
    d_save = torch.from_numpy(saveat).cuda()
    d_out = torch.empty((N, len(saveat), n), dtype=tdt, device="cuda")
    d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
    o = _lib.default_opts()
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = int(adaptive), prob.tspan[0], prob.tspan[1], dt, abstol, reltol
    if callback is not None:
        o.interp_points = 10
    for k, v in optkw.items():
        setattr(o, k, v)
    ms = []
    for i in range(reps):
        tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), len(saveat),
                                d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
        ms.append(tm.kernel_ms)
    st = d_st.cpu().numpy(); rc = d_rc.cpu().numpy()
    best = min(ms[1:]) if len(ms) > 1 else ms[0]
    es = np.dtype(npdt).itemsize
    byts = N * ((n + p.shape[1] + len(saveat) * n) * es + 20)
    print(json.dumps({"config": name, "N": N, "dtype": np.dtype(npdt).name, "ms": round(best, 3), "traj_per_s": N / best * 1e3,
                      "steps_per_s": float(st[:, :2].sum()) / best * 1e3, "mean_steps": float(st[:, :2].sum()) / N,
                      "success_frac": float((rc == 1).mean()), "events_mean": float(st[:, 3].mean()),
                      "GBps_algorithmic": byts / best / 1e6, "regs": tm.regs, "grid": tm.grid, "compile_s": round(tc, 1)}), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
if which in ("all", "3"):
    N = int(1_000_000 * scale)
    u0, p = W.robertson_params(N)
    for alg in (B.Rosenbrock23(), B.Rodas5(), B.Rodas5P(), B.FBDF()):
        run(f"cfg3 robertson {alg.name}", W.robertson_problem(), alg, u0, p, W.ROBERTSON_SAVEAT, 1e-6, abstol=1e-8, reltol=1e-6)
if which in ("all", "4"):
    N = int(10_000_000 * scale)
    for dt in (np.float32, np.float64):
        u0, p = W.gbm_params(N, dtype=dt)
        run("cfg4 gbm EM philox", W.gbm_problem(dt), B.EM(), u0, p, [1.0], 1 / 256, adaptive=False, seed=7)
        u0, p = W.lorenz_additive_params(N, dtype=dt)
        for alg in (B.EM(), B.SOSRA()):
            run(f"cfg4 stochastic lorenz {alg.name} philox", W.lorenz_additive_problem(dt), alg, u0, p, [10.0], 1 / 256, adaptive=False, seed=7, maxiters=10**6)
if which in ("all", "5"):
    N = int(1_000_000 * scale)
    u0, p = W.net16_params(N)
    for ns in (101, 1001):
        if ns == 1001 and N > 500_000:
            u0, p = u0[:500_000], p[:500_000]
        run(f"cfg5 net16 Vern7 event saveat{ns}", W.net16_problem(), B.Vern7(), u0, p, np.linspace(0, 10, ns), 0.01, abstol=1e-8, reltol=1e-8,
            callback=W.net16_callback(), reps=2)
