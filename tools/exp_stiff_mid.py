"""Stiff steppers on mid-size systems (n = 5..8): which LU variant (unrolled registers / rolled local memory) wins.
python tools/exp_stiff_mid.py n [N]   (B200ENS_DEFINES=B2_LU_ROLLED=1 forces the rolled variant)"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens as B
from b200ens import _lib
n = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
W_ = [1.3, 0.42, 6.1, 0.17, 2.9, 0.88, 4.4, 0.23, 7.7, 1.9, 0.35, 3.3, 0.61, 5.2, 1.1, 0.7]
def chain(du, u, p, t):
    acc = [0] * n
    for i in range(n - 1):
        fl = p[0] * W_[i] * u[i] - p[1] * W_[i + 1] * u[i + 1]
        acc[i] = acc[i] - fl; acc[i + 1] = acc[i + 1] + fl
    for i in range(n - 2):
        r = p[2] * u[i] * u[i + 1]
        acc[i] = acc[i] - r; acc[i + 1] = acc[i + 1] - r; acc[i + 2] = acc[i + 2] + r
    for i in range(n): du[i] = acc[i]
u0 = np.zeros((N, n)); u0[:, 0] = 1.0
rng = np.random.default_rng(3)
p = np.stack([10.0 ** rng.uniform(0, 3, N), 10.0 ** rng.uniform(-1, 1, N), 10.0 ** rng.uniform(0, 2, N)], axis=1)
prob = B.ODEProblem(chain, u0[0], (0.0, 10.0), p[0])
for alg in (B.Rodas5P(), B.FBDF()):
    t = time.time(); model = B.build_model(prob, alg); tc = time.time() - t
    sv = np.linspace(0, 10, 11)
    d_u0, d_p, d_s = torch.from_numpy(u0).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(sv).cuda()
    d_out = torch.empty((N, 11, n), dtype=torch.float64, device="cuda"); d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
    o = _lib.default_opts(); o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 1e-4, 1e-8, 1e-6
    ms = []
    for _ in range(3):
        tm = model.solve_device(o, 0, 0, N, d_u0.data_ptr(), d_p.data_ptr(), d_s.data_ptr(), 11, d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
        ms.append(tm.kernel_ms)
    st = d_st.cpu().numpy()
    print(json.dumps({"n": n, "alg": alg.name, "ms": round(min(ms[1:]), 3), "steps_per_s": float(st[:, :2].sum()) / min(ms[1:]) * 1e3,
                      "mean_steps": float(st[:, :2].sum()) / N, "ok": float((d_rc == 1).float().mean()), "info": model.info(), "jit_s": round(tc, 1),
                      "defines": os.environ.get("B200ENS_DEFINES", "")}), flush=True)
