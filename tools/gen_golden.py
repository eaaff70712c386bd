"""Generate tests/golden/*.json -- independent truths the CPU oracle is pinned against.

The reference's own tests assert no trajectory value, and Julia is not
available, so these vectors come from INDEPENDENT solvers, not from the reference:
  * lorenz / robertson: scipy DOP853 / Radau at rtol=atol~1e-13 on the problems of
    /root/reference/test/core.jl:22-30 and :39-46
  * linear: closed form 0.5*exp(1.01 t)   (test/core.jl:10-13)
  * philox: Random123 known-answer vectors for Philox4x32-10
  * vdp_mu100: van der Pol mu = 100 (second stiff pin), scipy Radau
  * net16_event: the 16-species network with its bolus ContinuousCallback (config 5), scipy DOP853 + terminal events +
    manual affect + restart
Run:  python tools/gen_golden.py     (needs scipy; output is committed)
"""
import json
import os

import numpy as np
from scipy.integrate import solve_ivp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def lor(t, u, s, r, b):
    return [s * (u[1] - u[0]), u[0] * (r - u[2]) - u[1], u[0] * u[1] - b * u[2]]


def rob(t, u, k1, k2, k3):
    return [-k1 * u[0] + k3 * u[1] * u[2], k1 * u[0] - k2 * u[1] ** 2 - k3 * u[1] * u[2], k2 * u[1] ** 2]


def rob_jac(t, u, k1, k2, k3):
    return [[-k1, k3 * u[2], k3 * u[1]], [k1, -2 * k2 * u[1] - k3 * u[2], -k3 * u[1]], [0, 2 * k2 * u[1], 0]]


ts = np.linspace(0, 1, 11)
sol = solve_ivp(lor, (0, 1), [1.0, 0.0, 0.0], method="DOP853", t_eval=ts, rtol=1e-13, atol=1e-13, args=(10.0, 28.0, 8.0 / 3.0))
json.dump({"problem": "lorenz test/core.jl:22-30", "u0": [1.0, 0.0, 0.0], "p": [10.0, 28.0, 8.0 / 3.0], "t": ts.tolist(),
           "u": sol.y.T.tolist(), "source": "scipy DOP853 rtol=atol=1e-13"}, open(os.path.join(OUT, "lorenz_t1.json"), "w"), indent=1)

ts = np.arange(0, 10.5, 1.0)
rows = []
for rho in (0.5, 14.0, 28.0):
    s = solve_ivp(lor, (0, 10), [1.0, 0.0, 0.0], method="DOP853", t_eval=ts, rtol=1e-13, atol=1e-13, args=(10.0, rho, 8.0 / 3.0))
    rows.append({"p": [10.0, rho, 8.0 / 3.0], "u": s.y.T.tolist()})
json.dump({"problem": "lorenz sweep, tspan (0,10) (BASELINE config 1)", "u0": [1.0, 0.0, 0.0], "t": ts.tolist(), "cases": rows,
           "source": "scipy DOP853 rtol=atol=1e-13 (chaotic cases are only usable at early times)"},
          open(os.path.join(OUT, "lorenz_t10.json"), "w"), indent=1)

# ---- config 1 as a SWEEP: 12 points of the ordered rho-sweep and 12 seeded "random" parameter sets (r (.) (10, 28, 8/3)),
# truth by DOP853 at 1e-13; `reliable` = number of leading save points at which a second DOP853 run at 1e-11 agrees to
# 1e-7 relative (beyond that the chaotic cases are not a truth any more).  Used by tests/test_oracle_golden.py and
# tests/test_spec_arith.py to put a real bound on the global error at the BASELINE tolerances.
ts = np.arange(0.0, 10.5, 1.0)
cases = [(10.0, 56.0 * k / 11, 8.0 / 3.0) for k in range(12)]
rng = np.random.default_rng(123)
for k in range(12):
    r = rng.random(3)
    cases.append((10.0 * r[0], 28.0 * r[1], 8.0 / 3.0 * r[2]))
rows = []
for (sg, rho, beta) in cases:
    A = solve_ivp(lor, (0, 10), [1.0, 0.0, 0.0], method="DOP853", t_eval=ts, rtol=1e-13, atol=1e-13, args=(sg, rho, beta)).y.T
    Bv = solve_ivp(lor, (0, 10), [1.0, 0.0, 0.0], method="DOP853", t_eval=ts, rtol=1e-11, atol=1e-11, args=(sg, rho, beta)).y.T
    agree = np.abs(A - Bv).max(axis=1) <= 1e-7 * (1 + np.abs(A).max(axis=1))
    rows.append({"p": [sg, rho, beta], "u": A.tolist(), "reliable": int(np.argmin(agree)) if not agree.all() else len(ts)})
json.dump({"problem": "lorenz parameter sweep, tspan (0,10), u0 = [1,0,0] (BASELINE config 1)", "u0": [1.0, 0.0, 0.0], "t": ts.tolist(),
           "cases": rows, "source": "scipy DOP853 rtol=atol=1e-13; `reliable` leading save points confirmed by a 1e-11 run"},
          open(os.path.join(OUT, "lorenz_t10_sweep.json"), "w"), indent=1)

ts = 10.0 ** np.arange(-5, 6)
s = solve_ivp(rob, (0, 1e5), [1.0, 0.0, 0.0], method="Radau", jac=rob_jac, t_eval=ts, rtol=1e-12, atol=1e-15, args=(0.04, 3e7, 1e4))
json.dump({"problem": "robertson test/core.jl:39-46", "u0": [1.0, 0.0, 0.0], "p": [0.04, 3e7, 1e4], "t": ts.tolist(),
           "u": s.y.T.tolist(), "source": "scipy Radau rtol=1e-12 atol=1e-15"}, open(os.path.join(OUT, "robertson.json"), "w"), indent=1)

ts = np.linspace(0, 1, 11)
json.dump({"problem": "u'=1.01u test/core.jl:10-13", "u0": [0.5], "p": [1.01], "t": ts.tolist(),
           "u": (0.5 * np.exp(1.01 * ts)).reshape(-1, 1).tolist(), "source": "closed form"},
          open(os.path.join(OUT, "linear.json"), "w"), indent=1)

json.dump({"source": "Random123 kat_vectors, Philox4x32-10 (SURVEY.md B.9)", "vectors": [
    {"ctr": ["00000000"] * 4, "key": ["00000000"] * 2, "out": ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]},
    {"ctr": ["ffffffff"] * 4, "key": ["ffffffff"] * 2, "out": ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]},
    {"ctr": ["243f6a88", "85a308d3", "13198a2e", "03707344"], "key": ["a4093822", "299f31d0"],
     "out": ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]}]}, open(os.path.join(OUT, "philox_kat.json"), "w"), indent=1)
print("golden vectors written to", OUT)

# ---- config 5: 16-species network with the bolus ContinuousCallback, by an INDEPENDENT event integrator:
# scipy DOP853 with a terminal event on X0 - theta (both directions, like ContinuousCallback's default affect_neg! =
# affect!), the bolus applied by hand, integration restarted at the event time.
NET16_W = [1.3, 0.42, 6.1, 0.17, 2.9, 0.88, 4.4, 0.23, 7.7, 1.9, 0.35, 3.3, 0.61, 5.2, 1.1]
NET16_V = [0.7, 2.4, 0.19, 3.8, 0.52, 1.6, 0.11, 8.3, 0.93, 0.27, 4.9, 0.44, 2.2, 0.15, 6.6]
NET16_Z = [0.9, 0.31, 2.7, 0.14, 1.8, 0.66, 3.9, 0.21, 5.5, 0.48, 1.2, 0.12, 2.1, 0.77]


def net16(t, u, p):
    du = np.zeros(16)
    for i in range(15):
        fl = p[0] * NET16_W[i] * u[i] - p[1] * NET16_V[i] * u[i + 1]
        du[i] -= fl
        du[i + 1] += fl
    for i in range(14):
        r = p[2] * NET16_Z[i] * u[i] * u[i + 1]
        du[i] -= r
        du[i + 1] -= r
        du[i + 2] += r
    du[0] -= p[3] * u[0]
    return du


def net16_with_events(p, ts, t1=10.0):
    ev = lambda t, u, p: u[0] - p[4]
    ev.terminal = True
    u = np.zeros(16)
    u[0] = 1.0
    t, times, rows, k = 0.0, [], [], 0
    while True:
        grid = [x for x in ts[k:] if x > t or (x == t and k == 0)]
        s = solve_ivp(net16, (t, t1), u, method="DOP853", rtol=1e-12, atol=1e-14, args=(p,), events=ev, dense_output=True)
        t_end = s.t_events[0][0] if s.status == 1 else t1
        while k < len(ts) and ts[k] <= t_end:      # save points up to (and including) the event time: pre-affect state
            rows.append(s.sol(ts[k]).tolist())
            k += 1
        if s.status != 1:
            break
        times.append(float(t_end))
        u = s.y_events[0][0].copy()
        u[0] += p[5]
        t = t_end
    return times, rows


ts = np.linspace(0.0, 10.0, 11)
cases = []
for p in ([1.0, 0.5, 0.8, 0.3, 0.25, 0.5], [2.2, 0.5, 0.4, 0.3, 0.25, 0.5], [0.45, 0.5, 1.9, 0.3, 0.25, 0.5]):
    times, rows = net16_with_events(p, ts)
    cases.append({"p": p, "event_times": times, "u": rows})
json.dump({"problem": "16-species network + bolus ContinuousCallback (BASELINE config 5; oracle/models.c net16)", "t": ts.tolist(),
           "cases": cases, "source": "scipy DOP853 rtol=1e-12 atol=1e-14, terminal events + manual affect + restart"},
          open(os.path.join(OUT, "net16_event.json"), "w"), indent=1)
print("net16 event golden written:", [len(c["event_times"]) for c in cases], "events")

# ---- a second stiff pin for the Rosenbrock family: van der Pol, mu = 100, t in (0, 50), scipy Radau
def vdp(t, u, mu):
    return [u[1], mu * ((1 - u[0] ** 2) * u[1] - u[0])]


ts = np.linspace(0.0, 50.0, 11)
s = solve_ivp(vdp, (0, 50), [2.0, 0.0], method="Radau", t_eval=ts, rtol=1e-12, atol=1e-14, args=(100.0,))
json.dump({"problem": "van der Pol mu=100", "u0": [2.0, 0.0], "p": [100.0], "t": ts.tolist(), "u": s.y.T.tolist(),
           "source": "scipy Radau rtol=1e-12 atol=1e-14"}, open(os.path.join(OUT, "vdp_mu100.json"), "w"), indent=1)
