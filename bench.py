#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ensemble path (BASELINE.json metric).

metric : trajectories/sec, Lorenz parameter-sweep EnsembleProblem, Tsit5 adaptive
         (abstol 1e-6, reltol 1e-3, dt0 0.1), saveat 0:1:10, 1M trajectories per GPU
         (BASELINE.json configs[1]; random DiffEqGPU-style parameter sweep).
step   : one complete ensemble solve of the resident batch (one kernel launch).
value  : whole-job trajectories/s with u0/p/saveat already resident in HBM (CUDA events on
         the launching stream, L2 flushed between steps, max over ranks).
e2e    : the same solve through the public API / C ABI with HOST (pinned) buffers:
         H2D of u0,p + kernel + D2H of saveat outputs, retcodes and stats inside the timed region.
roofline: FP32 (or FP64) FMA-issue roofline -- 266 algorithmic flops per attempted Tsit5 step
         (68n + 6F + 14, n=3, F=8) x attempted steps / kernel time, against the
         FFMA/DFMA peak measured live by csrc/fma_peak.cu on the same GPU; plus the HBM view.
cpu_baseline / --impl reference: the CPU oracle (restated EnsembleThreads path, OpenMP over
         trajectories, all host cores) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_STEP = 266.0          # Tsit5 attempted step on Lorenz, SURVEY.md 8(d)
SAVEAT = np.arange(0.0, 10.5, 1.0)
TSPAN = (0.0, 10.0)
ABSTOL, RELTOL, DT0 = 1e-6, 1e-3, 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trajectories", type=int, default=1_000_000, help="trajectories per GPU")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--sweep", default="random", choices=["random", "ordered"])
    ap.add_argument("--refill", type=int, default=0)
    ap.add_argument("--stage", type=int, default=-1)
    ap.add_argument("--work-order", type=int, default=-1, help="-1 auto, 0 caller's order, 1 expected-work order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-big", action="store_true", help="skip the informational 10M-trajectory run")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-configs", action="store_true", help="skip the block of the other BASELINE configs (3, 4, 5, Float64)")
    return ap.parse_args()


def bench_config(a):
    """The `config` object of the JSON line: names the workload and nothing arm-specific, so that the GPU arm and the
    `--impl reference` arm print the SAME object (the driver compares them); launch details live under `launch`."""
    return {"workload": workload_name(a), "trajectories_per_gpu": a.trajectories,
            "l2": "GPU arm: flushed between timed steps (256 MiB memset); CPU arm: not applicable"}


def workload_name(a):
    return (f"lorenz_tsit5_adaptive_{a.dtype}_{a.trajectories}traj_per_gpu_{a.sweep}_sweep_saveat0:1:10"
            f"_abstol1e-6_reltol1e-3")


# ------------------------------------------------------------------ CPU oracle timing (baseline / reference arm)
def oracle_throughput(dtype, sweep, n_target, seconds, seed=0):
    """Time the CPU oracle on a bounded sample (first n trajectories of the same seeded workload)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    from b200ens import workloads as W

    oracle_py.build()
    # all host cores this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers,
    # which would silently cripple the reference arm at N > 1
    cores = max(int(oracle_py.lib().orc_max_threads()), len(os.sched_getaffinity(0)))
    u0, p = W.lorenz_params(n_target, sweep, seed, dtype)
    probe = min(n_target, 20000)
    t = time.perf_counter()
    oracle_py.solve("lorenz", "Tsit5", u0[:probe], p[:probe], TSPAN, SAVEAT, DT0, abstol=ABSTOL, reltol=RELTOL, dtype=dtype, nthreads=cores)
    rate = probe / (time.perf_counter() - t)
    n = int(max(probe, min(n_target, rate * seconds)))
    # the whole workload may take less than the time budget: repeat it (about `seconds` of wall time on all cores)
    reps = int(max(1, min(64, np.ceil(seconds * rate / n)))) if n >= n_target else 1
    t = time.perf_counter()
    for _ in range(reps):
        _, rc, st = oracle_py.solve("lorenz", "Tsit5", u0[:n], p[:n], TSPAN, SAVEAT, DT0, abstol=ABSTOL, reltol=RELTOL, dtype=dtype, nthreads=cores)
    el = (time.perf_counter() - t) / reps
    steps = int(st[:, 0].sum() + st[:, 1].sum())
    return {"value": n / el, "unit": "trajectories/s", "cores": cores, "kind": "port",
            "sample": f"first {n} trajectories of the workload x {reps} repeats, CPU oracle (C, OpenMP dynamic, -O2 -mfma -mavx2, "
                      f"{cores} threads), {el:.2f} s per pass ({el * reps * cores:.0f} core-seconds), {steps / el:.3g} attempted steps/s",
            "seconds": el, "n": n}


# ------------------------------------------------------------------ the other BASELINE.json configs (one JSON block)
# Algorithmic flops per ATTEMPTED step (FMA = 2, add/mul/div/max = 1; RNG integer work not counted):
#   Tsit5:  68 n + 6 F + 14                      Lorenz n = 3, F = 8                       -> 266
#   Vern7: 118 n + 10 F + 14                     16-species network n = 16, F = 190        -> 3802
#   Rosenbrock23 (n = 3, Robertson F = 13, J = 8):  W 9 + LU 18 + 3 solves x 18 + 3 F + ~60 of vector sums / norm -> 190
#   Rodas5 / Rodas5P (s = 8):  J 8 + W 9 + LU 18 + 8 solves x 18 + 8 F + a-sums 96 + C-sums 147 + 42 + norm/controller 29 -> 600
#   FBDF (n = 3, typical order k = 4, 2 Newton iterations): J 8 + W 9 + LU 18 + 2 x (F 13 + residual 9 + solve 18 + norm 12)
#       + Lagrange weights and re-sampling ~220 + divided differences and their norms ~160 + f(u_new) 13 + control ~20 -> 550 (rough)
#   EM: 2n + 2n + F + G + n (dW = sqrt(dt) z)     GBM n = 1: 7;  stochastic Lorenz n = 3, F = 8, G = 0: 23
#   SOSRA (3 drift stages, additive noise):  3 F + stage sums 2 x (4n + 2n) + update 12 n + chi2 3n + 6n (dW, dZ) -> 111 (n = 3)
CFG_FLOPS = {"Tsit5": 266.0, "Vern7_net16": 3802.0, "Rosenbrock23": 190.0, "Rodas5": 600.0, "Rodas5P": 600.0, "FBDF": 550.0,
             "EM_gbm": 7.0, "EM_lorenz": 23.0, "SOSRA_lorenz": 111.0}


def run_configs(dev, stream, peaks_tf, hbm_peak, cpu_seconds, cores):
    """Device-resident runs of BASELINE.json configs 3, 4, 5 (+ the Float64 headline): kernel ms (best of `reps`, CUDA
    events inside b200ens_solve_device), steps/s, the roofline that bounds each (FP32/FP64 FMA issue with the algorithmic
    flops above, or HBM with the algorithmic bytes for the dense-saveat run) and a CPU-arm number from the oracle on a
    bounded subset of the same seeded inputs.  Parity of every one of these launch shapes is in tests/."""
    import torch

    import b200ens as B
    from b200ens import _lib, workloads as W

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    out = {}

    def run(name, prob, alg, u0, p, saveat, dt, flops, omodel, adaptive=True, abstol=1e-6, reltol=1e-3, callback=None, reps=3,
            cpu_n=20000, bound=None, seed=0, maxiters=None):
        npdt = prob.u0.dtype
        f64 = npdt == np.float64
        tdt = torch.float64 if f64 else torch.float32
        N, n = u0.shape
        t = time.perf_counter()
        model = B.build_model(prob, alg, callback)
        jit_s = time.perf_counter() - t
        saveat = np.asarray(saveat, dtype=npdt)
        d_u0 = torch.from_numpy(np.ascontiguousarray(u0, dtype=npdt)).cuda()
        d_p = torch.from_numpy(np.ascontiguousarray(p, dtype=npdt)).cuda()
        d_save = torch.from_numpy(saveat).cuda()
        d_out = torch.empty((N, len(saveat), n), dtype=tdt, device="cuda")
        d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
        d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
        o = _lib.default_opts()
        o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = int(adaptive), prob.tspan[0], prob.tspan[1], dt, abstol, reltol
        o.seed = seed
        if maxiters:
            o.maxiters = maxiters
        if callback is not None:
            o.interp_points = 10
        ms = []
        for _ in range(reps):
            tm = model.solve_device(o, dev, stream.cuda_stream, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(), len(saveat),
                                    d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr())
            ms.append(tm.kernel_ms)
        st = d_st.cpu().numpy()
        rc = d_rc.cpu().numpy()
        best = min(ms[1:]) if len(ms) > 1 else ms[0]
        steps = float(st[:, :2].sum())
        es = np.dtype(npdt).itemsize
        alg_bytes = N * ((n + p.shape[1] + len(saveat) * n) * es + 4 + 16)
        tf = flops * steps / (best * 1e-3) / 1e12
        peak = peaks_tf["f64" if f64 else "f32"]
        gbs = alg_bytes / (best * 1e-3) / 1e9
        if bound is None:
            bound = "fma_fp64" if f64 else "fma_fp32"
        roof = ({"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "algorithmic_bytes": alg_bytes}
                if bound == "hbm" else
                {"bound": bound, "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "flops_per_step": flops})
        # CPU arm: the oracle on the first cpu_n trajectories of the same inputs, all host threads
        cpu = None
        if cpu_seconds > 0:
            k = min(N, cpu_n)
            kw = dict(abstol=abstol, reltol=reltol, adaptive=adaptive, dtype=npdt, nthreads=cores, seed=seed, want_stats=False)
            if maxiters:
                kw["maxiters"] = maxiters
            if callback is not None:
                kw["event"] = True
            t = time.perf_counter()
            oracle_py.solve(omodel, alg.name, u0[:k], p[:k], prob.tspan, saveat, dt, **kw)
            el = time.perf_counter() - t
            cpu = {"value": k / el, "unit": "trajectories/s", "cores": cores, "kind": "port", "sample": f"first {k} trajectories, {el:.2f} s"}
        del d_out
        torch.cuda.empty_cache()
        out[name] = {"N": N, "dtype": np.dtype(npdt).name, "n_save": int(len(saveat)), "kernel_ms": round(best, 4),
                     "traj_per_s": N / best * 1e3, "steps_per_s": steps / best * 1e3, "mean_steps": steps / N,
                     "success_frac": float(np.isin(rc, (1, 2)).mean()), "events_mean": float(st[:, 3].mean()),
                     "roofline": roof, "hbm_gbs_algorithmic": gbs, "regs": tm.regs, "launches": tm.launches, "jit_s": round(jit_s, 2),
                     "cpu_baseline": cpu, "speedup_vs_cpu_port": (N / best * 1e3) / cpu["value"] if cpu else None}

    SAVE11 = np.arange(0.0, 10.5, 1.0)
    # configs[1] in Float64 (config 1 is a Float64 reference run)
    u0, p = W.lorenz_params(1_000_000, "random", 0, np.float64)
    run("cfg2_lorenz_tsit5_f64_1M", W.lorenz_problem(np.float64, TSPAN), B.Tsit5(), u0, p, SAVE11, DT0, CFG_FLOPS["Tsit5"], "lorenz", cpu_n=1000000)
    # config 3: Robertson, Rosenbrock23 / Rodas5 / Rodas5P with the analytic Jacobian, 1M trajectories
    u0, p = W.robertson_params(1_000_000)
    for alg in (B.Rosenbrock23(), B.Rodas5(), B.Rodas5P(), B.FBDF()):
        run(f"cfg3_robertson_{alg.name}_f64_1M", W.robertson_problem(), alg, u0, p, W.ROBERTSON_SAVEAT, 1e-6, CFG_FLOPS[alg.name], "robertson",
            abstol=1e-8, reltol=1e-6, cpu_n=200000)
    # config 4: GBM (EM) and stochastic Lorenz (EM, SOSRA), 10M paths, Philox on the device
    for dt_ in (np.float32, np.float64):
        tag = np.dtype(dt_).name.replace("float", "f")
        u0, p = W.gbm_params(10_000_000, dtype=dt_)
        run(f"cfg4_gbm_EM_{tag}_10M", W.gbm_problem(dt_), B.EM(), u0, p, [1.0], 1 / 256, CFG_FLOPS["EM_gbm"], "gbm", adaptive=False, seed=7,
            cpu_n=1000000, reps=2)
    u0, p = W.lorenz_additive_params(10_000_000, dtype=np.float32)
    for alg, key in ((B.EM(), "EM_lorenz"), (B.SOSRA(), "SOSRA_lorenz")):
        run(f"cfg4_stochastic_lorenz_{alg.name}_f32_10M", W.lorenz_additive_problem(np.float32), alg, u0, p, [10.0], 1 / 256, CFG_FLOPS[key],
            "lorenz_additive", adaptive=False, seed=7, maxiters=10**6, cpu_n=100000, reps=2)
    del u0, p
    # config 5: 16-species network, Vern7 + ContinuousCallback, dense saveat
    u0, p = W.net16_params(1_000_000)
    run("cfg5_net16_vern7_event_f64_1M_saveat101", W.net16_problem(), B.Vern7(), u0, p, np.linspace(0, 10, 101), 0.01, CFG_FLOPS["Vern7_net16"],
        "net16", abstol=1e-8, reltol=1e-8, callback=W.net16_callback(), reps=2, cpu_n=50000)
    run("cfg5_net16_vern7_event_f64_500k_saveat1001", W.net16_problem(), B.Vern7(), u0[:500_000], p[:500_000], np.linspace(0, 10, 1001), 0.01,
        CFG_FLOPS["Vern7_net16"], "net16", abstol=1e-8, reltol=1e-8, callback=W.net16_callback(), reps=2, cpu_n=10000, bound="hbm")
    return out


def run_reference(a, rank, world):
    if rank != 0:
        return
    dtype = np.float32 if a.dtype == "f32" else np.float64
    vals = []
    base = None
    for i in range(a.warmup + a.steps):
        r = oracle_throughput(dtype, a.sweep, a.trajectories, max(1.0, a.cpu_seconds / max(1, a.steps)))
        if i >= a.warmup:
            vals.append(r)
        base = r
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals])) * 1e3
    line = {
        "impl": "reference", "metric": "trajectories/sec (Lorenz Tsit5, 1M traj)", "value": v, "unit": "trajectories/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": bench_config(a), "note": "restated CPU baseline (oracle port of the EnsembleThreads path), not Julia",
        "cpu_baseline": {"value": v, "unit": "trajectories/s", "cores": base["cores"], "kind": "port", "sample": base["sample"]},
        "e2e": {"value": v, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def numa_bind(dev):
    """Best effort: run this rank (and first-touch its pinned host buffers) on the NUMA node its GPU hangs off, so
    that H2D/D2H traffic of the ranks does not cross the socket interconnect.  Returns a short description."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return f"{bdf}: no NUMA affinity reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{bdf}: node {node}, {len(allowed)} cpus"
        return f"{bdf}: node {node} has no allowed cpus"
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__})"


def fma_peak(device, f64):
    lib = ctypes.CDLL(os.path.join(ROOT, "differentialequations.jl_b200", "csrc", "libb200peak.so"))
    lib.b200_fma_peak_tflops.restype = ctypes.c_double
    lib.b200_fma_peak_tflops.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return float(lib.b200_fma_peak_tflops(device, int(f64), 4096 if not f64 else 2048))


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return

    import torch
    import torch.distributed as dist

    import b200ens
    from b200ens import _lib, workloads as W

    dev = local_rank
    torch.cuda.set_device(dev)
    numa = numa_bind(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    f64 = a.dtype == "f64"
    npdt = np.float64 if f64 else np.float32
    tdt = torch.float64 if f64 else torch.float32
    N = a.trajectories
    n_save = len(SAVEAT)

    # weak scaling: every rank owns N trajectories of the seeded sweep (rank r = shard r of a world*N ensemble)
    t_pf = time.perf_counter()
    u0_h, p_h = W.lorenz_params(N, a.sweep, seed=rank, dtype=npdt)
    prob_func_ms = (time.perf_counter() - t_pf) * 1e3   # the vectorised prob_func (EnsembleProblem(prob; u0s, ps)): host numpy, outside e2e
    t_jit = time.perf_counter()
    model = b200ens.build_model(W.lorenz_problem(npdt, TSPAN), b200ens.Tsit5())
    jit_ms = (time.perf_counter() - t_jit) * 1e3         # trace + emit + NVRTC, or a hit in the on-disk cubin cache
    o = _lib.default_opts()
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, TSPAN[0], TSPAN[1], DT0, ABSTOL, RELTOL
    o.refill_threshold, o.stage_outputs, o.work_order = a.refill, a.stage, a.work_order
    o.device_mask = 1 << dev
    o.traj_offset = rank * N

    # ---- device-resident leg (value)
    d_u0 = torch.from_numpy(u0_h).cuda()
    d_p = torch.from_numpy(p_h).cuda()
    d_save = torch.from_numpy(SAVEAT.astype(npdt)).cuda()
    d_out = torch.empty((N, n_save, 3), dtype=tdt, device="cuda")
    d_rc = torch.zeros(N, dtype=torch.int32, device="cuda")
    d_st = torch.zeros((N, 4), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def step_device():
        return model.solve_device(o, dev, stream.cuda_stream, N, d_u0.data_ptr(), d_p.data_ptr(), d_save.data_ptr(),
                                  n_save, d_out.data_ptr(), d_rc.data_ptr(), d_st.data_ptr(), timed=False)

    sampler = ClockSampler(dev)   # samples every 100 ms from the warm-up to the end of the timed legs
    sampler.start()
    for _ in range(a.warmup):
        step_device()
    torch.cuda.synchronize()
    # kernels per device-resident step: 1 ensemble kernel (+ 2 of the expected-work ordering pre-pass when it is on)
    launches_per_step = int(model.solve_device(o, dev, stream.cuda_stream, N, d_u0.data_ptr(), d_p.data_ptr(),
                                               d_save.data_ptr(), n_save, d_out.data_ptr(), d_rc.data_ptr(),
                                               d_st.data_ptr(), timed=True).launches)
    peak_tf = fma_peak(dev, f64)  # measured FMA roofline denominator (also warms the clocks)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for s, e in evs:
        flush.zero_()          # L2 flush between timed steps (256 MiB > 126 MB L2), outside the events
        s.record(stream)
        step_device()
        e.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    kernel_ms = [s.elapsed_time(e) for s, e in evs]
    total_ms = float(sum(kernel_ms))
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / a.steps
    value = world * N / (ms_per_step * 1e-3)

    st = d_st.cpu().numpy()
    rc = d_rc.cpu().numpy()
    attempted = int(st[:, 0].sum() + st[:, 1].sum())
    ok = bool((rc == 1).all())
    my_ms = float(np.mean(kernel_ms))
    achieved_tf = FLOPS_PER_STEP * attempted / (my_ms * 1e-3) / 1e12
    es = 8 if f64 else 4
    alg_bytes = N * ((3 + 3 + n_save * 3) * es + 4 + 16)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(a.dtype)
    except Exception:
        pass
    roofline = {"bound": "fma_fp64" if f64 else "fma_fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf > 0 else None, "traffic": traffic,
                "peak_source": "measured live by csrc/fma_peak.cu (FFMA/DFMA register-operand chains); nominal "
                               + ("37.2" if f64 else "74.4") + " TFLOP/s",
                "flops_per_step": FLOPS_PER_STEP, "attempted_steps_per_launch": attempted,
                "steps_per_s": attempted / (my_ms * 1e-3), "kernel_ms": my_ms}
    roofline_hbm = {"bound": "hbm", "achieved": alg_bytes / (my_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": alg_bytes / (my_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": alg_bytes,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"}

    # ---- informational: the same kernel on 10M trajectories (configs[1] spans 1M-10M): the drain tail of the persistent
    # kernel is a fixed ~100 iterations, so the larger ensemble shows the tail-free rate.  Not the headline value.
    big = None
    if world == 1 and N == 1_000_000 and not a.no_big:
        try:
            NB = 10_000_000
            ub, pb = W.lorenz_params(NB, a.sweep, seed=rank, dtype=npdt)
            d_ub, d_pb = torch.from_numpy(ub).cuda(), torch.from_numpy(pb).cuda()
            d_ob = torch.empty((NB, n_save, 3), dtype=tdt, device="cuda")
            d_rb = torch.zeros(NB, dtype=torch.int32, device="cuda")
            d_sb = torch.zeros((NB, 4), dtype=torch.int32, device="cuda")
            msb = []
            for _ in range(3):
                tmb = model.solve_device(o, dev, stream.cuda_stream, NB, d_ub.data_ptr(), d_pb.data_ptr(), d_save.data_ptr(),
                                         n_save, d_ob.data_ptr(), d_rb.data_ptr(), d_sb.data_ptr(), timed=True)
                msb.append(tmb.kernel_ms)
            big = {"trajectories": NB, "ms": min(msb[1:]), "value": NB / (min(msb[1:]) * 1e-3), "all_success": bool((d_rb == 1).all().item())}
            del d_ub, d_pb, d_ob, d_rb, d_sb
        except Exception as e:  # noqa: BLE001 -- informational only
            big = {"error": str(e)[:200]}

    # ---- end-to-end leg through the public C-ABI call with pinned HOST buffers
    u0_pin = _lib.pinned_empty(u0_h.shape, npdt)
    p_pin = _lib.pinned_empty(p_h.shape, npdt)
    out_pin = _lib.pinned_empty((N, n_save, 3), npdt)
    rc_pin = _lib.pinned_empty((N,), np.int32)
    u0_pin[:] = u0_h
    p_pin[:] = p_h
    e2e_launches = 0
    # per-trajectory step statistics (sol.stats: naccept / nreject / nf, 16 B per trajectory) are optional diagnostics of the
    # ABI (stats = NULL) and are not requested here: the result of the step is the saveat array and the retcodes
    for _ in range(2):
        model.solve(o, u0_pin, p_pin, SAVEAT, out=out_pin, rc=rc_pin, want_stats=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        _, rc2, _, tm = model.solve(o, u0_pin, p_pin, SAVEAT, out=out_pin, rc=rc_pin, want_stats=False)
        e2e_launches += tm.launches
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * N * a.steps / e2e_s
    # ---- the same ensemble when the caller wants per-save-point statistics instead of the trajectories
    # (b200ens_solve_moments: EnsembleSummary / timestep_meanvar reduced on the device): 4 B of retcode per trajectory and
    # 2 x 33 doubles come back instead of 132 B per trajectory, so this path is NOT bound by the host link
    e2e_summary = None
    try:
        for _ in range(2):
            model.solve_moments(o, u0_pin, p_pin, SAVEAT, rc=rc_pin)
        if world > 1:
            dist.barrier()
        ts0 = time.perf_counter()
        for _ in range(a.steps):
            s_sum, s_sq, s_cnt, _, _ = model.solve_moments(o, u0_pin, p_pin, SAVEAT, rc=rc_pin)
        es_s = time.perf_counter() - ts0
        if world > 1:
            tt = torch.tensor([es_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            es_s = float(tt.item())
        mean_dev = s_sum / max(s_cnt, 1)
        mean_host = out_pin.astype(np.float64).mean(axis=0)
        e2e_summary = {"value": world * N * a.steps / es_s, "unit": "trajectories/s", "ms_per_step": es_s / a.steps * 1e3,
                       "h2d_bytes_per_step": int(u0_h.nbytes + p_h.nbytes), "d2h_bytes_per_step": int(4 * N + 2 * n_save * 3 * 8 + 8),
                       "count": int(s_cnt), "max_abs_mean_diff_vs_full_output": float(np.abs(mean_dev - mean_host).max())}
    except Exception as e:  # noqa: BLE001
        e2e_summary = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---- informational: the same ensemble with solve(...; save_idxs = [3]) -- only z(t) is saved: 44 + 4 B per trajectory
    # come back instead of 132 + 4, the host link stops being the bound.  Not the headline (another output contract).
    e2e_save_idxs = None
    try:
        if world == 1:
            model_z = b200ens.build_model(W.lorenz_problem(npdt, TSPAN), b200ens.Tsit5(), save_idxs=[2])
            out_z = _lib.pinned_empty((N, n_save, 1), npdt)
            for _ in range(2):
                model_z.solve(o, u0_pin, p_pin, SAVEAT, out=out_z, rc=rc_pin, want_stats=False)
            tz0 = time.perf_counter()
            for _ in range(a.steps):
                model_z.solve(o, u0_pin, p_pin, SAVEAT, out=out_z, rc=rc_pin, want_stats=False)
            ez_s = time.perf_counter() - tz0
            e2e_save_idxs = {"value": N * a.steps / ez_s, "unit": "trajectories/s", "ms_per_step": ez_s / a.steps * 1e3,
                             "save_idxs": [3], "d2h_bytes_per_step": int(out_z.nbytes + rc_pin.nbytes),
                             "identical_to_column_of_full_output": bool(np.array_equal(out_z[:, :, 0], out_pin[:, :, 2]))}
    except Exception as e:  # noqa: BLE001
        e2e_save_idxs = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    # ---- the host-link ceiling of THIS box at this GPU count: the same bytes, copies only (no kernel), all ranks at once.
    # e2e cannot beat it; `frac_of_host_ceiling` says how close the pipelined solve gets.
    host_ceiling = None
    try:
        t_out = torch.from_numpy(out_pin.reshape(-1))
        t_rc = torch.from_numpy(rc_pin)
        t_u0, t_p = torch.from_numpy(u0_pin), torch.from_numpy(p_pin)
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            tc = time.perf_counter()
            d_u0.copy_(t_u0, non_blocking=True)
            d_p.copy_(t_p, non_blocking=True)
            t_out.copy_(d_out.reshape(-1), non_blocking=True)
            t_rc.copy_(d_rc, non_blocking=True)
            torch.cuda.synchronize()
            el = time.perf_counter() - tc
            if world > 1:
                tt = torch.tensor([el], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                el = float(tt.item())
            best = min(best, el)
        moved = int(u0_h.nbytes + p_h.nbytes + out_pin.nbytes + 4 * N)
        host_ceiling = {"copy_only_ms": best * 1e3, "aggregate_gbs": world * moved / best / 1e9,
                        "traj_per_s_at_ceiling": world * N / best, "pinned": bool(t_out.is_pinned())}
    except Exception as e:  # noqa: BLE001
        host_ceiling = {"error": str(e)[:200]}
    clocks = sampler.stop()
    same = bool(np.array_equal(out_pin[:1000], d_out[:1000].cpu().numpy(), equal_nan=True))
    h2d = int(u0_h.nbytes + p_h.nbytes + SAVEAT.astype(npdt).nbytes)
    d2h = int(out_pin.nbytes + 4 * N)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = oracle_throughput(npdt, a.sweep, N, a.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    configs = None
    if rank == 0 and world == 1 and not a.no_configs:
        t_cfg = time.perf_counter()
        try:
            peaks_tf = {"f32": peak_tf if not f64 else fma_peak(dev, False), "f64": peak_tf if f64 else fma_peak(dev, True)}
            configs = run_configs(dev, stream, peaks_tf, hbm_peak, 0.0 if a.no_cpu_baseline else a.cpu_seconds,
                                  cpu["cores"] if cpu else len(os.sched_getaffinity(0)))
            configs["_seconds"] = round(time.perf_counter() - t_cfg, 1)
        except Exception as e:  # noqa: BLE001 -- the headline line must still print
            configs = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

    if rank == 0:
        line = {
            "metric": "trajectories/sec (Lorenz Tsit5, 1M traj)", "value": value, "unit": "trajectories/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": bench_config(a),
            "launch": {"parallelism": f"trajectory ranges sharded over {world} GPU(s), no collective",
                       "refill_threshold": a.refill, "stage_outputs": a.stage, "work_order": a.work_order,
                       "kernels_per_step": launches_per_step, "all_success": ok, "same_kernel_at_10M_trajectories": big, "numa": numa,
                       "regs": model.info()["regs"]},
            "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "trajectories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / a.steps * 1e3, "matches_device_leg": same, "host_ceiling": host_ceiling,
                    "summary_path": e2e_summary, "save_idxs_path": e2e_save_idxs,
                    "outside_the_timed_region": {"prob_func_vectorised_ms": prob_func_ms, "model_build_ms": jit_ms,
                                                 "note": "prob_func as a parameter matrix (numpy, once per ensemble) and the one-time trace + NVRTC JIT (cubins are cached on disk)"},
                    "frac_of_host_ceiling": (e2e_val / host_ceiling["traj_per_s_at_ceiling"]) if host_ceiling and "traj_per_s_at_ceiling" in host_ceiling else None},
            "gpu_launches": a.steps * launches_per_step + e2e_launches, "clocks": clocks,
            "trajectory_steps_per_s": world * attempted / (ms_per_step * 1e-3),
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
