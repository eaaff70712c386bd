"""Host-path (b200ens_solve) timing breakdown and raw PCIe bandwidth on the box."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200ens
from b200ens import _lib, workloads as W

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
# raw PCIe
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, "GB/s", 5 * 256 / 1024 / (time.perf_counter() - t) * 1.073741824)
for dt in (np.float32, np.float64):
    u0, p = W.lorenz_params(N, "random", 0, dt)
    model = b200ens.build_model(W.lorenz_problem(dt), b200ens.Tsit5())
    o = _lib.default_opts()
    o.adaptive, o.t0, o.t1, o.dt, o.abstol, o.reltol = 1, 0.0, 10.0, 0.1, 1e-6, 1e-3
    SAVEAT = np.arange(0.0, 10.5, 1.0)
    u0p = _lib.pinned_empty(u0.shape, dt); u0p[:] = u0
    pp = _lib.pinned_empty(p.shape, dt); pp[:] = p
    outp = _lib.pinned_empty((N, 11, 3), dt)
    rcp = _lib.pinned_empty((N,), np.int32); stp = _lib.pinned_empty((N, 4), np.int32)
    for chunk in (0, 65536, 131072, 262144, 524288, 1 << 20):
        if chunk: os.environ["B200ENS_CHUNK"] = str(chunk)
        else: os.environ.pop("B200ENS_CHUNK", None)
        best = None
        for i in range(5):
            t = time.perf_counter()
            _, _, _, tm = model.solve(o, u0p, pp, SAVEAT, out=outp, rc=rcp, stats=stp)
            el = (time.perf_counter() - t) * 1e3
            if best is None or el < best[0]: best = (el, tm.asdict())
        print(json.dumps({"dtype": np.dtype(dt).name, "chunk": chunk, "wall_ms": round(best[0], 3),
                          **{k: round(v, 3) if isinstance(v, float) else v for k, v in best[1].items()}}))
    # pageable buffers
    os.environ.pop("B200ENS_CHUNK", None)
    outn = np.empty((N, 11, 3), dtype=dt); rcn = np.empty(N, dtype=np.int32); stn = np.empty((N, 4), dtype=np.int32)
    for i in range(4):
        t = time.perf_counter(); model.solve(o, u0, p, SAVEAT, out=outn, rc=rcn, stats=stn); print("pageable wall_ms call", i, round((time.perf_counter() - t) * 1e3, 2))
