"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_b200ens.so")

ALG = {"Tsit5": 1, "Vern7": 2, "Rosenbrock23": 3, "Rodas5": 4, "Rodas5P": 5, "EM": 6, "SOSRA": 7, "Rodas4": 8, "SRIW1": 9, "FBDF": 10}
RC_SUCCESS, RC_TERMINATED, RC_MAXITERS, RC_DTLESSTHANMIN, RC_UNSTABLE, RC_DTNAN = 1, 2, 3, 4, 5, 6


class Stats(C.Structure):
    _fields_ = [("naccept", C.c_int32), ("nreject", C.c_int32), ("nf", C.c_int32), ("nevents", C.c_int32)]


class Opts(C.Structure):
    _fields_ = [
        ("alg", C.c_int32), ("n_state", C.c_int32), ("n_param", C.c_int32), ("adaptive", C.c_int32),
        ("t0", C.c_double), ("t1", C.c_double), ("dt", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double), ("qmin", C.c_double), ("qmax", C.c_double),
        ("gamma", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("qoldinit", C.c_double),
        ("maxiters", C.c_int64), ("n_save", C.c_int32), ("noise_injected", C.c_int32), ("seed", C.c_uint64), ("traj_offset", C.c_uint64),
        ("has_event", C.c_int32), ("event_terminate", C.c_int32), ("interp_points", C.c_int32),
        ("save_tstops", C.c_int32),
        ("rhs", C.c_void_p), ("jac", C.c_void_p), ("tgrad", C.c_void_p), ("noise", C.c_void_p),
        ("cond", C.c_void_p), ("affect", C.c_void_p), ("dcond", C.c_void_p), ("daffect", C.c_void_p),
        ("devent_terminate", C.c_int32), ("pad_", C.c_int32),
        ("abstol_vec", C.POINTER(C.c_double)), ("reltol_vec", C.POINTER(C.c_double)),
        ("vcond", C.c_void_p), ("vaffect", C.c_void_p), ("ncond", C.c_int32), ("vterm_mask", C.c_uint32),
        ("mass", C.POINTER(C.c_double)),
        ("every_t", C.c_void_p), ("save_everystep", C.c_int32), ("sde_adaptive", C.c_int32),
        ("noise_stream_len", C.c_int64), ("spec_arith", C.c_int32), ("n_tstops", C.c_int32),
        ("event_dir", C.c_int32), ("pad2_", C.c_int32), ("affect_neg", C.c_void_p),
        ("tstops", C.POINTER(C.c_double)),
    ]


def _arch_flags():
    """-mfma -mavx2 when this host has them (fma() becomes one instruction: 2.7x faster, same bits), else nothing."""
    try:
        flags = open("/proc/cpuinfo").read()
        return "-mfma -mavx2" if (" fma " in flags or " fma\n" in flags) and " avx2" in flags else ""
    except OSError:
        return ""


def build(force=False):
    want = _arch_flags()
    stamp = os.path.join(HERE, "liboracle_b200ens.flags")
    have = open(stamp).read().strip() if os.path.exists(stamp) else None
    if force or not os.path.exists(LIB) or have != want or any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB)
        for f in ("oracle.c", "oracle_impl.inc", "models.c", "oracle.h", "tableaus_gen.h", "Makefile")
    ):
        if os.path.exists(LIB):
            os.remove(LIB)   # the stamp, not make's timestamps, decides when the flags changed
        subprocess.check_call(["make", "-C", HERE, "-s", f"ARCHFLAGS={want}"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_model_fn.restype = C.c_void_p
        _lib.orc_model_fn.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        _lib.orc_fastpow.restype = C.c_float
        _lib.orc_fastpow.argtypes = [C.c_float, C.c_float]
        for name in ("orc_solve_f64", "orc_solve_f32"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def model_fn(model, which, f64):
    return lib().orc_model_fn(model.encode(), which.encode(), 1 if f64 else 0)


def compile_host_model(src, tag, workdir):
    """Compile emitted model code (C++ wrapper from codegen.host_wrapper_source) with the oracle's
    floating-point flags (no contraction) so that the oracle runs the SAME expression tree as the GPU."""
    import hashlib

    os.makedirs(workdir, exist_ok=True)
    h = hashlib.sha1((src + _arch_flags()).encode()).hexdigest()[:12]
    cpath = os.path.join(workdir, f"model_{tag}_{h}.cpp")
    so = os.path.join(workdir, f"model_{tag}_{h}.so")
    if not os.path.exists(so):
        open(cpath, "w").write(src)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                               *_arch_flags().split(), "-o", so, cpath, "-lm"])
    return C.CDLL(so)


def fns_from_host_model(dll, f64):
    """dict(rhs=ptr, jac=ptr, ...) for solve(fns=...) from a compile_host_model() library."""
    suf = "f64" if f64 else "f32"
    out = {}
    for key, nm in (("rhs", "b2_rhs"), ("jac", "b2_jac"), ("tgrad", "b2_tgrad"), ("noise", "b2_noise"),
                    ("cond", "b2_condition"), ("affect", "b2_affect"), ("dcond", "b2_dcondition"),
                    ("daffect", "b2_daffect"), ("vcond", "b2_vcondition"), ("vaffect", "b2_vaffect"), ("affect_neg", "b2_affect_neg")):
        try:
            out[key] = C.cast(getattr(dll, f"{nm}_{suf}"), C.c_void_p).value
        except AttributeError:
            out[key] = None
    return out


def solve(model, alg, u0, p, tspan, saveat, dt, abstol=1e-6, reltol=1e-3, adaptive=True, dtype=np.float64,
          maxiters=100000, dW=None, seed=0, event=False, terminate=False, interp_points=10, nthreads=0,
          fns=None, want_stats=True, save_tstops=None, traj_offset=0, devent=False, dterminate=False, ncond=0, vterm_mask=0, mass_matrix=None, save_everystep=0, sde_adaptive=False, spec_arith=False, tstops=None, event_dir=0, terminate_neg=False, **ctl):
    """Run the oracle.  model: built-in name, or fns = dict(rhs=ptr, jac=ptr, ...)."""
    L = lib()
    f64 = np.dtype(dtype) == np.float64
    u0 = np.ascontiguousarray(u0, dtype=dtype)
    p = np.ascontiguousarray(p, dtype=dtype)
    N, n = u0.shape
    m = p.shape[1]
    saveat = np.ascontiguousarray(saveat, dtype=dtype)
    o = Opts()
    o.alg = ALG[alg]
    o.n_state, o.n_param, o.adaptive = n, m, int(adaptive)
    keep = []   # per-component tolerances / mass matrix: arrays must outlive the call
    if mass_matrix is not None:
        mm = np.ascontiguousarray(mass_matrix, dtype=np.float64); assert mm.shape == (n, n)
        keep.append(mm); o.mass = mm.ctypes.data_as(C.POINTER(C.c_double))
    if np.ndim(abstol) > 0:
        av = np.ascontiguousarray(abstol, dtype=np.float64); assert av.shape == (n,)
        keep.append(av); o.abstol_vec = av.ctypes.data_as(C.POINTER(C.c_double)); abstol = float(av[0])
    if np.ndim(reltol) > 0:
        rv = np.ascontiguousarray(reltol, dtype=np.float64); assert rv.shape == (n,)
        keep.append(rv); o.reltol_vec = rv.ctypes.data_as(C.POINTER(C.c_double)); reltol = float(rv[0])
    o.t0, o.t1, o.dt, o.abstol, o.reltol = tspan[0], tspan[1], dt, abstol, reltol
    for k in ("dtmin", "dtmax", "qmin", "qmax", "gamma", "beta1", "beta2", "qoldinit"):
        setattr(o, k, ctl.get(k, -1.0))
    o.maxiters = maxiters
    o.n_save = len(saveat)
    o.noise_injected = 0 if dW is None else 1
    o.seed = seed
    o.traj_offset = traj_offset
    o.has_event, o.event_terminate, o.interp_points = int(event), int(bool(terminate)) | (4 if terminate_neg else 0), interp_points
    o.event_dir = int(event_dir)   # 0 both directions, +1 upcrossings only, -1 downcrossings only
    if save_tstops is None:   # everything interpolates, except FBDF on a mass-matrix problem (its Hermite output needs u')
        save_tstops = mass_matrix is not None and alg == "FBDF"
    o.save_tstops = int(save_tstops)
    o.sde_adaptive = int(bool(sde_adaptive))
    if tstops is not None and len(tstops):
        tv = np.ascontiguousarray(tstops, dtype=np.float64)
        keep.append(tv); o.tstops = tv.ctypes.data_as(C.POINTER(C.c_double)); o.n_tstops = int(tv.shape[0])
    o.spec_arith = int(bool(spec_arith))   # step control to the letter of SURVEY A.4 / A.5 (the yardstick, not the contract)
    if sde_adaptive and dW is not None:    # injected STANDARD NORMALS [N][len], consumed in order by RSwM
        o.noise_stream_len = int(np.asarray(dW).reshape(N, -1).shape[1])
    get = (lambda w: (fns or {}).get(w)) if fns is not None else (lambda w: model_fn(model, w, f64))
    o.rhs, o.jac, o.noise = get("rhs"), get("jac"), get("noise")
    o.tgrad = get("tgrad") if fns else None
    o.cond, o.affect = (get("cond"), get("affect")) if event and not ncond else (None, None)
    o.affect_neg = (fns or {}).get("affect_neg") if (event and not ncond and fns) else None
    if ncond:   # VectorContinuousCallback
        o.vcond, o.vaffect, o.ncond, o.vterm_mask = get("vcond"), get("vaffect"), int(ncond), int(vterm_mask)
    o.dcond, o.daffect = (get("dcond"), get("daffect")) if devent else (None, None)
    o.devent_terminate = int(dterminate)
    out = np.empty((N, len(saveat), n), dtype=dtype)
    rc = np.zeros(N, dtype=np.int32)
    stats = np.zeros((N, 4), dtype=np.int32) if want_stats else None
    if dW is not None:
        dW = np.ascontiguousarray(dW, dtype=dtype)
    fn = L.orc_solve_f64 if f64 else L.orc_solve_f32
    vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    times = None
    if save_everystep:   # n_save = len(saveat) is the capacity; the saveat values themselves are ignored
        times = np.empty((N, len(saveat)), dtype=dtype)
        o.every_t = times.ctypes.data_as(C.c_void_p)
        o.save_everystep = 1
    if sde_adaptive:     # also return W(t_end) of every accepted path: (out, rc, stats, W)
        times = np.empty((N, n), dtype=dtype)
        o.every_t = times.ctypes.data_as(C.c_void_p)
    err = fn(C.byref(o), C.c_int64(N), vp(u0), vp(p), vp(saveat), vp(dW), vp(out), vp(rc), vp(stats),
             C.c_int(nthreads))
    if err != 0:
        raise RuntimeError(f"oracle error {err}")
    if save_everystep or sde_adaptive:
        return out, rc, stats, times
    return out, rc, stats


def fastpow(x, y):
    return float(lib().orc_fastpow(C.c_float(x), C.c_float(y)))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def normals(seed, traj, block, f64):
    L = lib()
    if f64:
        z = (C.c_double * 2)()
        L.orc_normals_f64(C.c_uint64(seed), C.c_uint64(traj), C.c_uint64(block), z)
    else:
        z = (C.c_float * 4)()
        L.orc_normals_f32(C.c_uint64(seed), C.c_uint64(traj), C.c_uint64(block), z)
    return np.array(list(z))
