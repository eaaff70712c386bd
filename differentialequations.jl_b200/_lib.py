"""ctypes binding of libb200ens.so (include/b200ens.h).  Fails loudly when the library is
missing -- there is no Python/CPU fallback for the hot path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libb200ens.so")

F32, F64 = 0, 1
ALG_IDS = {"Tsit5": 1, "Vern7": 2, "Rosenbrock23": 3, "Rodas5": 4, "Rodas5P": 5, "EM": 6, "SOSRA": 7, "Rodas4": 8, "SRIW1": 9, "FBDF": 10}
MODEL_FAST_MATH = 1
MODEL_KSMEM = 4
MODEL_SPLIT = 8
MODEL_NOSPLIT = 16
MODEL_SDE_ADAPTIVE = 32

E_NODEVICE = -3


class B200EnsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200ens error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("n_state", C.c_int32), ("n_param", C.c_int32), ("dtype", C.c_int32),
        ("alg", C.c_int32), ("flags", C.c_uint32),
        ("rhs_src", C.c_char_p), ("jac_src", C.c_char_p), ("tgrad_src", C.c_char_p), ("noise_src", C.c_char_p),
        ("condition_src", C.c_char_p), ("affect_src", C.c_char_p), ("name", C.c_char_p),
        ("dcondition_src", C.c_char_p), ("daffect_src", C.c_char_p),
        ("save_idxs", C.POINTER(C.c_int32)), ("n_save_idxs", C.c_int32), ("reserved0", C.c_int32),
    ]


class Opts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("adaptive", C.c_int32),
        ("t0", C.c_double), ("t1", C.c_double), ("dt", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
        ("dtmin", C.c_double), ("dtmax", C.c_double),
        ("qmin", C.c_double), ("qmax", C.c_double), ("gamma", C.c_double), ("beta1", C.c_double),
        ("beta2", C.c_double), ("qoldinit", C.c_double),
        ("maxiters", C.c_int64), ("seed", C.c_uint64), ("traj_offset", C.c_uint64),
        ("noise_injected", C.c_int32), ("event_terminate", C.c_int32), ("interp_points", C.c_int32),
        ("save_tstops", C.c_int32), ("device_mask", C.c_uint32), ("refill_threshold", C.c_int32),
        ("block_threads", C.c_int32), ("stage_outputs", C.c_int32),
        ("work_order", C.c_int32), ("save_everystep", C.c_int32),
        ("abstol_vec", C.POINTER(C.c_double)), ("reltol_vec", C.POINTER(C.c_double)),
        ("noise_stream_len", C.c_int64), ("shard_blocks", C.c_int32), ("n_tstops", C.c_int32),
        ("tstops", C.POINTER(C.c_double)),
    ]


class Stats(C.Structure):
    _fields_ = [("naccept", C.c_int32), ("nreject", C.c_int32), ("nf", C.c_int32), ("nevents", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
        ("n_devices", C.c_int32), ("launches", C.c_int32),
        ("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32), ("regs", C.c_int32),
        ("kernel_ms_min", C.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "b200ens_abi_version", "b200ens_device_count", "b200ens_last_error", "b200ens_nvrtc_info", "b200ens_opts_init", "b200ens_compile",
    "b200ens_free", "b200ens_model_info", "b200ens_solve", "b200ens_solve_device", "b200ens_solve_moments",
    "b200ens_host_alloc", "b200ens_host_free",
]

_lib = None


def lib():
    """Load libb200ens.so (built in-tree by `make -C csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
            "(python __graft_entry__.py build).  The B200 ensemble path has no fallback.")
    L = C.CDLL(LIB_PATH)
    L.b200ens_abi_version.restype = C.c_int
    L.b200ens_nvrtc_info.restype = C.c_char_p
    L.b200ens_device_count.restype = C.c_int
    L.b200ens_last_error.restype = C.c_char_p
    L.b200ens_opts_init.argtypes = [C.POINTER(Opts)]
    L.b200ens_opts_init.restype = None
    L.b200ens_compile.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.b200ens_compile.restype = C.c_int
    L.b200ens_free.argtypes = [C.c_void_p]
    L.b200ens_free.restype = None
    L.b200ens_model_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32)]
    L.b200ens_model_info.restype = C.c_int
    L.b200ens_solve.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(Timing)]
    L.b200ens_solve.restype = C.c_int
    L.b200ens_solve_device.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.POINTER(Timing)]
    L.b200ens_solve_device.restype = C.c_int
    L.b200ens_solve_moments.argtypes = [C.c_void_p, C.POINTER(Opts), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                        C.c_void_p, C.POINTER(Timing)]
    L.b200ens_solve_moments.restype = C.c_int
    L.b200ens_host_alloc.argtypes = [C.c_size_t]
    L.b200ens_host_alloc.restype = C.c_void_p
    L.b200ens_host_free.argtypes = [C.c_void_p]
    L.b200ens_host_free.restype = None
    if L.b200ens_abi_version() != 7:
        raise ImportError("libb200ens ABI version mismatch")
    _lib = L
    return L


def check(code):
    if code != 0:
        raise B200EnsError(code, lib().b200ens_last_error().decode(errors="replace"))


def default_opts():
    o = Opts()
    lib().b200ens_opts_init(C.byref(o))
    return o


class Model:
    """A compiled (problem functions x algorithm x dtype) kernel: b200ens_model*."""

    def __init__(self, n_state, n_param, dtype, alg, rhs_src, jac_src=None, tgrad_src=None, noise_src=None,
                 condition_src=None, affect_src=None, name="model", fast_math=False,
                 dcondition_src=None, daffect_src=None, ksmem=False, split=None, sde_adaptive=False, save_idxs=None):
        L = lib()
        d = ModelDesc()
        d.struct_size = C.sizeof(ModelDesc)
        d.n_state, d.n_param = n_state, n_param
        d.dtype = F64 if np.dtype(dtype) == np.float64 else F32
        d.alg = ALG_IDS[alg] if isinstance(alg, str) else int(alg)
        d.flags = ((MODEL_FAST_MATH if fast_math else 0)
                   | (MODEL_KSMEM if ksmem else 0) | (MODEL_SPLIT if split is True else 0)
                   | (MODEL_NOSPLIT if split is False else 0) | (MODEL_SDE_ADAPTIVE if sde_adaptive else 0))
        enc = lambda s: s.encode() if s is not None else None
        d.rhs_src, d.jac_src, d.tgrad_src = enc(rhs_src), enc(jac_src), enc(tgrad_src)
        d.noise_src, d.condition_src, d.affect_src = enc(noise_src), enc(condition_src), enc(affect_src)
        d.name = enc(name)
        d.dcondition_src, d.daffect_src = enc(dcondition_src), enc(daffect_src)
        idx_arr = None
        if save_idxs is not None:   # solve(...; save_idxs): the output rows hold these components only
            idx_arr = (C.c_int32 * len(save_idxs))(*[int(i) for i in save_idxs])
            d.save_idxs, d.n_save_idxs = idx_arr, len(save_idxs)
        self.handle = C.c_void_p()
        logbuf = C.create_string_buffer(1 << 16)
        code = L.b200ens_compile(C.byref(d), C.byref(self.handle), logbuf, len(logbuf))
        self.log = logbuf.value.decode(errors="replace")
        check(code)
        self.n_state, self.n_param, self.dtype, self.alg = n_state, n_param, np.dtype(dtype), alg
        self.n_out = len(save_idxs) if save_idxs is not None else n_state   # entries of an output row
        self.has_event = condition_src is not None or dcondition_src is not None

    def info(self):
        cb, regs, smem, lmem = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().b200ens_model_info(self.handle, C.byref(cb), C.byref(regs), C.byref(smem), C.byref(lmem)))
        return {"cubin_bytes": cb.value, "regs": regs.value, "smem": smem.value, "lmem": lmem.value}

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                lib().b200ens_free(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ---- host-buffer solve (b200ens_solve)
    def solve(self, opts, u0, p, saveat, dW=None, want_stats=True, out=None, rc=None, stats=None):
        N = u0.shape[0]
        dt = self.dtype
        u0 = np.ascontiguousarray(u0, dtype=dt)
        p = np.ascontiguousarray(p, dtype=dt).reshape(N, self.n_param)
        saveat = np.ascontiguousarray(saveat, dtype=dt)
        n_save = saveat.shape[0]
        assert u0.shape == (N, self.n_state)
        if out is None:
            out = np.empty((N, n_save, self.n_out), dtype=dt)
        if rc is None:
            rc = np.zeros(N, dtype=np.int32)
        if stats is None and want_stats:
            stats = np.zeros((N, 4), dtype=np.int32)
        if dW is not None:
            dW = np.ascontiguousarray(dW, dtype=dt)
        tm = Timing()
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        check(lib().b200ens_solve(self.handle, C.byref(opts), N, vp(u0), vp(p), vp(saveat), n_save, vp(dW), vp(out),
                                  None, vp(rc), vp(stats), C.byref(tm)))
        return out, rc, stats, tm

    # ---- every accepted step (opts.save_everystep = 1): capacity slots per trajectory, per-trajectory step times
    def solve_everystep(self, opts, u0, p, capacity):
        N = u0.shape[0]
        dt = self.dtype
        u0 = np.ascontiguousarray(u0, dtype=dt)
        p = np.ascontiguousarray(p, dtype=dt).reshape(N, self.n_param)
        out = np.empty((N, capacity, self.n_out), dtype=dt)
        times = np.empty((N, capacity), dtype=dt)
        rc = np.zeros(N, dtype=np.int32)
        stats = np.zeros((N, 4), dtype=np.int32)
        tm = Timing()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        o = copy_opts(opts)
        o.save_everystep = 1
        check(lib().b200ens_solve(self.handle, C.byref(o), N, vp(u0), vp(p), None, int(capacity), None, vp(out), vp(times),
                                  vp(rc), vp(stats), C.byref(tm)))
        return out, times, rc, stats, tm

    # ---- ensemble moments without shipping trajectories to the host (b200ens_solve_moments)
    def solve_moments(self, opts, u0, p, saveat, dW=None, rc=None):
        N = u0.shape[0]
        dt = self.dtype
        u0 = np.ascontiguousarray(u0, dtype=dt)
        p = np.ascontiguousarray(p, dtype=dt).reshape(N, self.n_param)
        saveat = np.ascontiguousarray(saveat, dtype=dt)
        n_save = saveat.shape[0]
        s = np.zeros((n_save, self.n_out))
        q = np.zeros((n_save, self.n_out))
        cnt = C.c_int64(0)
        if rc is None:
            rc = np.zeros(N, dtype=np.int32)
        if dW is not None:
            dW = np.ascontiguousarray(dW, dtype=dt)
        tm = Timing()
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        check(lib().b200ens_solve_moments(self.handle, C.byref(opts), N, vp(u0), vp(p), vp(saveat), n_save, vp(dW),
                                          vp(s), vp(q), C.byref(cnt), vp(rc), C.byref(tm)))
        return s, q, cnt.value, rc, tm

    # ---- device-buffer solve (b200ens_solve_device); pointers are raw ints (e.g. torch .data_ptr())
    def solve_device(self, opts, device, stream, N, d_u0, d_p, d_saveat, n_save, d_out, d_rc, d_stats=None,
                     d_dW=None, timed=True):
        tm = Timing()
        check(lib().b200ens_solve_device(self.handle, C.byref(opts), device, stream, N, d_u0, d_p, d_saveat, n_save,
                                         d_dW, d_out, d_rc, d_stats, C.byref(tm) if timed else None))
        return tm


def copy_opts(o):
    """A by-value copy of a b200ens_opts struct."""
    c = Opts()
    C.memmove(C.byref(c), C.byref(o), C.sizeof(Opts))
    c._tol_keep = getattr(o, "_tol_keep", None)   # keep per-component tolerance arrays alive with the copy
    return c


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc memory (zero-staging H2D/D2H)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib().b200ens_host_alloc(max(n, 1))
    if not ptr:
        raise MemoryError(lib().b200ens_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _pinned[arr.__array_interface__["data"][0]] = ptr
    return arr


_pinned = {}
