"""The C-ABI library: loads, exports every symbol include/b200ens.h declares, struct layouts match the
ctypes mirror, JIT compilation works without a GPU, and solving without a device fails loudly."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "b200ens.h")


def test_exports_every_declared_symbol(B):
    L = B._lib.lib()
    declared = set(re.findall(r"\b(b200ens_[a-z_]+)\s*\(", open(HDR).read()))
    assert declared == set(B._lib.EXPORTS), declared ^ set(B._lib.EXPORTS)
    nm = subprocess.check_output(["nm", "-D", "--defined-only", B._lib.LIB_PATH], text=True)
    for sym in declared:
        assert hasattr(L, sym) and re.search(rf"\bT {sym}\b", nm), sym
    assert L.b200ens_abi_version() == 7


def test_struct_layouts_match_header(B, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "b200ens.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(b200ens_model_desc),'
                   ' sizeof(b200ens_opts), sizeof(b200ens_stats), sizeof(b200ens_timing)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    L = B._lib
    assert sizes == [C.sizeof(L.ModelDesc), C.sizeof(L.Opts), C.sizeof(L.Stats), C.sizeof(L.Timing)]


def test_header_is_plain_c_and_has_no_torch_types():
    txt = open(HDR).read()
    assert 'extern "C"' in txt and "torch" not in txt.lower() and "at::" not in txt


def test_kernel_args_layout_is_mirrored():
    """struct B2Args must be declared identically on the host and in the kernel header."""
    grab = lambda path: re.sub(r"//.*", "", re.search(r"struct B2Args \{(.*?)\n\};", open(path).read(), re.S).group(1))
    norm = lambda s: re.sub(r"\s+", " ", s.replace("b200ens_stats", "B2Stats")).strip()
    host = grab(os.path.join(ROOT, "differentialequations.jl_b200", "csrc", "b200ens.cpp"))
    dev = grab(os.path.join(ROOT, "differentialequations.jl_b200", "csrc", "kernels", "b2_common.cuh"))
    assert norm(host) == norm(dev)


@pytest.mark.parametrize("alg", ["Tsit5", "Vern7", "Rosenbrock23", "Rodas4", "Rodas5", "Rodas5P", "EM", "SOSRA"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_every_stepper_jit_compiles_for_sm100a_without_a_gpu(B, alg, dtype):
    from b200ens import workloads as W

    prob = W.lorenz_additive_problem(dtype) if alg in ("EM", "SOSRA") else W.lorenz_problem(dtype)
    m = B.build_model(prob, getattr(B, alg)())
    info = m.info()
    assert info["cubin_bytes"] > 10000 and 0 < info["regs"] <= 255
    assert "sm_100a" in m.log


@pytest.mark.parametrize("alg", ["SRIW1", "SOSRA"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_adaptive_sde_kernel_jit_compiles(B, alg, dtype):
    """B200ENS_MODEL_SDE_ADAPTIVE (include/b200ens.h): the RSwM kernel compiles for sm_100a; its local-memory
    frame is the stack of remembered Brownian increments, not register spills."""
    from b200ens import workloads as W

    m = B.build_model(W.lorenz_additive_problem(dtype), getattr(B, alg)(), sde_adaptive=True)
    info = m.info()
    n, es = 3, np.dtype(dtype).itemsize
    assert 0 < info["regs"] <= 255 and info["lmem"] >= 48 * (1 + 2 * n) * es
    with pytest.raises(B.B200EnsError) as e:
        B.build_model(W.lorenz_additive_problem(dtype), B.EM(), sde_adaptive=True)
    assert e.value.code == -6
    with pytest.raises(B.B200EnsError) as e:      # shipped, un-gated: reaches the library, which has no CPU fallback
        B.solve(W.lorenz_additive_problem(dtype), getattr(B, alg)(), adaptive=True, dt=0.1, saveat=1.0)
    assert e.value.code == -3


def test_compile_errors_are_reported(B):
    with pytest.raises(B.B200EnsError) as e:
        B.Model(3, 3, np.float64, "Tsit5", "__device__ void b2_rhs(real* du, const real* u, const real* p, real t) { du[0] = nope; }")
    assert e.value.code == -2 and "nope" in str(e.value)
    with pytest.raises(B.B200EnsError) as e:   # Rosenbrock without a Jacobian: no AD / finite-difference fallback
        B.Model(3, 3, np.float64, "Rodas5P", "__device__ void b2_rhs(real* du, const real* u, const real* p, real t) {}")
    assert e.value.code == -6


def test_no_cpu_fallback_without_device(B):
    L = B._lib.lib()
    if L.b200ens_device_count() > 0:
        pytest.skip("a GPU is visible here")
    from b200ens import workloads as W

    with pytest.raises(B.B200EnsError) as e:
        B.solve(W.lorenz_problem(), B.Tsit5(), saveat=1.0, dt=0.1)
    assert e.value.code == B._lib.E_NODEVICE and "no CPU fallback" in str(e.value)


def test_saveat_grid_is_validated_before_any_device_work(B):
    """include/b200ens.h: saveat is ascending and inside tspan -- checked on the host, so it also fails without a GPU."""
    from b200ens import workloads as W

    for dtype in (np.float64, np.float32):
        m = B.build_model(W.lorenz_problem(dtype), B.Tsit5())
        o = B._lib.default_opts()
        o.t0, o.t1, o.dt = 0.0, 10.0, 0.1
        u0, p = W.lorenz_params(4, "random", seed=1, dtype=dtype)
        for bad in ([0.0, 5.0, 4.0], [0.0, 10.5], [-1.0, 2.0], [1.0, float("nan")]):
            with pytest.raises(B.B200EnsError) as e:
                m.solve(o, u0, p, bad)
            assert e.value.code == -1 and "saveat" in str(e.value), bad


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may reference it."""
    pkg = os.path.join(ROOT, "differentialequations.jl_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "orc_solve" not in txt, os.path.join(dp, fn)
