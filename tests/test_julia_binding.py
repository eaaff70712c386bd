"""julia/EnsembleB200.jl is the reference-side binding a maintainer would add; Julia is not installed here, so the file
cannot be executed.  What CAN be checked mechanically is everything that would otherwise fail silently at the FFI:
struct layouts (field order, C type, count) against include/b200ens.h, the `ccall` signatures against the C prototypes,
and the enum tables.  (the ctypes mirror gets the same treatment in tests/test_abi.py.)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "b200ens.h")
JL = os.path.join(ROOT, "differentialequations.jl_b200", "julia", "EnsembleB200.jl")

C2JL = {"uint32_t": "UInt32", "int32_t": "Int32", "int64_t": "Int64", "uint64_t": "UInt64", "double": "Float64",
        "const char*": "Cstring", "const double*": "Ptr{Float64}", "const int32_t*": "Ptr{Int32}"}


def c_struct_fields(name):
    txt = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), txt, re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"((?:const )?\w+\s*\*?)\s*(.*)", decl)
        ctype = m.group(1).replace(" *", "*").strip()
        for nm in m.group(2).split(","):
            nm = nm.strip()
            if nm.startswith("*"):
                fields.append((nm[1:].strip(), ctype + "*"))
            else:
                fields.append((nm, ctype))
    return fields


def jl_struct_fields(name):
    txt = re.sub(r"#.*", "", open(JL).read())
    body = re.search(r"struct %s\b(.*?)\bend\b" % name, txt, re.S).group(1)
    body = re.sub(r"%s\(\)\s*=\s*new\(\)" % name, "", body)
    return [(a, b) for a, b in re.findall(r"(\w+)::([\w{}]+)", body)]


def test_struct_layouts_match_the_header():
    for cname, jname in (("b200ens_model_desc", "ModelDesc"), ("b200ens_opts", "Opts"), ("b200ens_stats", "Stats"),
                         ("b200ens_timing", "Timing")):
        c = c_struct_fields(cname)
        j = jl_struct_fields(jname)
        assert [n for n, _ in c] == [n for n, _ in j], (cname, [n for n, _ in c], [n for n, _ in j])
        for (n, ct), (_, jt) in zip(c, j):
            assert C2JL[ct] == jt, (cname, n, ct, jt)


def test_ccalls_name_exported_symbols_with_the_right_arity():
    hdr = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(b200ens_[a-z_]+)\s*\(([^;{]*?)\)\s*;", hdr, re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",")])
    jl = open(JL).read()
    calls = re.findall(r"ccall\(\(:(\w+), LIB\),\s*[\w{}]+,\s*\(([^)]*)\)", jl, re.S)
    assert len(calls) >= 7
    for sym, argt in calls:
        assert sym in protos, sym
        n = len([a for a in argt.split(",") if a.strip()])
        assert n == protos[sym], (sym, n, protos[sym])
    used = {c[0] for c in calls}
    for must in ("b200ens_abi_version", "b200ens_compile", "b200ens_solve", "b200ens_free", "b200ens_last_error",
                 "b200ens_opts_init", "b200ens_host_alloc", "b200ens_host_free"):
        assert must in used, must


def test_enum_tables_match_the_header():
    hdr = open(HDR).read()
    jl = open(JL).read()
    algs = dict(re.findall(r"B200ENS_(TSIT5|VERN7|ROSENBROCK23|RODAS5P|RODAS5|RODAS4|EM|SOSRA|SRIW1|FBDF) = (\d+)", hdr))
    jl_algs = {k.upper(): v for k, v in re.findall(r":(\w+) => (\d+)", re.search(r"const ALG_IDS = Dict\((.*?)\)", jl).group(1))}
    assert jl_algs == algs
    rcs = [n for n, _ in sorted(re.findall(r"B200ENS_RC_(\w+) = (\d+)", hdr), key=lambda x: int(x[1]))]
    jl_rcs = re.findall(r"ReturnCode\.(\w+)", re.search(r"const RETCODES = \((.*?)\)", jl, re.S).group(1))
    assert [r.lower() for r in jl_rcs] == [r.lower() for r in rcs]
    assert re.search(r"const ABI_VERSION = (\d+)", jl).group(1) == re.search(r"#define B200ENS_ABI_VERSION (\d+)", hdr).group(1)


def test_nothing_is_silently_dropped():
    """Every solve keyword the binding accepts is used, unknown ones are an error, and callbacks reach the model desc."""
    jl = open(JL).read()
    sig = re.search(r"function __solve\(.*?\)\n", jl, re.S).group(0)
    kws = re.findall(r"(\w+)\s*=", sig.split(";", 1)[1])
    body = jl.split(sig, 1)[1]
    for kw in kws:
        assert re.search(r"\b%s\b" % kw, body), f"keyword {kw} is accepted but never used"
    assert "isempty(kwargs) || error" in body
    assert "callback_sources(callback" in body and "cptr(csrc), cptr(asrc)" in body and "cptr(dcsrc), cptr(dasrc)" in body
    assert "eprob.output_func(" in body and "eprob.reduction(" in body and "batch_size" in body
    assert "EnsembleContext(" not in re.sub(r"#.*", "", jl)      # no invented constructor
