"""RHS / Jacobian / noise / callback code generation: symbolic form -> CUDA C.

Stand-in for the Symbolics/ModelingToolkit front-end the north star names (SURVEY.md 2.2
E7; README hyperlink /root/reference/README.md:54): the problem's functions are traced
with sympy symbols (as Symbolics traces a Julia `f`), the analytic Jacobian is derived
symbolically, and each function is emitted as a `__device__` CUDA-C function with the
fixed names include/b200ens.h documents (b2_rhs, b2_jac, b2_tgrad, b2_noise,
b2_condition, b2_affect).  The same text compiles as host C++ (tests build it with g++ for
the CPU oracle) given `#define __device__` -- nothing in it is CUDA-specific.

Numerical contract: literals are emitted as `(real)(<17 digits>)`, integer powers up to 4
are expanded to products, nothing is contracted into FMAs (the kernels are compiled with
--fmad=false), so the expression tree is the same on the GPU and in the oracle.
"""
import inspect

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter


class _RealPrinter(C99CodePrinter):
    """C printer that keeps all arithmetic in `real` (float or double)."""

    def _print_Float(self, expr):
        return "(real)(%.17g)" % float(expr)

    def _print_Integer(self, expr):
        return "(real)(%d)" % int(expr)

    def _print_Rational(self, expr):
        return "((real)(%d) / (real)(%d))" % (expr.p, expr.q)

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= int(e) <= 4:
            bs = self.parenthesize(b, 50)  # PRECEDENCE["Mul"]
            return "(" + " * ".join([bs] * int(e)) + ")"
        if e == -1:
            return "((real)(1) / %s)" % self.parenthesize(b, 100)
        if e.is_Integer and -4 <= int(e) <= -2:
            bs = self.parenthesize(b, 50)
            return "((real)(1) / (" + " * ".join([bs] * (-int(e))) + "))"
        if e == sp.Rational(1, 2):
            return "sqrt(%s)" % self._print(b)
        return "pow(%s, %s)" % (self._print(b), self._print(e))

    def _print_Symbol(self, expr):
        return expr.name


_printer = _RealPrinter({"math_macros": {}})


def _c(expr):
    return _printer.doprint(sp.sympify(expr))


class _Vec:
    """Indexable vector of symbols standing for a device array (u[i], p[i], du[i])."""

    def __init__(self, name, n):
        self.syms = [sp.Symbol(f"{name}[{i}]", real=True) for i in range(n)]
        self.vals = list(self.syms)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.vals[i]
        return self.vals[i]

    def __setitem__(self, i, v):
        self.vals[i] = sp.sympify(v)

    def __len__(self):
        return len(self.vals)

    def __iter__(self):
        return iter(self.vals)


def _nparams(fn):
    try:
        return len(inspect.signature(fn).parameters)
    except (TypeError, ValueError):
        return None


def trace_vector_fn(fn, n_state, n_param):
    """Trace f(u,p,t) -> sequence  or  f!(du,u,p,t) (in place) into sympy expressions."""
    u, p, t = _Vec("u", n_state), _Vec("p", max(n_param, 1)), sp.Symbol("t", real=True)
    if _nparams(fn) == 4:
        du = _Vec("du", n_state)
        du.vals = [sp.Integer(0)] * n_state
        fn(du, u, p, t)
        out = list(du.vals)
    else:
        out = fn(u, p, t)
        if not hasattr(out, "__len__"):
            out = [out]
        out = [sp.sympify(x) for x in out]
    if len(out) != n_state:
        raise ValueError(f"function returned {len(out)} components for n_state={n_state}")
    return out, u.syms, p.syms, t


def check_user_jacobian(jac, exprs, usyms, n_state, n_param):
    """ODEFunction(f; jac = ...): the device Jacobian is always derived from the traced f (exact, and the expression tree
    the oracle shares); a user-supplied jac(u,p,t) -> matrix or jac!(J,u,p,t) is traced too and must agree with it
    symbolically -- a wrong analytic Jacobian is an error here instead of a silently slower Rosenbrock solve."""
    u, p, t = _Vec("u", n_state), _Vec("p", max(n_param, 1)), sp.Symbol("t", real=True)
    if _nparams(jac) == 4:
        J = sp.zeros(n_state, n_state).as_mutable()
        jac(J, u, p, t)
    else:
        J = sp.Matrix(jac(u, p, t))
    if J.shape != (n_state, n_state):
        raise ValueError(f"jac returned a {J.shape} matrix for n_state={n_state}")
    ref = sp.Matrix(exprs).jacobian(sp.Matrix(usyms))
    for i in range(n_state):
        for j in range(n_state):
            if sp.simplify(sp.sympify(J[i, j]) - ref[i, j]) != 0:
                raise ValueError(f"jac[{i},{j}] = {J[i, j]} disagrees with d f[{i}] / d u[{j}] = {ref[i, j]}")


def _emit_body(assign_targets, exprs, cse=True):
    lines = []
    if cse:
        repl, red = sp.cse(list(exprs), symbols=sp.numbered_symbols("x"), order="none")
    else:
        repl, red = [], list(exprs)
    for s, e in repl:
        lines.append(f"    const real {s} = {_c(e)};")
    for tgt, e in zip(assign_targets, red):
        lines.append(f"    {tgt} = {_c(e)};")
    return "\n".join(lines)


_SIG = "__device__ __forceinline__ void {name}(real* __restrict__ {out}, const real* __restrict__ u, const real* __restrict__ p, const real t)"


def emit_rhs(exprs):
    n = len(exprs)
    body = _emit_body([f"du[{i}]" for i in range(n)], exprs)
    return _SIG.format(name="b2_rhs", out="du") + " {\n    (void)p; (void)t;\n" + body + "\n}\n"


def emit_jac(exprs, usyms):
    n = len(exprs)
    J = sp.Matrix(exprs).jacobian(sp.Matrix(usyms))
    entries = [J[i, j] for i in range(n) for j in range(n)]
    body = _emit_body([f"J[{k}]" for k in range(n * n)], entries)
    return _SIG.format(name="b2_jac", out="J") + " {\n    (void)p; (void)t;\n" + body + "\n}\n"


def emit_tgrad(exprs, t):
    d = [sp.diff(e, t) for e in exprs]
    if all(x == 0 for x in d):
        return None
    body = _emit_body([f"dT[{i}]" for i in range(len(d))], d)
    return _SIG.format(name="b2_tgrad", out="dT") + " {\n    (void)p; (void)t;\n" + body + "\n}\n"


def emit_mass(M):
    """Constant mass matrix as a compile-time table: after unrolling the zero entries vanish from the kernel."""
    M = np.asarray(M, dtype=np.float64)
    vals = ", ".join(repr(float(v)) for v in M.reshape(-1))
    return ("#undef B2_HAS_MASS\n#define B2_HAS_MASS 1\n"
            f"static constexpr double B2_MASS_[{M.size}] = {{{vals}}};\n")


def emit_noise(exprs):
    body = _emit_body([f"g[{i}]" for i in range(len(exprs))], exprs)
    return _SIG.format(name="b2_noise", out="g") + " {\n    (void)u; (void)p; (void)t;\n" + body + "\n}\n"


class _TraceIntegrator:
    """What `condition(u,t,integrator)` / `affect!(integrator)` see while being traced."""

    def __init__(self, n_state, n_param):
        object.__setattr__(self, "_frozen", False)
        self.u = _Vec("u", n_state)
        # p and t are READ-ONLY while tracing: only changes to integrator.u can be emitted (the kernel's affect functions
        # take p as const), so `integrator.p[i] = ...` (parameter switches, dosing) or `integrator.t = ...` raise instead
        # of being dropped silently
        self.p = _ReadOnlySeq(_Vec("p", max(n_param, 1)).syms, "integrator.p")
        self.t = sp.Symbol("t", real=True)
        self.terminated = False
        object.__setattr__(self, "_frozen", True)

    def __setattr__(self, name, value):
        if getattr(self, "_frozen", False) and name not in ("terminated",):
            if name == "u":
                raise NotImplementedError("affect! must modify integrator.u in place (integrator.u[i] = ...), not rebind it")
            raise NotImplementedError(f"affect! may only modify integrator.u (tried to set integrator.{name}): writes to p or t "
                                      "cannot be emitted for the device and are not silently dropped")
        object.__setattr__(self, name, value)


class _ReadOnlySeq(tuple):
    """integrator.p while tracing: indexable and iterable, assignment raises."""

    def __new__(cls, items, what):
        obj = super().__new__(cls, items)
        obj._what = what
        return obj

    def __setitem__(self, i, v):
        raise NotImplementedError(f"affect! may only modify integrator.u (tried to assign {self._what}[{i}]): parameter writes "
                                  "cannot be emitted for the device and are not silently dropped")


def terminate_b(integrator):
    """Julia's terminate!(integrator)."""
    integrator.terminated = True


def emit_callback(cb, n_state, n_param):
    """ContinuousCallback(condition, affect!) -> (condition_src, affect_src, terminate)."""
    integ = _TraceIntegrator(n_state, n_param)
    try:
        g = cb.condition(integ.u, integ.t, integ)
    except NotImplementedError:
        raise   # a precise rejection (e.g. affect! writing integrator.p) is the message the user needs
    except Exception as e:  # not expressible symbolically -> reject (no CPU fallback), SURVEY 7.3
        raise NotImplementedError(
            "ContinuousCallback condition is not symbolically traceable; EnsembleB200 only accepts "
            "callbacks that can be emitted as CUDA C") from e
    # which components of u the condition reads: the event search builds its interpolation polynomial only for these
    gs = sp.sympify(g)
    mask = 0
    for i, sym in enumerate(integ.u.syms):
        if gs.has(sym):
            mask |= 1 << i
    cond_src = (f"#undef B2_COND_MASK\n#define B2_COND_MASK 0x{mask:x}u\n"
                "__device__ __forceinline__ real b2_condition(const real* __restrict__ u, const real* __restrict__ p, "
                "const real t) {\n    (void)u; (void)p; (void)t;\n    return " + _c(g) + ";\n}\n")
    direction = getattr(cb, "direction", 0)
    if direction:
        cond_src = f"#undef B2_EVENT_DIR\n#define B2_EVENT_DIR {direction}\n" + cond_src

    def trace_affect(fn, name):
        it = _TraceIntegrator(n_state, n_param)
        try:
            fn(it)
        except NotImplementedError:
            raise   # a precise rejection (e.g. affect! writing integrator.p) is the message the user needs
        except Exception as e:
            raise NotImplementedError("ContinuousCallback affect! is not symbolically traceable") from e
        lines = []
        changed = [(i, v) for i, (s, v) in enumerate(zip(it.u.syms, it.u.vals)) if v != s]
        # evaluate all right-hand sides before assigning (affect! sees the pre-event state)
        for i, v in changed:
            lines.append(f"    const real n{i} = {_c(v)};")
        for i, _ in changed:
            lines.append(f"    u[{i}] = n{i};")
        src = (f"__device__ __forceinline__ void {name}(real* __restrict__ u, const real* __restrict__ p, "
               "const real t) {\n    (void)u; (void)p; (void)t;\n" + "\n".join(lines) + "\n}\n")
        return src, bool(it.terminated)

    affect_neg = getattr(cb, "affect_neg", cb.affect)
    if direction < 0:     # downcrossings only: the one affect that can run is affect_neg!
        aff_src, term = trace_affect(affect_neg, "b2_affect")
        return cond_src, aff_src, int(term)
    aff_src, term = trace_affect(cb.affect, "b2_affect")
    if direction == 0 and affect_neg is not cb.affect:
        neg_src, term_neg = trace_affect(affect_neg, "b2_affect_neg")
        aff_src += "#undef B2_HAS_AFFECT_NEG\n#define B2_HAS_AFFECT_NEG 1\n" + neg_src
        return cond_src, aff_src, int(term) | (4 if term_neg else 0)
    return cond_src, aff_src, int(term)


def emit_vector_callback(cb, n_state, n_param):
    """VectorContinuousCallback(condition!, affect!, len) -> (vcondition_src, vaffect_src, terminate).
    condition(out, u, t, integrator) fills out[0..len); affect(integrator, idx) is traced once per 1-based idx."""
    nc = cb.len
    integ = _TraceIntegrator(n_state, n_param)
    out = _Vec("g", nc)
    try:
        cb.condition(out, integ.u, integ.t, integ)
    except NotImplementedError:
        raise   # a precise rejection (e.g. affect! writing integrator.p) is the message the user needs
    except Exception as e:
        raise NotImplementedError("VectorContinuousCallback condition is not symbolically traceable; EnsembleB200 "
                                  "only accepts callbacks that can be emitted as CUDA C") from e
    gs = [sp.sympify(v) for v in out.vals]
    if any(v == sym for v, sym in zip(gs, out.syms)):
        raise ValueError("VectorContinuousCallback: condition! must assign every out[i]")
    mask = 0
    for i, sym in enumerate(integ.u.syms):
        if any(g.has(sym) for g in gs):
            mask |= 1 << i
    body = _emit_body([f"g[{k}]" for k in range(nc)], gs)
    direction = getattr(cb, "direction", 0)
    vaffect = cb.affect_neg if direction < 0 else cb.affect
    cond_src = ((f"#undef B2_EVENT_DIR\n#define B2_EVENT_DIR {direction}\n" if direction else "") +
                f"#undef B2_COND_MASK\n#define B2_COND_MASK 0x{mask:x}u\n#define B2_NCOND {nc}\n"
                "__device__ __forceinline__ void b2_vcondition(real* __restrict__ g, const real* __restrict__ u, "
                "const real* __restrict__ p, const real t) {\n    (void)u; (void)p; (void)t;\n" + body + "\n}\n")
    cases, terminated = [], []
    for k in range(nc):
        it = _TraceIntegrator(n_state, n_param)
        try:
            vaffect(it, k + 1)
        except NotImplementedError:
            raise   # a precise rejection (e.g. affect! writing integrator.p) is the message the user needs
        except Exception as e:
            raise NotImplementedError("VectorContinuousCallback affect! is not symbolically traceable") from e
        changed = [(i, v) for i, (sy, v) in enumerate(zip(it.u.syms, it.u.vals)) if v != sy]
        lines = [f"        const real n{i} = {_c(v)};" for i, v in changed] + [f"        u[{i}] = n{i};" for i, _ in changed]
        cases.append(f"    case {k}: {{\n" + "\n".join(lines) + "\n    }} break;".replace("}}", "}"))
        terminated.append(bool(it.terminated))
    tmask = sum(1 << k for k, tm in enumerate(terminated) if tm)   # which event indices call terminate!(integrator)
    aff_src = (f"#define B2_VTERM_MASK 0x{tmask:x}u\n"
               "__device__ __forceinline__ void b2_vaffect(real* __restrict__ u, const real* __restrict__ p, const real t, "
               "const int idx) {\n    (void)u; (void)p; (void)t;\n    switch (idx) {\n" + "\n".join(cases) +
               "\n    default: break;\n    }\n}\n")
    return cond_src, aff_src, False   # termination is per event index (B2_VTERM_MASK), not the global flag


class _TimeProxy:
    """The `t` a DiscreteCallback condition is traced with.  Julia's `t == 4.0` on a symbolic time is a symbolic
    equation (upstream's dosing idiom: condition(u, t, integrator) = t == 4.0 together with tstops = [4.0]); Python's
    `==` on a sympy Symbol is structural and answers False.  This proxy forwards arithmetic to the symbol and turns
    the comparisons into sympy relationals."""

    def __init__(self, sym):
        self._s = sym

    def _sympy_(self):
        return self._s

    def __eq__(self, other):
        return sp.Eq(self._s, sp.sympify(other))

    def __ne__(self, other):
        return sp.Ne(self._s, sp.sympify(other))

    __hash__ = None

    def __lt__(self, o): return self._s < sp.sympify(o)
    def __le__(self, o): return self._s <= sp.sympify(o)
    def __gt__(self, o): return self._s > sp.sympify(o)
    def __ge__(self, o): return self._s >= sp.sympify(o)
    def __add__(self, o): return self._s + sp.sympify(o)
    def __radd__(self, o): return sp.sympify(o) + self._s
    def __sub__(self, o): return self._s - sp.sympify(o)
    def __rsub__(self, o): return sp.sympify(o) - self._s
    def __mul__(self, o): return self._s * sp.sympify(o)
    def __rmul__(self, o): return sp.sympify(o) * self._s
    def __truediv__(self, o): return self._s / sp.sympify(o)
    def __rtruediv__(self, o): return sp.sympify(o) / self._s
    def __pow__(self, o): return self._s ** sp.sympify(o)
    def __neg__(self): return -self._s


def emit_discrete_callback(cb, n_state, n_param):
    """DiscreteCallback(condition, affect!) -> (dcondition_src, daffect_src, terminate).  condition(u,t,integrator)
    must trace to a sympy relational / boolean (e.g. `t >= 0.5`, `(u[0] > 1) & (t < 3)`)."""
    integ = _TraceIntegrator(n_state, n_param)
    try:
        g = cb.condition(integ.u, _TimeProxy(integ.t), integ)
        g = sp.sympify(g)
        if not (g.is_Relational or g.is_Boolean or g in (sp.true, sp.false)):
            raise TypeError("condition must be a boolean expression")
    except NotImplementedError:
        raise   # a precise rejection (e.g. affect! writing integrator.p) is the message the user needs
    except Exception as e:
        raise NotImplementedError(
            "DiscreteCallback condition is not symbolically traceable to a boolean expression; EnsembleB200 only "
            "accepts callbacks that can be emitted as CUDA C") from e
    cond_src = ("__device__ __forceinline__ bool b2_dcondition(const real* __restrict__ u, const real* __restrict__ p, "
                "const real t) {\n    (void)u; (void)p; (void)t;\n    return " + _c(g) + ";\n}\n")
    integ2 = _TraceIntegrator(n_state, n_param)
    try:
        cb.affect(integ2)
    except NotImplementedError:
        raise   # a precise rejection (e.g. affect! writing integrator.p) is the message the user needs
    except Exception as e:
        raise NotImplementedError("DiscreteCallback affect! is not symbolically traceable") from e
    changed = [(i, v) for i, (s, v) in enumerate(zip(integ2.u.syms, integ2.u.vals)) if v != s]
    lines = [f"    const real n{i} = {_c(v)};" for i, v in changed] + [f"    u[{i}] = n{i};" for i, _ in changed]
    aff_src = ("__device__ __forceinline__ void b2_daffect(real* __restrict__ u, const real* __restrict__ p, "
               "const real t) {\n    (void)u; (void)p; (void)t;\n" + "\n".join(lines) + "\n}\n")
    return cond_src, aff_src, bool(integ2.terminated)


HOST_PRELUDE = """// host build of emitted model code (tests: compiled with g++ for the CPU oracle)
#include <cmath>
using std::sqrt; using std::pow; using std::sin; using std::cos; using std::exp; using std::log; using std::fabs;
#define __device__
#define __forceinline__ inline
"""


def host_wrapper_source(sources, names=("b2_rhs", "b2_jac", "b2_tgrad", "b2_noise", "b2_condition", "b2_affect", "b2_affect_neg",
                                        "b2_dcondition", "b2_daffect", "b2_vcondition", "b2_vaffect")):
    """C++ translation unit exposing the emitted functions for float and double with C linkage
    (oracle side of the parity tests)."""
    parts = [HOST_PRELUDE]
    for suf, ty in (("f32", "float"), ("f64", "double")):
        parts.append(f"namespace ns_{suf} {{\ntypedef {ty} real;\n" + "\n".join(s for s in sources if s) + "\n}\n")
    parts.append('extern "C" {\n')
    present = "\n".join(s for s in sources if s)
    for suf, ty in (("f32", "float"), ("f64", "double")):
        for nm in names:
            if nm + "(" not in present:
                continue
            if nm == "b2_condition":
                parts.append(f"{ty} {nm}_{suf}(const {ty}* u, const {ty}* p, {ty} t) {{ return ns_{suf}::{nm}(u, p, t); }}\n")
            elif nm == "b2_dcondition":
                parts.append(f"int {nm}_{suf}(const {ty}* u, const {ty}* p, {ty} t) {{ return ns_{suf}::{nm}(u, p, t) ? 1 : 0; }}\n")
            elif nm == "b2_vaffect":
                parts.append(f"void {nm}_{suf}({ty}* u, const {ty}* p, {ty} t, int idx) {{ ns_{suf}::{nm}(u, p, t, idx); }}\n")
            elif nm in ("b2_affect", "b2_affect_neg", "b2_daffect"):
                parts.append(f"void {nm}_{suf}({ty}* u, const {ty}* p, {ty} t) {{ ns_{suf}::{nm}(u, p, t); }}\n")
            else:
                parts.append(f"void {nm}_{suf}({ty}* o, const {ty}* u, const {ty}* p, {ty} t) {{ ns_{suf}::{nm}(o, u, p, t); }}\n")
    parts.append("}\n")
    return "".join(parts)
