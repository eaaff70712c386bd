"""Coefficient tables (SURVEY.md Appendix B -> tools/tableaus.json -> generated headers) must satisfy
their nominal order conditions (rooted-tree theory, tools/rk_trees.py)."""
import json
import os

import numpy as np

from rk_trees import order_residuals, trees, gamma, order, phi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = json.load(open(os.path.join(ROOT, "tools", "tableaus.json")))


def _tsit5():
    t = {k: float(v) for k, v in T["tsit5"].items()}
    A = np.zeros((7, 7))
    for i in range(2, 8):
        for j in range(1, i):
            A[i - 1, j - 1] = t[f"a{i}{j}"]
    bt = np.array([t[f"btilde{i}"] for i in range(1, 8)])
    return t, A, A[6].copy(), bt


def test_tsit5_order5_embedded4():
    t, A, b, bt = _tsit5()
    res = order_residuals(A, b, 6)
    assert max(res[q] for q in range(1, 6)) < 1e-13 and res[6] > 1e-5
    res = order_residuals(A, b - bt, 5)
    assert max(res[q] for q in range(1, 5)) < 1e-13 and res[5] > 1e-5
    c = np.array([0, t["c2"], t["c3"], t["c4"], t["c5"], 1, 1])
    assert np.abs(A.sum(1) - c).max() < 1e-14


def test_tsit5_dense_output_order4():
    t, A, b, _ = _tsit5()
    r = np.zeros((7, 4))
    r[0, 0] = t["r11"]
    for i in range(1, 8):
        for k in range(2, 5):
            r[i - 1, k - 1] = t[f"r{i}{k}"]
    for th in (0.1, 0.37, 0.5, 0.9, 1.0):
        bth = sum(r[:, k] * th ** (k + 1) for k in range(4))
        for q in range(1, 5):
            for tr in trees(q):
                assert abs(bth @ phi(tr, A) - th ** q / gamma(tr)) < 1e-12
    assert np.abs(r.sum(1) - b).max() < 1e-13   # b_i(1) = b_i


def _vern7():
    A = np.zeros((16, 16))
    for src in ("vern7", "vern7_extra"):
        for k, v in T[src].items():
            if k.startswith("a"):
                A[int(k[1:3]) - 1, int(k[3:5]) - 1] = float(v)
    b = np.zeros(16)
    bt = np.zeros(16)
    for k, v in T["vern7"].items():
        if k.startswith("btilde"):
            bt[int(k[6:]) - 1] = float(v)
        elif k.startswith("b"):
            b[int(k[1:]) - 1] = float(v)
    return A, b, bt


def test_vern7_order7_embedded6():
    A, b, bt = _vern7()
    res = order_residuals(A[:10, :10], b[:10], 8)
    assert max(res[q] for q in range(1, 8)) < 1e-13 and res[8] > 1e-6
    res = order_residuals(A[:10, :10], (b - bt)[:10], 7)
    assert max(res[q] for q in range(1, 7)) < 1e-13 and res[7] > 1e-5


def test_vern7_derived_dense_output_order6():
    """tools/derive_vern7_dense.py: order<=6 continuous conditions and continuity at theta=1."""
    A, b, _ = _vern7()
    D = json.load(open(os.path.join(ROOT, "tools", "vern7_dense.json")))
    r = np.zeros((16, 6))
    for s, coef in D["r"].items():
        r[int(s) - 1] = [float(x) for x in coef]
    assert np.abs(r.sum(1) - b).max() < 1e-13
    for th in (0.2, 0.5, 0.77):
        bth = sum(r[:, k] * th ** (k + 1) for k in range(6))
        for q in range(1, 7):
            for tr in trees(q):
                assert abs(bth @ phi(tr, A) - th ** q / gamma(tr)) < 2e-12


def test_sosra_roessler_conditions():
    s = {k: float(v) for k, v in T["sosra"].items()}
    al = np.array([s["alpha1"], s["alpha2"], s["alpha3"]])
    b1 = np.array([s["beta11"], s["beta12"], s["beta13"]])
    b2 = np.array([s["beta21"], s["beta22"], s["beta23"]])
    A0 = np.array([[0, 0, 0], [s["A021"], 0, 0], [s["A031"], s["A032"], 0]])
    B0 = np.array([[0, 0, 0], [s["B021"], 0, 0], [s["B031"], s["B032"], 0]])
    c1 = np.array([s["c11"], s["c12"], s["c13"]])
    e = np.ones(3)
    conds = [al @ e - 1, b1 @ e - 1, b2 @ e, al @ B0 @ e - 1, al @ A0 @ e - 0.5, al @ (B0 @ e) ** 2 - 1.5,
             b1 @ c1 - 1, b2 @ c1 + 1]
    assert max(abs(x) for x in conds) < 1e-14


def test_generated_headers_match_json():
    for path in ("oracle/tableaus_gen.h", "differentialequations.jl_b200/csrc/kernels/tableaus_gen.cuh"):
        txt = open(os.path.join(ROOT, path)).read()
        for name, tab in T.items():
            for k, v in tab.items():
                assert f"#define B2T_{name.upper()}_{k} " in txt


def test_rodas_dense_weights_are_reproducible_and_consistent():
    """tools/rodas_dense.json (the dense-output weights of Rodas4 / Rodas5 / Rodas5P in the kernels' stage variables) is what
    tools/derive_rodas_dense.py derives from tools/tableaus.json, the generated headers carry the same numbers, and
    m_i(1) reproduces the step itself: y(t + h) = U_s + k_s, i.e. weights (a_s1, .., a_s,s-1, 1)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stored = json.load(open(os.path.join(root, "tools", "rodas_dense.json")))
    tabs = json.load(open(os.path.join(root, "tools", "tableaus.json")))
    hdr = open(os.path.join(root, "oracle", "tableaus_gen.h")).read()
    for name in ("rodas4", "rodas5", "rodas5p"):
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "derive_rodas_dense.py"), name], capture_output=True, text=True)
        fresh = json.loads(out.stdout)
        assert "max residual" in out.stderr and float(out.stderr.split("max residual")[1].split()[0]) < 1e-12
        m = [[float(v) for v in row] for row in stored[name]["m"]]
        for i, row in enumerate(m):
            for p_, v in enumerate(row):
                assert abs(v - fresh["m"][i][p_]) < 1e-11 * max(1.0, abs(v))
                assert f"#define B2T_{name.upper()}_H{i + 1}{p_ + 1} " in hdr
        s = stored[name]["stages"]
        last = [float(tabs[name][f"a{s - (2 if s == 8 else 1)}{j + 1}"]) for j in range(s - (3 if s == 8 else 2))]
        # chained rows: a_s = (a_{nexp}, 1, .., 1)
        want = last + [1.0] * (s - len(last))
        got = [sum(row) for row in m]
        assert max(abs(a - b) for a, b in zip(got, want)) < 1e-12, (got, want)
