"""Derive an order-6 dense-output table for Vern7 from its extra stages.

SURVEY.md B.3: the recalled extra-stage rows (11..16) satisfy stage-order 6 but
not the order-7 continuous conditions, so upstream's order-7 interpolant cannot
be reproduced; instead solve the order<=6 continuous conditions

    sum_i r[i,k] Phi_i(tau) = [k == |tau|] / gamma(tau)      (all |tau| <= 6)

for polynomial weights b_i(theta) = sum_{k=1..6} r[i,k] theta^k over the stages
{1,4,5,6,7,8,9,11,...,16}, with the continuity constraint b_i(1) = b_i (so the
interpolant meets u_new exactly at theta=1), minimum-norm solution, in 50-digit
arithmetic.  Writes tools/vern7_dense.json.  DEVIATION from upstream (order 6
instead of order 7) recorded in DESIGN.md.
"""
import json
import os
import sys

import mpmath as mp

sys.path.insert(0, os.path.dirname(__file__))
from rk_trees import trees, gamma, order  # noqa: E402

mp.mp.dps = 50
HERE = os.path.dirname(__file__)
T = json.load(open(os.path.join(HERE, "tableaus.json")))

S = 16
A = mp.zeros(S, S)
for src in ("vern7", "vern7_extra"):
    for k, v in T[src].items():
        if k.startswith("a"):
            A[int(k[1:3]) - 1, int(k[3:5]) - 1] = mp.mpf(v)
b = [mp.mpf(0)] * S
for k, v in T["vern7"].items():
    if k.startswith("b") and not k.startswith("btilde"):
        b[int(k[1:]) - 1] = mp.mpf(v)

stages = [1, 4, 5, 6, 7, 8, 9, 11, 12, 13, 14, 15, 16]
P = 6


def phi(t):
    out = mp.ones(S, 1)
    for c in t:
        w = A * phi(c)
        out = mp.matrix([out[i] * w[i] for i in range(S)])
    return out


all_trees = [t for q in range(1, P + 1) for t in trees(q)]
Phi = [phi(t) for t in all_trees]

ns = len(stages)
nunk = ns * P  # r[i,k] at index i*P + (k-1)
rows = []
rhs = []
for t, ph in zip(all_trees, Phi):
    for k in range(1, P + 1):
        row = [mp.mpf(0)] * nunk
        for ii, s in enumerate(stages):
            row[ii * P + (k - 1)] = ph[s - 1]
        rows.append(row)
        rhs.append(mp.mpf(1) / gamma(t) if order(t) == k else mp.mpf(0))
for ii, s in enumerate(stages):  # continuity at theta = 1
    row = [mp.mpf(0)] * nunk
    for k in range(P):
        row[ii * P + k] = mp.mpf(1)
    rows.append(row)
    rhs.append(b[s - 1])

M = mp.matrix(rows)
y = mp.matrix(rhs)
# minimum-norm least squares through the SVD pseudo-inverse
U, sv, V = mp.svd_r(M)
tol = sv[0] * mp.mpf(10) ** (-25)
x = mp.zeros(nunk, 1)
rank = 0
for j in range(len(sv)):
    if sv[j] > tol:
        rank += 1
        coef = sum(U[i, j] * y[i] for i in range(M.rows)) / sv[j]
        for i in range(nunk):
            x[i] += coef * V[j, i]
res = M * x - y
maxres = max(abs(res[i]) for i in range(res.rows))
print("unknowns", nunk, "equations", M.rows, "rank", rank, "max residual", mp.nstr(maxres, 5))

out = {"stages": stages, "powers": list(range(1, P + 1)), "r": {}}
for ii, s in enumerate(stages):
    out["r"][str(s)] = [mp.nstr(x[ii * P + k], 20) for k in range(P)]
out["max_residual"] = mp.nstr(maxres, 5)
json.dump(out, open(os.path.join(HERE, "vern7_dense.json"), "w"), indent=1)
