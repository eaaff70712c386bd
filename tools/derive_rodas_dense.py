"""Derive an order-4 dense output for Rodas5P (and Rodas5 / Rodas4: order 3) from the Rosenbrock order conditions
(Hairer & Wanner IV.7): sum_i b_i(theta) Phi_i(t) = theta^rho(t) P_t(gamma / theta) for the eight trees of order <= 4.
The tableaus in tools/tableaus.json are in the transformed (implementation) form; they are converted back to
(alpha, Gamma, b), the discrete conditions are checked at theta = 1, the polynomial weights b_i(theta) are solved for and
returned in the transformed variables: y(t + theta h) = y + sum_i m_i(theta) k_i with the kernel's k_i.
python tools/derive_rodas_dense.py [rodas5p|rodas5|rodas4]  -> JSON with m[i][p] (coefficient of theta^(p+1))"""
import json, os, sys
import numpy as np
import mpmath as mp

mp.mp.dps = 40
name = sys.argv[1] if len(sys.argv) > 1 else "rodas5p"
T = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tableaus.json")))[name]
s = 8 if "C81" in T else 6
g = mp.mpf(T["gamma"])
A = mp.zeros(s, s); C = mp.zeros(s, s)
nexp = s - 2 if s == 8 else s - 1   # rows with their own a-coefficients; the rest chain: a_i = (a_{i-1}, 1)
for i in range(1, nexp):
    for j in range(i):
        A[i, j] = mp.mpf(T[f"a{i+1}{j+1}"])
for i in range(nexp, s):
    for j in range(i - 1):
        A[i, j] = A[i - 1, j]
    A[i, i - 1] = 1
for i in range(1, s):
    for j in range(i):
        C[i, j] = mp.mpf(T[f"C{i+1}{j+1}"])
Ginv = mp.eye(s) / g - C                 # C = diag(1/gamma) - Gamma^{-1}
Gam = Ginv ** -1
alpha = A * Gam                          # a = alpha Gamma^{-1}
m = mp.matrix([A[s - 1, j] for j in range(s - 1)] + [1])   # y1 = U_s + k_s
b = (m.T * Gam).T
beta = alpha + Gam
for i in range(s):
    beta[i, i] = 0
    for j in range(i + 1, s):
        beta[i, j] = 0; alpha[i, j] = 0
ai = mp.matrix([sum(alpha[i, j] for j in range(i)) for i in range(s)])
bp = mp.matrix([sum(beta[i, j] for j in range(i)) for i in range(s)])
one = mp.matrix([1] * s)
Phi = [one, bp, mp.matrix([ai[i] ** 2 for i in range(s)]), beta * bp, mp.matrix([ai[i] ** 3 for i in range(s)]),
       mp.matrix([ai[i] * sum(alpha[i, j] * bp[j] for j in range(s)) for i in range(s)]),
       beta * mp.matrix([ai[i] ** 2 for i in range(s)]), beta * (beta * bp)]
rho = [1, 2, 3, 3, 4, 4, 4, 4]
# P_t(gamma) as coefficient lists in gamma: P = sum_q pc[q] gamma^q
P = [[1], [mp.mpf(1) / 2, -1], [mp.mpf(1) / 3], [mp.mpf(1) / 6, -1, 1], [mp.mpf(1) / 4], [mp.mpf(1) / 8, -mp.mpf(1) / 3],
     [mp.mpf(1) / 12, -mp.mpf(1) / 3], [mp.mpf(1) / 24, -mp.mpf(1) / 2, mp.mpf(3) / 2, -1]]
print("discrete conditions at theta = 1 (residuals):", file=sys.stderr)
for t in range(8):
    lhs = sum(b[i] * Phi[t][i] for i in range(s))
    rhs = sum(pc * g ** q for q, pc in enumerate(P[t]))
    print(f"  tree {t} rho {rho[t]}: {mp.nstr(lhs - rhs, 5)}", file=sys.stderr)
print("c_i check (alpha row sums vs c):", [mp.nstr(ai[i], 8) for i in range(s)], file=sys.stderr)
ntree = 8 if s == 8 else 4             # Rodas4 (6 stages): order-3 dense output from the four trees of order <= 3
deg = 4 if s == 8 else 3
# unknowns b_i(theta) = sum_{p=1..deg} B[i][p] theta^p.  theta^rho P(gamma/theta) = sum_q pc[q] gamma^q theta^(rho-q).
# The tree matrix is rank-deficient for these tableaus (c_6 = c_7 = c_8 = 1 and the chained rows), so the conditions leave
# a family of solutions: all conditions for all powers plus b_i(1) = b_i are solved together in the minimum-norm sense
# (numpy SVD) and the residual is checked.
# in the transformed variables: b = m Gamma, so sum_i b_i Phi_i = sum_j m_j (Gamma Phi)_j; the minimum norm is taken over
# the coefficients of m_j(theta), the numbers the kernels multiply the stage values with
Psi = [Gam * Phi[t] for t in range(8)]
nun = s * deg
rows, rhsv = [], []
for t in range(ntree):
    for p in range(1, deg + 1):
        row = np.zeros(nun)
        for i in range(s):
            row[i * deg + (p - 1)] = float(Psi[t][i])
        r = mp.mpf(0)
        for q, pc in enumerate(P[t]):
            if rho[t] - q == p:
                r += pc * g ** q
        rows.append(row); rhsv.append(float(r))
for i in range(s):
    row = np.zeros(nun)
    row[i * deg:(i + 1) * deg] = 1.0
    rows.append(row); rhsv.append(float(m[i]))
Mx, rv = np.array(rows), np.array(rhsv)
sol, res, rank, sv = np.linalg.lstsq(Mx, rv, rcond=1e-12)
print("rank", rank, "of", Mx.shape, "max residual", float(np.max(np.abs(Mx @ sol - rv))), file=sys.stderr)
Mt = mp.zeros(s, deg)
for i in range(s):
    for p in range(deg):
        Mt[i, p] = mp.mpf(float(sol[i * deg + p]))
out = {"name": name, "stages": s, "degree": deg, "m": [[float(Mt[i, p]) for p in range(deg)] for i in range(s)],
       "m_str": [[mp.nstr(Mt[i, p], 20) for p in range(deg)] for i in range(s)]}
print(json.dumps(out, indent=1))
