"""Rooted-tree order conditions for Runge-Kutta tableaux (Butcher theory).

Used by tests/test_tableaus.py (the tables in tools/tableaus.json must satisfy
their nominal order) and by tools/derive_vern7_dense.py (dense-output weights).

A tree is a sorted tuple of child trees; the leaf is ().
  order |t|      = 1 + sum |child|
  density gamma  = |t| * prod gamma(child)
  elementary weight vector Phi(t)_i = prod_child (A @ Phi(child))_i , Phi(leaf)=1
Order condition:  b . Phi(t) = 1/gamma(t)  for all |t| <= p.
Continuous:       b(theta) . Phi(t) = theta^|t| / gamma(t).
"""
from functools import lru_cache
from itertools import combinations_with_replacement

import numpy as np


@lru_cache(maxsize=None)
def trees(n):
    """All rooted trees with n vertices."""
    if n == 1:
        return ((),)
    out = set()
    for part in _partitions(n - 1):
        # part: tuple of child orders (non-increasing)
        pools = [trees(k) for k in part]
        out.update(_products(part, pools))
    return tuple(sorted(out))


def _partitions(n, maxpart=None):
    if maxpart is None:
        maxpart = n
    if n == 0:
        yield ()
        return
    for k in range(min(n, maxpart), 0, -1):
        for rest in _partitions(n - k, k):
            yield (k,) + rest


def _products(part, pools):
    # group equal orders so that children multisets are not duplicated
    groups = {}
    for k in part:
        groups[k] = groups.get(k, 0) + 1
    choices = [[]]
    for k, mult in groups.items():
        new = []
        for combo in combinations_with_replacement(trees(k), mult):
            for c in choices:
                new.append(c + list(combo))
        choices = new
    return {tuple(sorted(c)) for c in choices}


def order(t):
    return 1 + sum(order(c) for c in t)


def gamma(t):
    g = order(t)
    for c in t:
        g *= gamma(c)
    return g


def phi(t, A):
    """Elementary weight vector (one entry per stage)."""
    s = A.shape[0]
    out = np.ones(s, dtype=A.dtype)
    for c in t:
        out = out * (A @ phi(c, A))
    return out


def order_residuals(A, b, p):
    """max |b.Phi - 1/gamma| per order 1..p."""
    res = {}
    for q in range(1, p + 1):
        r = 0.0
        for t in trees(q):
            r = max(r, abs(float(b @ phi(t, A)) - 1.0 / gamma(t)))
        res[q] = r
    return res
