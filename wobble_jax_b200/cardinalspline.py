"""The cardinal B-spline basis with the reference's names (``jabble/cardinalspline.py``).

``CardinalSplineKernel(n)`` is the n!-scaled Irwin-Hall basis function of degree n.  Its
piecewise-polynomial table ``alphas`` is host data (a constructor-time constant; the CUDA kernels
carry the same polynomials written out per degree).  Evaluating it, ``kernel(t)``, runs on the GPU:
B(t) is the cardinal spline whose only non-zero control point sits on the knot at 0, which is what
the forward kernel evaluates (``jb_plan_forward``) -- there is no host evaluation path.
"""

import math

import numpy as np


def _alpha_table(n):
    """alphas[j, k] = coefficient of s^j on piece k (s = t + (n+1)/2 in [k, k+1)) of n! times the
    density of a sum of n+1 uniforms: sum_{i <= k} (-1)^i C(n+1, i) (s - i)^n, expanded in s."""
    return np.array([[sum((-1) ** i * math.comb(n + 1, i) * math.comb(n, j) * float(-i) ** (n - j)
                          for i in range(k + 1)) for k in range(n + 1)] for j in range(n + 1)], dtype=np.float64)


def _alpha_recursion(j, k, n):
    """cardinalspline.py:6-18 -- entry (j, k) of the table of the basis with n pieces (degree n - 1).
    The reference recurses over k; this is the closed form of the same sum."""
    return float(_alpha_table(n - 1)[j, k])


class CardinalSplineKernel:
    """cardinalspline.py:21-55 -- ``alphas`` (n+1, n+1) and ``kernel(t)`` = B(t), zero outside
    ``|t| < (n+1)/2``.  Degrees 1..3 evaluate on the CUDA path; others raise ``ValueError``."""

    def __init__(self, n):
        self.n = int(n)
        self.alphas = _alpha_table(self.n)

    def __call__(self, x, *args):
        from . import model as M
        t = np.asarray(x, dtype=np.float64)
        flat = t.reshape(-1)
        if flat.size == 0:
            return np.zeros(t.shape)
        if not np.all(np.isfinite(flat)):
            raise ValueError("CardinalSplineKernel: non-finite argument")
        half = (self.n + 1) / 2.0
        # points outside the support are zero by definition (cond1 * cond2 of the reference) and
        # need no knots: only the inside ones go to the device
        inside = np.abs(flat) < half
        out = np.zeros(flat.shape)
        if inside.any():
            m = self.n + 3                                   # knots -m .. m around the impulse at 0
            xp = np.arange(-m, m + 1, dtype=np.float64)
            ap = np.zeros(xp.shape)
            ap[m] = 1.0
            out[inside] = M.cardinal_vmap_model(flat[inside], xp, ap, self, half)
        return out.reshape(t.shape)
