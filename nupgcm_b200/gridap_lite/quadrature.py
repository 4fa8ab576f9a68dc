"""Simplex quadrature rules in barycentric coordinates.

Stands in for ``Measure(Ω, degree)`` / ``Measure(Γ, degree)`` with ``degree=4``
(reference ``src/meshes.jl:29-37``).  Each rule is ``(bary, w)`` with ``bary`` of shape
``(nq, d+1)`` and ``w`` summing to 1; multiply by the simplex measure.

The tetrahedral rule Gridap 0.20.3 picks for degree 4 is not pinned by any fixture of the
reference, so the rule is a parameter: ``"keast11"`` (default, exact to
degree 4, one negative weight), ``"keast15"`` (degree 5, used to bound the rule's influence).
"""
from __future__ import annotations

import itertools

import numpy as np


def _perms(vals):
    return sorted(set(itertools.permutations(vals)))


def _orbit_rule(orbits):
    pts, wts = [], []
    for vals, w in orbits:
        for p in _perms(vals):
            pts.append(p)
            wts.append(w)
    bary = np.array(pts, dtype=np.float64)
    w = np.array(wts, dtype=np.float64)
    if abs(w.sum() - 1.0) > 1e-9:
        raise AssertionError("quadrature weights do not sum to 1")
    return bary, w / w.sum()


def segment(degree: int = 4):
    """Gauss-Legendre on a segment (3 points: exact to degree 5)."""
    x, w = np.polynomial.legendre.leggauss(max(1, degree // 2 + 1))
    t = 0.5 * (x + 1.0)
    return np.stack([1.0 - t, t], axis=1), 0.5 * w


def triangle(degree: int = 4):
    """6-point degree-4 rule (Strang-Fix / Dunavant)."""
    if degree > 4:
        raise ValueError("triangle rules up to degree 4 only")
    a1, w1 = 0.445948490915965, 0.223381589678011
    a2, w2 = 0.091576213509771, 0.109951743655322
    return _orbit_rule([((a1, a1, 1 - 2 * a1), w1), ((a2, a2, 1 - 2 * a2), w2)])


def tetrahedron(rule: str = "keast11"):
    if rule == "keast4":      # degree 2
        a = 0.1381966011250105
        return _orbit_rule([((a, a, a, 1 - 3 * a), 0.25)])
    if rule == "keast11":     # degree 4
        a = 0.0714285714285714
        b = 0.399403576166799
        return _orbit_rule([
            ((0.25, 0.25, 0.25, 0.25), -0.01315555555555556 * 6),
            ((a, a, a, 1 - 3 * a), 0.007622222222222222 * 6),
            ((b, b, 0.5 - b, 0.5 - b), 0.02488888888888889 * 6),
        ])
    if rule == "keast15":     # degree 5, positive weights
        a = 1.0 / 3.0
        b = 0.0909090909090909
        c = 0.0665501535736643
        return _orbit_rule([
            ((0.25, 0.25, 0.25, 0.25), 0.0302836780970892 * 6),
            ((a, a, a, 1 - 3 * a), 0.00602678571428572 * 6),
            ((b, b, b, 1 - 3 * b), 0.0116452490860290 * 6),
            ((c, c, 0.5 - c, 0.5 - c), 0.0109491415613865 * 6),
        ])
    raise ValueError(f"unknown tetrahedral rule {rule!r}")


def simplex(dim: int, rule: str | None = None):
    if dim == 1:
        return segment()
    if dim == 2:
        return triangle()
    if dim == 3:
        return tetrahedron(rule or "keast11")
    raise ValueError(dim)
