"""Real Clebsch-Gordan (Wigner-3j) tables for the kernel generator (host side).

The per-path coupling tensor ``C^{l1 l2 l3}[i, j, k]`` that the reference obtains
from e3nn's ``o3.wigner_3j`` when ``TensorProductScatter`` constructs its
``o3.TensorProduct`` (``nequip/nn/_tp_scatter_base.py:24-31``).  e3nn's convention
(0.6.x ``e3nn/o3/_wigner.py``): SU(2) coefficients from the Racah sum, rotated to
the real basis with ``Q_l`` (m<0 sine-like, m>0 cosine-like, overall ``(-i)^l``),
Frobenius-normalised.  Only the sparse list of non-zeros is consumed by
``codegen.py``.
"""
from __future__ import annotations

from fractions import Fraction
from functools import lru_cache
from math import factorial, sqrt
from typing import List, Tuple


def _racah(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    """<j1 m1; j2 m2 | j3 m3> for integer spins."""
    if m1 + m2 != m3:
        return 0.0
    f = factorial
    pref = Fraction(
        (2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
        f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
    )
    lo = max(-j1 + j2 + m3, -j1 + m1, 0)
    hi = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)
    acc = Fraction(0)
    for v in range(lo, hi + 1):
        term = Fraction(
            f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
        )
        acc += -term if (v + j2 + m2) % 2 else term
    val = sqrt(float(pref)) * float(acc)
    return val


def _real_to_complex(l: int) -> List[List[complex]]:
    """Q_l[m_complex, m_real] (rows: complex m=-l..l, cols: real index)."""
    n = 2 * l + 1
    q = [[0j] * n for _ in range(n)]
    r = 1 / sqrt(2)
    for m in range(-l, 0):
        q[l + m][l - m] = r
        q[l + m][l + m] = -1j * r
    q[l][l] = 1
    for m in range(1, l + 1):
        s = -1.0 if m % 2 else 1.0
        q[l + m][l + m] = s * r
        q[l + m][l - m] = 1j * s * r
    ph = (-1j) ** l
    return [[ph * v for v in row] for row in q]


@lru_cache(maxsize=None)
def real_w3j(l1: int, l2: int, l3: int) -> Tuple[Tuple[Tuple[float, ...], ...], ...]:
    """Dense real tensor ``[2l1+1][2l2+1][2l3+1]`` with unit Frobenius norm."""
    if not abs(l1 - l2) <= l3 <= l1 + l2:
        raise ValueError(f"({l1},{l2},{l3}) violates the triangle rule")
    n1, n2, n3 = 2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    # su2[a][b][c] with complex-basis indices
    su2 = {}
    for a in range(n1):
        for b in range(n2):
            m3 = (a - l1) + (b - l2)
            if abs(m3) <= l3:
                v = _racah(l1, a - l1, l2, b - l2, l3, m3)
                if v != 0.0:
                    su2[(a, b, m3 + l3)] = v
    out = [[[0.0] * n3 for _ in range(n2)] for _ in range(n1)]
    nrm = 0.0
    for i in range(n1):
        for j in range(n2):
            for k in range(n3):
                s = 0j
                for (a, b, c), v in su2.items():
                    qa, qb = Q1[a][i], Q2[b][j]
                    if qa == 0 or qb == 0:
                        continue
                    # conj(Q3^T)[k, c] = conj(Q3[c][k])
                    qc = Q3[c][k].conjugate()
                    if qc == 0:
                        continue
                    s += qa * qb * qc * v
                if abs(s.imag) > 1e-12:
                    raise AssertionError("real w3j has an imaginary part")
                out[i][j][k] = s.real
                nrm += s.real * s.real
    nrm = sqrt(nrm)
    return tuple(
        tuple(tuple((v / nrm if abs(v / nrm) > 1e-14 else 0.0) for v in row) for row in plane)
        for plane in out
    )


def sparse_w3j(l1: int, l2: int, l3: int) -> List[Tuple[int, int, int, float]]:
    """Non-zeros ``(i, j, k, value)``."""
    C = real_w3j(l1, l2, l3)
    return [
        (i, j, k, C[i][j][k])
        for i in range(2 * l1 + 1)
        for j in range(2 * l2 + 1)
        for k in range(2 * l3 + 1)
        if C[i][j][k] != 0.0
    ]
