"""Real-basis Wigner-3j tensors and Wigner-D matrices (TEST INFRASTRUCTURE ONLY).

Restates the published algorithm of ``e3nn==0.6.x`` ``e3nn/o3/_wigner.py``
(``_su2_clebsch_gordan``, ``change_basis_real_to_complex``,
``_so3_clebsch_gordan``; dependency pinned at reference ``pyproject.toml:21``,
NOT vendored).  Reference call sites that consume these numbers:
``nequip/nn/_tp_scatter_base.py:24-31`` (o3.TensorProduct builds its w3j
buffers from ``o3.wigner_3j``) and SURVEY.md Appendix A.2.

    C = Re( sum Q1[i,j] Q2[k,l] conj(Q3^T)[m,n] CG_su2[i,k,n] ),  C <- C/||C||_F
"""
from fractions import Fraction
from functools import lru_cache
from math import factorial

import numpy as np


def _f(n: int) -> int:
    if n < 0:
        raise ValueError
    return factorial(n)


def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    """<j1 m1 j2 m2 | j3 m3> via the Racah sum (integer j only are needed here).
    Exact rational arithmetic under the square root."""
    if m3 != m1 + m2:
        return 0.0
    vmin = max(-j1 + j2 + m3, -j1 + m1, 0)
    vmax = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)
    C = Fraction(
        (2 * j3 + 1) * _f(j3 + j1 - j2) * _f(j3 - j1 + j2) * _f(j1 + j2 - j3) * _f(j3 + m3) * _f(j3 - m3),
        _f(j1 + j2 + j3 + 1) * _f(j1 - m1) * _f(j1 + m1) * _f(j2 - m2) * _f(j2 + m2),
    )
    S = Fraction(0)
    for v in range(vmin, vmax + 1):
        S += Fraction(
            (-1) ** (v + j2 + m2) * _f(j2 + j3 + m1 - v) * _f(j1 - m1 + v),
            _f(v) * _f(j3 - j1 + j2 - v) * _f(j3 + m3 - v) * _f(v + j1 - j2 - m3),
        )
    sign = -1.0 if S < 0 else 1.0
    return sign * float(np.sqrt(float(C * S * S)))


def su2_cg(j1: int, j2: int, j3: int) -> np.ndarray:
    mat = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1))
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def change_basis_real_to_complex(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = 1 / np.sqrt(2)
        q[l + m, l - abs(m)] = -1j / np.sqrt(2)
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m / np.sqrt(2)
        q[l + m, l - abs(m)] = 1j * (-1) ** m / np.sqrt(2)
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def _w3j_cached(l1: int, l2: int, l3: int):
    Q1 = change_basis_real_to_complex(l1)
    Q2 = change_basis_real_to_complex(l2)
    Q3 = change_basis_real_to_complex(l3)
    C = su2_cg(l1, l2, l3).astype(np.complex128)
    C = np.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, np.conj(Q3.T), C)
    assert np.abs(C.imag).max() < 1e-12
    C = C.real
    C = C / np.linalg.norm(C)
    C[np.abs(C) < 1e-14] = 0.0
    C.setflags(write=False)
    return C


def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real w3j tensor ``[2l1+1, 2l2+1, 2l3+1]``, Frobenius norm 1."""
    assert abs(l1 - l2) <= l3 <= l1 + l2
    return _w3j_cached(l1, l2, l3)


# ---------------------------------------------------------------------------
# Wigner-D in the same real basis (for equivariance checks only)
# ---------------------------------------------------------------------------
def su2_generators(j: int) -> np.ndarray:
    m = np.arange(-j, j)
    raising = np.diag(-np.sqrt(j * (j + 1) - m * (m + 1)), k=-1)
    m = np.arange(-j + 1, j + 1)
    lowering = np.diag(np.sqrt(j * (j + 1) - m * (m - 1)), k=1)
    m = np.arange(-j, j + 1)
    return np.stack(
        [
            0.5 * (raising + lowering),
            np.diag(1j * m),
            -0.5j * (raising - lowering),
        ],
        axis=0,
    )


def so3_generators(l: int) -> np.ndarray:
    X = su2_generators(l)
    Q = change_basis_real_to_complex(l)
    X = np.conj(Q.T) @ X @ Q
    assert np.abs(X.imag).max() < 1e-12
    return X.real


def _expm(A: np.ndarray) -> np.ndarray:
    from scipy.linalg import expm

    return expm(A)


def wigner_D(l: int, alpha: float, beta: float, gamma: float) -> np.ndarray:
    """e3nn convention: D = exp(alpha X_y) exp(beta X_x) exp(gamma X_y)
    (YXY Euler angles, y is the polar axis)."""
    X = so3_generators(l)
    return _expm(alpha * X[1]) @ _expm(beta * X[0]) @ _expm(gamma * X[1])


def rotation_matrix(alpha: float, beta: float, gamma: float) -> np.ndarray:
    """Cartesian rotation for the same angles; equals ``wigner_D(1, ...)``."""
    return wigner_D(1, alpha, beta, gamma)
