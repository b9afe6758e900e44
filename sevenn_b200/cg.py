"""Real-basis Wigner-3j (Clebsch-Gordan) tensors in the e3nn convention.

The reference gets these from the third-party package ``e3nn`` (``o3.wigner_3j``,
unpinned ``e3nn>=0.5.0``, reference ``pyproject.toml:24``; call sites
``sevenn/nn/convolution.py:100`` via ``o3.TensorProduct`` and
``sevenn/nn/cue_helper.py:36``).  e3nn is not vendored in the reference tree, so this
module restates its published algorithm from the mathematics:

  1. SU(2) Clebsch-Gordan coefficients <j1 m1 j2 m2 | j3 m3> (Racah's formula),
  2. change of basis from complex to e3nn's real spherical-harmonic basis
     (m ordered -l..l, the polar axis is y), including the (-i)^l phase that makes
     the coupling tensor real,
  3. Frobenius normalisation (sum of squares == 1).

Parity with the reference is pinned in ``tests/test_cg.py`` against the Wigner-3j
buffers that the shipped SevenNet checkpoints carry
(``*_convolution.convolution._compiled_main_left_right._w3j_l1_l2_l3``), exported to
``tests/golden/w3j_reference.npz`` by ``tools/make_golden.py``.
"""
from __future__ import annotations

import functools
from fractions import Fraction
from math import factorial

import numpy as np


def _su2_cg_coeff(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    """<j1 m1; j2 m2 | j3 m3> for integer spins (Racah's closed form)."""
    if m3 != m1 + m2:
        return 0.0
    vmin = max(-j1 + j2 + m3, -j1 + m1, 0)
    vmax = min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3)

    def f(n: int) -> int:
        return factorial(n)

    c = Fraction(
        (2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3)
        * f(j3 + m3) * f(j3 - m3),
        f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2),
    )
    s = Fraction(0)
    for v in range(vmin, vmax + 1):
        s += Fraction(
            (-1) ** (v + j2 + m2) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
            f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3),
        )
    return float(np.sqrt(float(c)) * float(s))


def _su2_cg(j1: int, j2: int, j3: int) -> np.ndarray:
    out = np.zeros((2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1), dtype=np.float64)
    if not (abs(j1 - j2) <= j3 <= j1 + j2):
        return out
    for m1 in range(-j1, j1 + 1):
        for m2 in range(-j2, j2 + 1):
            m3 = m1 + m2
            if abs(m3) <= j3:
                out[j1 + m1, j2 + m2, j3 + m3] = _su2_cg_coeff(j1, m1, j2, m2, j3, m3)
    return out


def _real_to_complex(l: int) -> np.ndarray:
    """Unitary q with  Y^complex_m = sum_m' q[m, m'] Y^real_m'  (e3nn phase choice)."""
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    s2 = 1.0 / np.sqrt(2.0)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s2
        q[l + m, l - abs(m)] = -1j * s2
    q[l, l] = 1.0
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s2
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s2
    return (-1j) ** l * q


@functools.lru_cache(maxsize=None)
def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Real-basis coupling tensor C[i, j, k], shape (2l1+1, 2l2+1, 2l3+1), ||C||_F = 1."""
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        raise ValueError(f'({l1},{l2},{l3}) violates the triangle rule')
    q1, q2, q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    c = _su2_cg(l1, l2, l3).astype(np.complex128)
    c = np.einsum('ij,kl,mn,ikn->jlm', q1, q2, np.conj(q3.T), c)
    assert np.abs(c.imag).max() < 1e-9, 'coupling tensor is not real'
    c = c.real.copy()
    c[np.abs(c) < 1e-14] = 0.0
    c /= np.linalg.norm(c)
    c.setflags(write=False)
    return c


def tp_path_coefficients(l1: int, l2: int, l3: int) -> np.ndarray:
    """sqrt(2l3+1) * w3j: the per-path constant tensor of the 'uvu' tensor product with
    e3nn's default ``irrep_normalization='component'``, ``path_normalization='element'``
    and one path per output slot."""
    return np.sqrt(2 * l3 + 1) * wigner_3j(l1, l2, l3)
