"""Real spherical harmonics in e3nn's convention (``o3.SphericalHarmonics`` with
``normalize=True, normalization='component'``; reference call site
``sevenn/nn/edge_embedding.py:164-185``): the polar axis is y, m runs -l..l,
sum_m Y_lm(u)^2 = 2l+1 on the unit sphere, Y_0 = 1.

The closed forms for l <= 3 are restated from SURVEY Appendix A.3 (verified there against
the reference's golden vectors) as sympy polynomials in the unit vector (x, y, z); the
numpy evaluator, the oracle and the CUDA code generator (``csrc/gen_kernels.py``) are all
derived from this one table.  ``tests/test_cg.py`` checks the component normalisation and
that w3j(l1,l2,l1+l2) . Y_l1 . Y_l2 is parallel to +Y_(l1+l2) -- the property that fixes
ordering and signs.
"""
from __future__ import annotations

import functools
from typing import List

import numpy as np
import sympy as sp

X, Y, Z = sp.symbols('x y z', real=True)
LMAX_SUPPORTED = 3


@functools.lru_cache(maxsize=None)
def sh_polynomials(lmax: int) -> List[sp.Expr]:
    """List of (lmax+1)^2 sympy polynomials in the *unit* vector components."""
    if lmax > LMAX_SUPPORTED:
        raise NotImplementedError(f'lmax {lmax} > {LMAX_SUPPORTED}')
    x, y, z = X, Y, Z
    s3, s5, s7 = sp.sqrt(3), sp.sqrt(5), sp.sqrt(7)
    out: List[sp.Expr] = [sp.Integer(1)]
    if lmax >= 1:
        out += [s3 * x, s3 * y, s3 * z]
    s20 = s3 * x * z
    s24 = (s3 / 2) * (z * z - x * x)
    if lmax >= 2:
        out += [s5 * s20, s5 * s3 * x * y, s5 * (y * y - (x * x + z * z) / 2),
                s5 * s3 * y * z, s5 * s24]
    if lmax >= 3:
        c = sp.sqrt(30) / 6
        q = 4 * y * y - x * x - z * z
        out += [s7 * c * (s20 * z + s24 * x),
                s7 * s5 * s20 * y,
                s7 * (sp.sqrt(6) / 4) * q * x,
                s7 * sp.Rational(1, 2) * y * (2 * y * y - 3 * (x * x + z * z)),
                s7 * (sp.sqrt(6) / 4) * z * q,
                s7 * s5 * s24 * y,
                s7 * c * (s24 * z - s20 * x)]
    return [sp.expand(e) for e in out]


@functools.lru_cache(maxsize=None)
def _sh_lambda(lmax: int):
    return sp.lambdify((X, Y, Z), sh_polynomials(lmax), 'numpy')


def spherical_harmonics(lmax: int, vec: np.ndarray) -> np.ndarray:
    """vec [..., 3] (not necessarily unit) -> [..., (lmax+1)^2]; the vector is normalised first."""
    vec = np.asarray(vec, dtype=np.float64)
    u = vec / np.linalg.norm(vec, axis=-1, keepdims=True)
    vals = _sh_lambda(lmax)(u[..., 0], u[..., 1], u[..., 2])
    vals = [np.broadcast_to(np.asarray(v, dtype=np.float64), u.shape[:-1]) for v in vals]
    return np.stack(vals, axis=-1)
