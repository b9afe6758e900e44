"""CPU ORACLE for the D3 dispersion correction -- test infrastructure, NOT product code.

A numpy fp64 restatement of the reference's CUDA D3 (``sevenn/pair_e3gnn/pair_d3_for_ase.cu``): all atom
pairs x all lattice translations inside the cutoffs, exactly the sums the reference kernels take (the
reference evaluates them in fp32 with fp64 accumulators; this oracle is fp64 throughout).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` may import it.

Parity is PINNED against the reference's own golden values (``tests/unit_tests/test_calculator.py:192-238``:
NaCl primitive cell and an H2O molecule, PBE / Becke-Johnson damping) in ``tests/test_d3_oracle.py``.

Reference lines followed:
  units, wrapping of positions into the cell     pair_d3_for_ase.cu:873-978, 1170-1216   (bohr; a -= floor(a))
  lattice repetitions per cutoff                 :979-1003
  coordination number                            :1004-1058   CN_i = sum 1/(1+exp(-K1 (rcov_i+rcov_j)/r + K1)), r^2 <= cnthr
  C6_ij(CN_i, CN_j) and its CN derivatives       :765-845     Gaussian (K3 = -4) interpolation of the reference C6
  pair energy / force, BJ and zero damping       :1534-1745, 1263-1496
  CN chain-rule force                            :1797-1962
  energy / force / virial units                  :1986-2006
  functional parameters                          :394-631     (weights/d3_params.npz, tools/convert_d3_params.py)
"""
from __future__ import annotations

import json
import os

import numpy as np

AU_TO_ANG = 0.52917726
AU_TO_EV = 27.21138505
K1, K3 = 16.0, -4.0
_PARAMS = None


def d3_params():
    global _PARAMS
    if _PARAMS is None:
        f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'weights', 'd3_params.npz'))
        p = {k: f[k] for k in ('r0ab', 'c6ref', 'cnref', 'mxc', 'r2r4', 'rcov')}
        p['functionals'] = json.loads(bytes(f['functionals']).decode())
        _PARAMS = p
    return _PARAMS


def damping_parameters(damping: str, functional: str):
    """``setfuncpar`` (:608-631): s6, s8 = s18, a1 = rs6, a2 = rs8 = rs18, alp6 = alp, alp8 = alp + 2."""
    p = d3_params()['functionals'][damping][functional]
    return dict(s6=p['s6'], s8=p['s18'], a1=p['rs6'], a2=p['rs18'], alp6=p['alp'], alp8=p['alp'] + 2.0)


def _translations(lat, cutoff, pbc):
    """all lattice translations n1 a1 + n2 a2 + n3 a3, |n_k| <= rep_k (:979-1003)"""
    a1, a2, a3 = lat
    rep = []
    for (u, v, w) in ((a2, a3, a1), (a3, a1, a2), (a1, a2, a3)):
        c = np.cross(u, v)
        h = abs(np.dot(c, w)) / np.linalg.norm(c)
        rep.append(int(abs(cutoff / h)) + 1)
    rep = [r if p else 0 for r, p in zip(rep, pbc)]
    g = np.stack(np.meshgrid(*[np.arange(-r, r + 1) for r in rep], indexing='ij'), -1).reshape(-1, 3)
    return g @ lat, g


def d3_reference(numbers, positions, cell, pbc=(True, True, True), damping='damp_bj', functional='pbe',
                 vdw_cutoff=9000.0, cn_cutoff=1600.0, mimic_fp32=False):
    """Energy (eV), forces (eV/A), virial sum_pairs f (x) r (eV, 3x3; the reference's ``sigma``) of the D3
    correction.  ``cell`` rows are lattice vectors in Angstrom (any orientation).

    ``mimic_fp32``: round to float32 where the reference kernels hold floats and the result is sensitive
    to it -- positions and translations (:1170-1216, ``float **x``), the per-pair CN terms (:1031-1034) and,
    above all, the coordination numbers entering the Gaussian C6 weights (``const float cni = cn[iat]``,
    :780-783: d ln w / d CN = 8 (CN - CN_ref) ~ 80 for highly coordinated atoms, so the float rounding of CN
    moves C6 by ~1e-5 relative).  Used to pin the reference's golden numbers to their own tolerance."""
    f32 = (lambda v: np.asarray(v, dtype=np.float32).astype(np.float64)) if mimic_fp32 else (lambda v: v)
    P = d3_params()
    z = np.asarray(numbers, dtype=np.int64) - 1
    lat = np.asarray(cell, dtype=np.float64) / AU_TO_ANG
    x = np.asarray(positions, dtype=np.float64) / AU_TO_ANG
    frac = x @ np.linalg.inv(lat)
    frac -= np.floor(frac)                                   # :1206 (all directions, as the reference does)
    x = f32(frac @ lat)
    n = len(z)
    dp = damping_parameters(damping, functional)
    rcov, r2r4 = P['rcov'][z], P['r2r4'][z]

    def pairs(cut2, strict):
        tau, g = _translations(lat, np.sqrt(cut2), pbc)
        tau = f32(tau)
        d = x[None, :, None, :] - x[:, None, None, :] + tau[None, None, :, :]        # [i, j, t, 3] = x_j - x_i + tau
        r2 = (d ** 2).sum(-1)
        self_img = (np.arange(n)[:, None, None] == np.arange(n)[None, :, None]) & (np.abs(g).sum(1) == 0)[None, None, :]
        m = ((r2 < cut2) if strict else (r2 <= cut2)) & ~self_img
        i, j, t = np.nonzero(m)
        return i, j, d[i, j, t], np.sqrt(r2[i, j, t])

    # coordination numbers (every ordered pair i <- j counts once for i)
    i, j, d, r = pairs(cn_cutoff, False)
    rc = rcov[i] + rcov[j]
    cn = np.zeros(n)
    np.add.at(cn, i, f32(1.0 / (1.0 + np.exp(-K1 * (rc / r - 1.0)))))
    cn_exact = cn
    cn = f32(cn)

    # C6_ij and derivatives for all atom pairs (:765-845)
    cnr, mxc = f32(P['cnref'][z]), P['mxc'][z]              # [n,5], [n]
    valid = np.arange(5)[None, :] < mxc[:, None]
    w = np.where(valid, np.exp(K3 * f32((cnr - cn[:, None]) ** 2)), 0.0)              # [n,5]
    dw = w * 2.0 * K3 * (cn[:, None] - cnr)
    c6r = f32(P['c6ref'][z[:, None], z[None, :]])           # [n,n,5,5]
    num = np.einsum('ijab,ia,jb->ij', c6r, w, w)
    den = np.einsum('ia,jb->ij', w, w)
    dnum_i = np.einsum('ijab,ia,jb->ij', c6r, dw, w)
    dden_i = np.einsum('ia,jb->ij', dw, w)
    dnum_j = np.einsum('ijab,ia,jb->ij', c6r, w, dw)
    dden_j = np.einsum('ia,jb->ij', w, dw)
    ok = den > 1e-99                                          # :824-844: else the C6 of the nearest reference, no derivative
    sden = np.where(ok, den, 1.0)
    near = np.argmin(np.where(valid, (cnr - cn[:, None]) ** 2, np.inf), axis=1)
    c6_near = c6r[np.arange(n)[:, None], np.arange(n)[None, :], near[:, None], near[None, :]]
    c6 = np.where(ok, num / sden, c6_near)
    dc6_i = np.where(ok, (dnum_i - c6 * dden_i) / sden, 0.0)
    dc6_j = np.where(ok, (dnum_j - c6 * dden_j) / sden, 0.0)

    # pair terms inside the vdW cutoff; ordered pairs, each carrying half of the pair energy
    i, j, d, r = pairs(vdw_cutoff, False)
    C6 = c6[i, j]
    if damping == 'damp_bj':
        r42x3 = 3.0 * r2r4[i] * r2r4[j]
        R0 = dp['a1'] * np.sqrt(r42x3) + dp['a2']
        t6, t8 = 1.0 / (r ** 6 + R0 ** 6), 1.0 / (r ** 8 + R0 ** 8)
        g = dp['s6'] * t6 + dp['s8'] * r42x3 * t8                                       # E_pair = -C6 g
        dg = -(6.0 * dp['s6'] * r ** 5 * t6 ** 2 + 8.0 * dp['s8'] * r42x3 * r ** 7 * t8 ** 2)
    elif damping == 'damp_zero':
        r0 = P['r0ab'][z[i], z[j]] / AU_TO_ANG
        r42 = r2r4[i] * r2r4[j]
        t6 = (dp['a1'] * r0 / r) ** dp['alp6']
        t8 = (dp['a2'] * r0 / r) ** dp['alp8']
        d6, d8 = 1.0 / (1.0 + 6.0 * t6), 1.0 / (1.0 + 6.0 * t8)
        g = dp['s6'] * d6 / r ** 6 + 3.0 * dp['s8'] * r42 * d8 / r ** 8
        dg = (dp['s6'] * (-6.0 * d6 / r ** 7 + 6.0 * dp['alp6'] * t6 * d6 ** 2 / r ** 7)
              + 3.0 * dp['s8'] * r42 * (-8.0 * d8 / r ** 9 + 6.0 * dp['alp8'] * t8 * d8 ** 2 / r ** 9))
    else:
        raise ValueError('damping must be damp_bj or damp_zero')
    energy = -0.5 * (C6 * g).sum()
    dEdr = -0.5 * C6 * dg                                    # of this ordered half pair
    fvec = dEdr[:, None] * d / r[:, None]                    # dE/d(x_j - x_i) direction
    forces = np.zeros((n, 3))
    np.add.at(forces, i, fvec)
    np.add.at(forces, j, -fvec)
    sigma = -(fvec[:, :, None] * d[:, None, :]).sum(0)       # reference: sigma += vec (x) rij with vec = -dE/dr r^ ... (sign below)
    dc6i = np.zeros(n)                                       # = -dE/dCN_i
    np.add.at(dc6i, i, 0.5 * g * dc6_i[i, j])
    np.add.at(dc6i, j, 0.5 * g * dc6_j[i, j])

    # CN chain rule (:1797-1962)
    i, j, d, r = pairs(cn_cutoff, True)
    rc = rcov[i] + rcov[j]
    ex = np.exp(-K1 * (rc / r - 1.0))
    dcn = -K1 * rc * ex / (r * r * (1.0 + ex) ** 2)          # d cnf / dr
    x1 = 0.5 * dcn * (dc6i[i] + dc6i[j])                     # ordered pairs: half each; dE/dr = -x1
    vec = x1[:, None] * d / r[:, None]
    np.add.at(forces, i, -vec)
    np.add.at(forces, j, vec)
    sigma += (vec[:, :, None] * d[:, None, :]).sum(0)
    return dict(energy=energy * AU_TO_EV, forces=forces * AU_TO_EV / AU_TO_ANG, sigma=sigma * AU_TO_EV, cn=cn,
                c6=c6, dc6i=dc6i)


def ase_results(numbers, positions, cell, pbc=(True, True, True), **kw):
    """What ``D3Calculator.calculate`` returns (``sevenn/calculator.py:528-614``): energy, forces, and the ASE
    Voigt stress ``-(xx, yy, zz, yz, xz, xy) / volume`` of the virial tensor."""
    out = d3_reference(numbers, positions, cell, pbc, **kw)
    s = out['sigma']
    vol = abs(np.linalg.det(np.asarray(cell, dtype=np.float64)))
    stress = -np.array([s[0, 0], s[1, 1], s[2, 2], s[1, 2], s[0, 2], s[0, 1]]) / vol
    return dict(energy=out['energy'], forces=out['forces'], stress=stress)
