"""Host-side neighbour-list construction (the step *before* the hot path).

Restates what the reference's graph builder produces
(``sevenn/train/dataload.py:32-129``: matscipy/ASE ``neighbour_list('ijDS')``):
directed edges i <- j for every pair (including periodic images and self-images at
non-zero shift) with |r_j - r_i + S.cell| < cutoff, ``edge_index[0] = i`` (centre),
``edge_index[1] = j`` (neighbour), ``edge_vec = r_j - r_i + S.cell``, sorted by centre i.

Two implementations with identical output sets: a brute-force one over all image shifts
(any cell, any size of cutoff relative to the cell; O(N^2 n_images)) for small systems and a
binned cell list (vectorised numpy) for the large synthetic benchmark cells.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def _n_images(cell: np.ndarray, pbc, cutoff: float) -> np.ndarray:
    vol = abs(np.linalg.det(cell))
    n = np.zeros(3, dtype=np.int64)
    for a in range(3):
        if not pbc[a]:
            continue
        b, c = cell[(a + 1) % 3], cell[(a + 2) % 3]
        height = vol / np.linalg.norm(np.cross(b, c))
        n[a] = int(np.ceil(cutoff / height))
    return n


def neighbor_list_brute(pos, cell, pbc, cutoff: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    pos = np.asarray(pos, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    n = len(pos)
    if abs(np.linalg.det(cell)) < 1e-12:
        nimg = np.zeros(3, dtype=np.int64)
        cell = np.eye(3)
    else:
        nimg = _n_images(cell, pbc, cutoff)
    shifts = np.array([(a, b, c)
                       for a in range(-nimg[0], nimg[0] + 1)
                       for b in range(-nimg[1], nimg[1] + 1)
                       for c in range(-nimg[2], nimg[2] + 1)], dtype=np.float64)
    ii, jj, vv, ss = [], [], [], []
    for S in shifts:
        d = pos[None, :, :] - pos[:, None, :] + (S @ cell)[None, None, :]   # [i, j, 3]
        r2 = (d * d).sum(-1)
        mask = r2 < cutoff * cutoff
        if not S.any():
            mask &= ~np.eye(n, dtype=bool)
        i, j = np.nonzero(mask)
        ii.append(i), jj.append(j), vv.append(d[i, j]), ss.append(np.broadcast_to(S, (len(i), 3)))
    i = np.concatenate(ii)
    j = np.concatenate(jj)
    v = np.concatenate(vv) if len(i) else np.zeros((0, 3))
    s = np.concatenate(ss) if len(i) else np.zeros((0, 3))
    order = np.lexsort((j, i))
    return np.stack([i[order], j[order]]).astype(np.int64), v[order], s[order]


def neighbor_list_cells(pos, cell, cutoff: float) -> Tuple[np.ndarray, np.ndarray]:
    """Fully periodic, orthorhombic-or-triclinic cell with every cell height >= 3*cutoff is
    not required, only >= 2*cutoff (minimum image inside the 27-bin stencil).  Returns
    (edge_index [2,E] sorted by centre, edge_vec [E,3])."""
    pos = np.asarray(pos, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64).reshape(3, 3)
    inv = np.linalg.inv(cell)
    frac = pos @ inv
    frac -= np.floor(frac)
    vol = abs(np.linalg.det(cell))
    heights = np.array([vol / np.linalg.norm(np.cross(cell[(a + 1) % 3], cell[(a + 2) % 3]))
                        for a in range(3)])
    nb = np.maximum(np.floor(heights / cutoff).astype(np.int64), 1)
    if (nb < 3).any():
        ei, ev, _ = neighbor_list_brute(pos, cell, True, cutoff)
        return ei, ev
    b = np.minimum((frac * nb).astype(np.int64), nb - 1)
    lin = (b[:, 0] * nb[1] + b[:, 1]) * nb[2] + b[:, 2]
    order = np.argsort(lin, kind='stable')
    lin_sorted = lin[order]
    nbins = int(nb.prod())
    start = np.searchsorted(lin_sorted, np.arange(nbins))
    count = np.searchsorted(lin_sorted, np.arange(nbins), side='right') - start
    wrapped = frac @ cell
    out_i, out_j, out_v = [], [], []
    # loop over the 27 stencil offsets; vectorise over (bin, atom-in-bin) pairs
    for da in (-1, 0, 1):
        for db in (-1, 0, 1):
            for dc in (-1, 0, 1):
                nbv = b + np.array([da, db, dc])
                sh = np.floor_divide(nbv, nb)              # image shift of the neighbour bin
                nbw = nbv - sh * nb
                nlin = (nbw[:, 0] * nb[1] + nbw[:, 1]) * nb[2] + nbw[:, 2]
                cnt = count[nlin]                           # neighbours available per atom i
                tot = int(cnt.sum())
                if tot == 0:
                    continue
                i_rep = np.repeat(np.arange(len(pos)), cnt)
                offs = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                j_idx = order[np.repeat(start[nlin], cnt) + offs]
                d = wrapped[j_idx] - wrapped[i_rep] + (sh[i_rep].astype(np.float64) @ cell)
                r2 = (d * d).sum(-1)
                m = (r2 < cutoff * cutoff) & ~((i_rep == j_idx) & (r2 < 1e-20))
                out_i.append(i_rep[m]), out_j.append(j_idx[m]), out_v.append(d[m])
    i = np.concatenate(out_i)
    j = np.concatenate(out_j)
    v = np.concatenate(out_v)
    o = np.lexsort((j, i))
    return np.stack([i[o], j[o]]).astype(np.int64), v[o]


def build_graph(pos, cell, pbc, cutoff: float):
    """(edge_index, edge_vec) for any system; picks the binned builder for large periodic cells."""
    pbc3 = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    if len(pos) > 400 and pbc3.all():
        return neighbor_list_cells(pos, cell, cutoff)
    ei, ev, _ = neighbor_list_brute(pos, cell, pbc3, cutoff)
    return ei, ev


# ---- synthetic benchmark cells ------------------------
def diamond_si(nx: int, ny: int, nz: int, a: float = 5.431, sigma: float = 0.05, seed: int = 0):
    basis = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                      [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij'), -1).reshape(-1, 3)
    pos = ((g[:, None, :] + basis[None, :, :]) * a).reshape(-1, 3)
    pos = pos + np.random.RandomState(seed).normal(scale=sigma, size=pos.shape)
    cell = np.diag([nx * a, ny * a, nz * a])
    return pos, cell, np.full(len(pos), 14, dtype=np.int64)


def rocksalt_nacl(nx: int, ny: int, nz: int, a: float = 5.64, sigma: float = 0.05, seed: int = 0):
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    basis = np.concatenate([fcc, fcc + np.array([.5, 0, 0])])
    z = np.array([11] * 4 + [17] * 4)
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing='ij'), -1).reshape(-1, 3)
    pos = ((g[:, None, :] + basis[None, :, :]) * a).reshape(-1, 3)
    pos = pos + np.random.RandomState(seed).normal(scale=sigma, size=pos.shape)
    cell = np.diag([nx * a, ny * a, nz * a])
    return pos, cell, np.tile(z, len(g))
