"""Symmetry properties of the energy/force path: invariance of the energy under rotation, translation
and relabelling of the atoms, covariance of forces and virial.  Checked for the CPU oracle in fp64
(tight) and for the CUDA engine through its positions entry point (fp32 tolerances)."""
import numpy as np
import pytest

from helpers import golden_vectors, model_weights, oracle, species_of


def _rotation(seed):
    q, r = np.linalg.qr(np.random.RandomState(seed).normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _full(v6):   # (xx, yy, zz, xy, yz, zx) -> 3x3
    return np.array([[v6[0], v6[3], v6[5]], [v6[3], v6[1], v6[4]], [v6[5], v6[4], v6[2]]])


def _variants(system, seed=3):
    """(name, positions, cell, permutation, rotation) of transformed copies of a periodic system."""
    pos = np.asarray(system['positions'], dtype=np.float64)
    cell = np.asarray(system['cell'], dtype=np.float64)
    n = len(pos)
    R = _rotation(seed)
    rng = np.random.RandomState(seed)
    shift = rng.uniform(-3, 3, size=3)
    perm = rng.permutation(n)
    eye = np.eye(3)
    return [('rotated', pos @ R.T, cell @ R.T, np.arange(n), R),
            ('translated', pos + shift, cell, np.arange(n), eye),
            ('permuted', pos[perm], cell, perm, eye)]


@pytest.mark.parametrize('model', ['sevennet_0', 'sevennet_l3i5'])
def test_oracle_symmetries(model):
    from sevenn_b200.neighbors import build_graph
    meta, _ = model_weights(model)
    ora = oracle(model)
    sysd = golden_vectors()['7net0_hfo2_0']['system']
    z = np.asarray(sysd['numbers'])

    def run(pos, cell, zz):
        ei, ev = build_graph(pos, cell, True, 5.0)
        o = ora.forward(species_of(meta, zz), ei, ev, volume=abs(np.linalg.det(cell)))
        return float(o['energy']), o['forces'].numpy(), o['virial'].numpy()

    e0, f0, v0 = run(np.asarray(sysd['positions'], float), np.asarray(sysd['cell'], float), z)
    for name, pos, cell, perm, R in _variants(sysd):
        e, f, v = run(pos, cell, z[perm])
        assert abs(e - e0) < 1e-8, name
        assert np.allclose(f, f0[perm] @ R.T, atol=1e-8), name
        assert np.allclose(_full(v), R @ _full(v0) @ R.T, atol=1e-7), name


@pytest.mark.gpu
@pytest.mark.parametrize('model', ['sevennet_0', 'sevennet_l3i5'])
def test_engine_symmetries(model):
    from sevenn_b200.engine import B200Engine
    meta, arrays = model_weights(model)
    eng = B200Engine(meta, arrays)
    sysd = golden_vectors()['7net0_hfo2_0']['system']
    z = np.asarray(sysd['numbers'])
    sp = species_of(meta, z).astype(np.int32)
    e0, _, f0, v0, n0 = eng.compute_positions(sp, sysd['positions'], sysd['cell'], True)
    for name, pos, cell, perm, R in _variants(sysd):
        e, _, f, v, n = eng.compute_positions(sp[perm], pos, cell, True)
        assert n == n0, name
        assert abs(e - e0) < 5e-5, (name, e - e0)
        assert np.allclose(f, f0[perm] @ R.T, atol=5e-5), name
        assert np.allclose(_full(v), R @ _full(v0) @ R.T, atol=5e-4), name
