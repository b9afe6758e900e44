"""Batched evaluation (disjoint-union graph) and the TorchSim-shaped adapter against the oracle and
the reference's golden vectors, one structure at a time."""
import types

import numpy as np
import pytest

from helpers import golden_vectors, model_weights, oracle, species_of, system_graph

pytestmark = pytest.mark.gpu

CASES = ['7net0_nacl_rattled', '7net0_hfo2_0', '7net0_h2o_rattled', '7net0_hfo2_1', '7net0_single_o']


def _systems():
    out = []
    for k in CASES:
        s = golden_vectors()[k]['system']
        out.append(dict(numbers=s['numbers'], positions=s['positions'], cell=s['cell'], pbc=bool(s['pbc']), key=k))
    return out


def test_batched_evaluator_matches_per_structure_oracle():
    from sevenn_b200.batch import BatchedEvaluator
    from sevenn_b200.engine import B200Engine
    meta, arrays = model_weights('sevennet_0')
    ev = BatchedEvaluator(B200Engine(meta, arrays))
    systems = _systems()
    res = ev.split(ev.compute(systems))
    ora = oracle('sevennet_0')
    n_edges = 0
    for s, r in zip(systems, res):
        g = golden_vectors()[s['key']]
        ei, evec, vol = system_graph(g['system'], 5.0)
        n_edges += ei.shape[1]
        ref = ora.forward(species_of(meta, s['numbers']), ei, evec, volume=vol)
        assert abs(r['energy'] - float(ref['energy'])) < 2e-5, s['key']
        assert np.allclose(r['energies'], ref['atomic_energy'].numpy(), atol=1e-5), s['key']
        assert np.allclose(r['forces'], ref['forces'].numpy(), atol=3e-5), s['key']
        assert np.allclose(r['virial'], ref['virial'].numpy(), atol=5e-4), s['key']
        assert abs(r['energy'] - g['energy']) < max(g['atol']['energy'], 5e-5), s['key']     # the reference's own numbers
        assert np.allclose(r['forces'], g['forces'], atol=max(g['atol']['forces'], 5e-5)), s['key']
    assert ev.engine.n_edges == n_edges
    # order independence: the same structures reversed give the same per-structure numbers
    res2 = ev.split(ev.compute(systems[::-1]))[::-1]
    for a, b in zip(res, res2):
        assert abs(a['energy'] - b['energy']) < 1e-6 and np.allclose(a['forces'], b['forces'], atol=2e-6)


def test_torchsim_shaped_model():
    import torch
    from sevenn_b200.batch import SevenNetModel
    keys = ['7net0_nacl_rattled', '7net0_hfo2_0', '7net0_hfo2_1']
    gs = [golden_vectors()[k] for k in keys]
    state = types.SimpleNamespace(
        positions=torch.tensor(np.concatenate([g['system']['positions'] for g in gs]), dtype=torch.float32),
        row_vector_cell=torch.tensor(np.stack([g['system']['cell'] for g in gs]), dtype=torch.float32),
        pbc=torch.tensor([True, True, True]),
        atomic_numbers=torch.tensor(np.concatenate([g['system']['numbers'] for g in gs])),
        system_idx=torch.tensor(np.concatenate([[i] * len(g['system']['numbers']) for i, g in enumerate(gs)])))
    model = SevenNetModel('7net-0', device='cuda')
    out = model(state)
    assert out['energy'].shape == (3,) and out['stress'].shape == (3, 3, 3) and out['forces'].shape[1] == 3
    e = out['energy'].cpu().numpy()
    f = out['forces'].cpu().numpy()
    st = out['stress'].cpu().numpy()
    a = 0
    for i, g in enumerate(gs):
        n = len(g['system']['numbers'])
        assert abs(e[i] - g['energy']) < max(g['atol']['energy'], 1e-4)     # float32 positions in, float32 energy out
        assert np.allclose(f[a:a + n], g['forces'], atol=2e-4)
        a += n
        assert np.allclose(st[i], st[i].T)
    v = golden_vectors()['7net0_nacl_rattled']['ase_stress']   # ASE Voigt (xx,yy,zz,yz,xz,xy)
    full = np.array([[v[0], v[5], v[4]], [v[5], v[1], v[3]], [v[4], v[3], v[2]]])
    assert np.allclose(st[0], full, atol=2e-5)
    with pytest.raises(NotImplementedError):
        SevenNetModel('7net-0', compute_atomic_virial=True)
    with pytest.raises(ValueError):
        SevenNetModel('7net-0', dtype=torch.float64)
