"""Protocol test of the ML-IAP front-end (sevenn_b200/mliap.py) on the CPU: a periodic cell is presented
the way LAMMPS does -- owned atoms plus ghost copies for the periodic images, pairs pointing at ghosts,
forward/reverse exchange callbacks -- and must give the energy and forces of the plain periodic graph.
The engine is the CPU stand-in of tests/test_parallel_gloo.py (same stage interface as B200Engine)."""
import numpy as np
import torch

from sevenn_b200.engine import (STAGE_BWD_END, STAGE_BWD_LAYER_A, STAGE_BWD_LAYER_B, STAGE_FWD_BEGIN, STAGE_FWD_END,
                                STAGE_FWD_LAYER)
from sevenn_b200.neighbors import build_graph, rocksalt_nacl
from test_parallel_gloo import FakeEngine


class LammpsData:
    """What LAMMPS' MLIAPData exposes to a unified model (the subset the wrapper uses)."""

    def __init__(self, pos, cell, elems, cutoff):
        n = len(pos)
        ei, ev = build_graph(pos, cell, True, cutoff)
        shift = np.rint((ev - (pos[ei[1]] - pos[ei[0]])) @ np.linalg.inv(cell)).astype(int)
        ghosts, pair_j = {}, []
        for e in range(ei.shape[1]):
            s = tuple(shift[e])
            if s == (0, 0, 0):
                pair_j.append(int(ei[1, e]))
            else:
                key = (int(ei[1, e]),) + s
                pair_j.append(n + ghosts.setdefault(key, len(ghosts)))
        self.owner = torch.tensor([k[0] for k in ghosts], dtype=torch.long)
        self.nlocal, self.ntotal, self.npairs = n, n + len(ghosts), ei.shape[1]
        order = np.random.RandomState(0).permutation(self.npairs)            # LAMMPS promises no pair order
        self.pair_i, self.pair_j, self.rij = ei[0][order], np.array(pair_j)[order], ev[order]
        self.elems = np.concatenate([elems, elems[self.owner.numpy()]])
        self.eatoms = np.zeros(n)
        self.energy = None
        self.f = None
        self.calls = []

    def forward_exchange(self, src, dst, vec_len):
        assert src.shape == dst.shape == (self.ntotal, vec_len)
        self.calls.append('fwd')
        dst[:self.nlocal] = src[:self.nlocal]
        dst[self.nlocal:] = src[self.owner]

    def reverse_exchange(self, src, dst, vec_len):
        assert src.shape == dst.shape == (self.ntotal, vec_len)
        self.calls.append('rev')
        dst.zero_()
        dst[:self.nlocal] = src[:self.nlocal]
        dst[:self.nlocal].index_add_(0, self.owner, src[self.nlocal:])

    def update_pair_forces_gpu(self, fij):
        assert fij.dtype == torch.float64 and fij.shape == (self.npairs, 3)
        f = torch.zeros(self.ntotal, 3, dtype=torch.float64)
        f.index_add_(0, torch.as_tensor(self.pair_i), fij)
        f.index_add_(0, torch.as_tensor(self.pair_j), -fij)
        f[:self.nlocal].index_add_(0, self.owner, f[self.nlocal:])             # LAMMPS reverse-communicates ghost forces
        self.f = f[:self.nlocal]


class SortingEngine(FakeEngine):
    """FakeEngine + the centre-sorting that B200Engine.set_graph does (returns the permutation)."""

    def set_graph(self, species, edge_index, edge_vec, n_local=None):
        ei = torch.as_tensor(edge_index).long()
        perm = torch.argsort(ei[0], stable=True)
        super().set_graph(species, ei[:, perm], torch.as_tensor(edge_vec)[perm], n_local)
        return dict(perm=perm)


def test_mliap_wrapper_protocol_matches_periodic_graph():
    from sevenn_b200.mliap import SevenNetMLIAPWrapper
    pos, cell, z = rocksalt_nacl(2, 2, 2, sigma=0.08, seed=7)
    elems = (z == z.min()).astype(np.int64)
    data = LammpsData(pos, cell, elems, 5.0)
    assert data.ntotal > data.nlocal
    wrapper = SevenNetMLIAPWrapper('stand-in', engine=SortingEngine())
    wrapper.compute_forces(data)

    ser = FakeEngine()
    ei, ev = build_graph(pos, cell, True, 5.0)
    ser.set_graph(elems, ei, ev)
    ser.run_stage(STAGE_FWD_BEGIN)
    for t in range(ser.T):
        ser.run_stage(STAGE_FWD_LAYER, t)
    ser.run_stage(STAGE_FWD_END)
    for t in range(ser.T - 1, -1, -1):
        ser.run_stage(STAGE_BWD_LAYER_A, t)
        if t > 0:
            ser.run_stage(STAGE_BWD_LAYER_B, t)
    ser.run_stage(STAGE_BWD_END)

    assert abs(float(data.energy) - float(ser.energy[0])) < 1e-9 * abs(float(ser.energy[0]))
    assert np.allclose(data.eatoms.sum(), float(ser.energy[0]))
    assert np.allclose(data.f.numpy(), ser.forces.numpy(), atol=1e-10)
    assert data.calls == ['fwd'] * (ser.T - 1) + ['rev'] * (ser.T - 1)       # no exchange for layer 0
    assert wrapper.element_types[0] == 'X' and wrapper.element_types[14] == 'Si' and wrapper.rcutfac == 0.0


def test_mliap_wrapper_handles_empty_and_refuses_modal():
    import pytest
    from sevenn_b200.mliap import SevenNetMLIAPWrapper
    wrapper = SevenNetMLIAPWrapper('stand-in', engine=FakeEngine())

    class Empty:
        nlocal = 0
        ntotal = 0
        npairs = 0
    wrapper.compute_forces(Empty())          # returns without touching the engine (mliap.py:180-181)
    with pytest.raises(NotImplementedError):
        SevenNetMLIAPWrapper('7net-0', modal='mpa')
