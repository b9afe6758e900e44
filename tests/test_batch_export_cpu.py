"""CPU checks of the host-side pieces added around the engine: the flat model file written for C/C++
hosts (sevenn_b200/export.py, read back here byte by byte the way examples/host_entry.cpp does) and
the batching logic of sevenn_b200/batch.py (union graph, per-structure reductions) with an engine
stand-in that evaluates the union graph with the oracle."""
import ctypes
import struct

import numpy as np
import torch

from helpers import golden_vectors, model_weights, oracle, species_of, system_graph


def test_flat_model_file_round_trip(tmp_path):
    from sevenn_b200.engine import S7bModelDesc, default_table_knots, prepare_params
    from sevenn_b200.export import export_flat
    from sevenn_b200.spec import build_spec
    meta, arrays = model_weights('sevennet_0')
    path = str(tmp_path / 'm.s7b')
    export_flat(path, meta, arrays)
    spec = build_spec(meta)
    want = prepare_params(spec, arrays, 'table', default_table_knots(spec))
    with open(path, 'rb') as f:
        assert f.read(8) == b'S7BMODEL'
        assert struct.unpack('<i', f.read(4))[0] == 1
        d = S7bModelDesc.from_buffer_copy(f.read(ctypes.sizeof(S7bModelDesc)))
        assert (d.n_layers, d.lmax_filter, d.num_species, d.n_basis) == (5, 2, spec.num_species, 8)
        assert abs(d.cutoff - 5.0) < 1e-6 and d.table_knots == default_table_knots(spec)
        n_arrays, n_types = struct.unpack('<ii', f.read(8))
        tm = dict(struct.unpack('<ii', f.read(8)) for _ in range(n_types))
        assert tm == spec.type_map and n_arrays == len(want)
        seen = set()
        for _ in range(n_arrays):
            name = f.read(32).rstrip(b'\0').decode()
            layer, numel = struct.unpack('<iq', f.read(12))
            data = np.frombuffer(f.read(4 * numel), dtype=np.float32)
            ref = np.ascontiguousarray(want[(name, layer)], dtype=np.float32).ravel()
            assert data.shape == ref.shape and np.array_equal(data.view(np.uint32), ref.view(np.uint32)), (name, layer)
            seen.add((name, layer))
        assert seen == set(want) and f.read(1) == b''


class OracleEngine:
    """Implements the slice of B200Engine that BatchedEvaluator uses, on the CPU oracle."""
    torch = torch
    device = torch.device('cpu')

    def __init__(self, name):
        from sevenn_b200.spec import build_spec
        self.meta, _ = model_weights(name)
        self.spec = build_spec(self.meta)
        self.ora = oracle(name)

    def set_positions(self, sp, pos, cell, pbc):
        from sevenn_b200.neighbors import build_graph
        cell = np.zeros((3, 3)) if cell is None else np.asarray(cell, float)
        ei, ev = build_graph(np.asarray(pos, float), cell, bool(np.all(pbc)), self.spec.cutoff)
        order = np.argsort(ei[0], kind='stable')
        self._ga = (np.concatenate([[0], np.cumsum(np.bincount(ei[0], minlength=len(sp)))]), ei[1][order], ev[order])

    def graph_arrays(self):
        rp, s, ev = self._ga
        return torch.tensor(rp, dtype=torch.int32), torch.tensor(s, dtype=torch.int32), torch.tensor(ev, dtype=torch.float32)

    def set_graph_csr(self, species, rowptr, src, ev, n_local):
        self._graph = dict(species=species, rowptr=rowptr, src=src, edge_vec=ev)
        self.n_nodes = self.n_local = n_local
        self.n_edges = len(src)

    def compute(self):
        g = self._graph
        dst = torch.repeat_interleave(torch.arange(self.n_local), (g['rowptr'][1:] - g['rowptr'][:-1]).long())
        self.out = self.ora.forward(g['species'].numpy().astype(np.int64), np.stack([dst.numpy(), g['src'].numpy()]),
                                    g['edge_vec'].numpy().astype(np.float64), volume=0.0)

    def buffer(self, name, shape=None, **kw):
        return self.out[name].float().reshape(shape)


def test_batched_evaluator_union_graph_and_reductions():
    from sevenn_b200.batch import BatchedEvaluator
    keys = ['7net0_nacl_rattled', '7net0_hfo2_0', '7net0_h2o_rattled', '7net0_single_o']
    eng = OracleEngine('sevennet_0')
    ev = BatchedEvaluator(eng)
    systems = []
    for k in keys:
        s = golden_vectors()[k]['system']
        systems.append(dict(numbers=s['numbers'], positions=s['positions'], cell=s['cell'], pbc=bool(s['pbc'])))
    out = ev.compute(systems)
    res = ev.split(out)
    assert out['energy'].shape == (4,) and out['virial'].shape == (4, 6)
    assert eng._graph['rowptr'][-1] == eng.n_edges and len(eng._graph['rowptr']) == eng.n_nodes + 1
    a = 0
    for k, s, r in zip(keys, systems, res):
        g = golden_vectors()[k]
        n = len(s['numbers'])
        ei, evec, vol = system_graph(g['system'], 5.0)
        ref = eng.ora.forward(species_of(eng.meta, s['numbers']), ei, evec, volume=vol)
        assert abs(r['energy'] - g['energy']) < max(g['atol']['energy'], 3e-5), k
        assert np.allclose(r['forces'], g['forces'], atol=max(g['atol']['forces'], 2e-5)), k
        assert np.allclose(r['virial'], ref['virial'].numpy(), atol=2e-5), k
        src = eng._graph['src'][eng._graph['rowptr'][a]:eng._graph['rowptr'][a + n]]
        assert src.numel() == 0 or (int(src.min()) >= a and int(src.max()) < a + n)     # no cross-structure edges
        a += n
    # unknown element -> ValueError, as the reference's type-map lookup
    try:
        ev.set_batch([dict(numbers=[118], positions=[[0, 0, 0]], cell=np.eye(3) * 10, pbc=True)])
    except ValueError as e:
        assert 'not known' in str(e)
    else:
        raise AssertionError('expected ValueError')
