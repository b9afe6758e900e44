"""Several structures in one pass of the engine (the batched callers of the hot path).

The reference evaluates a batch as one disjoint-union graph: ``AtomGraphSequential`` with
``is_batch_data`` (``sevenn/nn/sequential.py:99-108,143``), per-graph energy from ``AtomReduce``
(``nn/linear.py:127-141``) and per-graph stress from the per-atom virial scattered by ``batch``
(``nn/force_output.py:216-228``).  The engine works on any CSR graph, so a batch is the concatenation
of the per-structure graphs with node offsets; the per-structure reductions run as torch segment sums
on the device (plumbing).  ``SevenNetModel`` mirrors the TorchSim adapter ``sevenn/torchsim.py:56-292``:
same constructor keywords, ``forward(state) -> {'energy' [B], 'forces' [n,3], 'stress' [B,3,3]}``
with the sign / Voigt handling of ``torchsim.py:286-290``.  ``torch_sim`` is not installed here, so
``state`` is duck-typed: ``positions, row_vector_cell (or cell), pbc, atomic_numbers, system_idx``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .engine import B200Engine


class BatchedEvaluator:
    def __init__(self, engine: B200Engine):
        self.engine = engine

    def set_batch(self, systems: Sequence[dict]):
        """systems: dicts with ``species`` (indices) or ``numbers`` (atomic numbers), ``positions``
        [n,3], ``cell`` [3,3] and ``pbc``.  Builds every neighbour list on the device
        (``s7b_engine_set_positions_host``) and installs the union graph."""
        eng, torch = self.engine, self.engine.torch
        tm = eng.spec.type_map
        rowptrs, srcs, evs, species, counts, offs = [], [], [], [], [], [0]
        e_off = 0
        for s in systems:
            if 'species' in s:
                sp = np.asarray(s['species'], dtype=np.int32)
            else:
                try:
                    sp = np.array([tm[int(z)] for z in s['numbers']], dtype=np.int32)
                except KeyError as e:
                    raise ValueError(f'atomic number {e} is not known to this model') from None
            eng.set_positions(sp, s['positions'], s.get('cell'), s.get('pbc', False))
            rp, src, ev = eng.graph_arrays()
            rowptrs.append(rp[(1 if rowptrs else 0):].to(torch.int64) + e_off)
            srcs.append(src.to(torch.int64) + offs[-1])
            evs.append(ev.clone())
            species.append(torch.as_tensor(sp, device=eng.device))
            counts.append(len(sp))
            offs.append(offs[-1] + len(sp))
            e_off += int(src.shape[0])
        if not counts:
            raise ValueError('empty batch')
        cat = torch.cat
        self.counts = counts
        self.system_idx = torch.repeat_interleave(torch.arange(len(counts), device=eng.device),
                                                  torch.tensor(counts, device=eng.device))
        eng.set_graph_csr(cat(species).to(torch.int32).contiguous(), cat(rowptrs).to(torch.int32).contiguous(),
                          cat(srcs).to(torch.int32).contiguous(), cat(evs).contiguous(), offs[-1])
        return self

    def compute(self, systems: Optional[Sequence[dict]] = None) -> dict:
        """-> dict of device tensors: energy [B] f64, atomic_energy [n], forces [n,3], virial [B,6] f64
        (= -sum r (x) f per structure, order xx,yy,zz,xy,yz,zx), n_edges."""
        if systems is not None:
            self.set_batch(systems)
        eng, torch = self.engine, self.engine.torch
        eng.compute()
        B = len(self.counts)
        ae = eng.buffer('atomic_energy', shape=(eng.n_local,)).clone()
        energy = torch.zeros(B, dtype=torch.float64, device=eng.device).index_add_(0, self.system_idx, ae.double())
        forces = eng.buffer('forces', shape=(eng.n_nodes, 3)).clone()
        virial = torch.zeros(B, 6, dtype=torch.float64, device=eng.device)
        if eng.n_edges:
            g = eng._graph
            fe = eng.buffer('edge_force', shape=(eng.n_edges, 3)).double()
            ev = g['edge_vec'].double()
            v6 = ev[:, [0, 1, 2, 0, 1, 2]] * fe[:, [0, 1, 2, 1, 2, 0]]
            virial.index_add_(0, self.system_idx[g['src'].long()], -v6)
        return dict(energy=energy, atomic_energy=ae, forces=forces, virial=virial, n_edges=eng.n_edges)

    def split(self, out: dict) -> List[dict]:
        """Per-structure numpy results."""
        res, a = [], 0
        ae, f = out['atomic_energy'].cpu().numpy(), out['forces'].cpu().numpy()
        e, v = out['energy'].cpu().numpy(), out['virial'].cpu().numpy()
        for b, n in enumerate(self.counts):
            res.append(dict(energy=float(e[b]), energies=ae[a:a + n], forces=f[a:a + n], virial=v[b]))
            a += n
        return res


class SevenNetModel:
    """TorchSim-style model wrapper (``sevenn/torchsim.py:56``): ``model(state)`` evaluates all systems of
    the state in one engine pass."""

    def __init__(self, model='7net-0', *, modal=None, neighbor_list_fn=None, enable_cueq=False,
                 enable_flash=False, enable_oeq=False, compute_atomic_virial=False, device='auto',
                 dtype=None, radial: str = 'table'):
        import torch
        if compute_atomic_virial:   # torchsim.py:112-116
            raise NotImplementedError('compute_atomic_virial is not supported for SevenNet TorchSim interface.')
        if modal is not None:
            raise NotImplementedError('multi-fidelity models are out of scope')
        if enable_cueq or enable_flash or enable_oeq:
            raise ValueError('enable_cueq/flash/oeq select other accelerators; this model always runs the sevenn_b200 engine')
        if neighbor_list_fn is not None:
            raise ValueError('the neighbour list is built by the engine on the device; neighbor_list_fn is not used')
        if dtype is not None and dtype is not torch.float32:   # torchsim.py:135-139
            raise ValueError(f'SevenNet currently only supports {torch.float32}, but received different dtype: {dtype}')
        dev = torch.device('cuda' if device == 'auto' else device)
        if dev.type != 'cuda':
            raise RuntimeError('sevenn_b200 has no CPU path; pass a CUDA device')
        from .calculator import resolve_model
        meta, arrays = resolve_model(model) if isinstance(model, str) else model
        self.engine = B200Engine(meta, arrays, radial=radial, device=dev.index)
        self._device, self._dtype = self.engine.device, torch.float32
        self.cutoff = torch.tensor(self.engine.spec.cutoff)
        self.type_map = self.engine.spec.type_map
        self.modal = None
        self.implemented_properties = ['energy', 'forces', 'stress']
        self._batch = BatchedEvaluator(self.engine)

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return self._dtype

    def forward(self, state, **kwargs):
        torch = self.engine.torch
        pos = torch.as_tensor(state.positions).detach().cpu().double().numpy()
        cells = getattr(state, 'row_vector_cell', None)
        if cells is None:   # SimState.cell holds column vectors
            cells = torch.as_tensor(state.cell).transpose(-1, -2)
        cells = torch.as_tensor(cells).detach().cpu().double().numpy().reshape(-1, 3, 3)
        numbers = torch.as_tensor(state.atomic_numbers).cpu().numpy()
        sys_idx = torch.as_tensor(state.system_idx).cpu().numpy()
        pbc = np.broadcast_to(np.asarray(torch.as_tensor(state.pbc).cpu().numpy(), dtype=bool), (3,))
        B = int(sys_idx.max()) + 1
        if np.any(np.diff(sys_idx) < 0):
            raise ValueError('system_idx must be sorted')
        systems = [dict(numbers=numbers[sys_idx == b], positions=pos[sys_idx == b], cell=cells[b], pbc=pbc)
                   for b in range(B)]
        out = self._batch.compute(systems)
        vol = torch.as_tensor(np.abs(np.linalg.det(cells)), device=self._device)
        s = (out['virial'] / vol[:, None])                      # 'inferred_stress', (xx,yy,zz,xy,yz,zx)
        v = -s[:, [0, 1, 2, 4, 5, 3]]                           # ASE Voigt, sign of torchsim.py:286-290
        stress = torch.stack([torch.stack([v[:, 0], v[:, 5], v[:, 4]], -1),
                              torch.stack([v[:, 5], v[:, 1], v[:, 3]], -1),
                              torch.stack([v[:, 4], v[:, 3], v[:, 2]], -1)], -2)
        return {'energy': out['energy'].to(self._dtype), 'forces': out['forces'], 'stress': stress.to(self._dtype)}

    __call__ = forward
