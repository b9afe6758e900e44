"""D3 dispersion front-end: ``D3Calculator`` / ``SevenNetD3Calculator`` with the constructor and result
keys of the reference (``sevenn/calculator.py:236-314, 387-618``) on top of the cell-list CUDA kernels of
``csrc/d3_kernels.cuh`` (C ABI ``s7b_d3_*``), plus the multi-GPU driver the reference does not have
(its D3 is single-GPU and limited to 46 340 atoms, ``docs/source/user_guide/d3.md:7,53``).

Unlike the reference binding there is no LAMMPS-frame rotation: the library takes lattice vectors in any
orientation, so forces and the virial come back in the caller's frame.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Optional

import numpy as np

from .engine import check, load_library

_PARAMS = None
AU_TO_ANG = 0.52917726


def d3_tables():
    """Grimme's D3 reference tables, converted from the reference's ``pair_d3_pars.h`` by
    ``tools/convert_d3_params.py`` (``weights/d3_params.npz``)."""
    global _PARAMS
    if _PARAMS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'weights', 'd3_params.npz')
        f = np.load(path)
        p = {k: f[k] for k in ('r0ab', 'c6ref', 'cnref', 'mxc', 'r2r4', 'rcov')}
        p['functionals'] = json.loads(bytes(f['functionals']).decode())
        _PARAMS = p
    return _PARAMS


class D3Engine:
    """One D3 evaluator on one GPU (thin ctypes host of ``s7b_d3_*``)."""

    def __init__(self, damping_type: str = 'damp_bj', functional_name: str = 'pbe', vdw_cutoff: float = 9000.0,
                 cn_cutoff: float = 1600.0, device: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise NotImplementedError('CPU + D3 is not implemented')       # same message class as calculator.py:421
        self.torch = torch
        self.lib = load_library()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        damping_type, functional_name = damping_type.lower(), functional_name.lower()
        if damping_type not in ('damp_bj', 'damp_zero'):
            raise ValueError('Error: Invalid damping type.')
        T = d3_tables()
        if functional_name not in T['functionals'][damping_type]:
            raise ValueError(f'Functional name unknown: {functional_name}')
        p = T['functionals'][damping_type][functional_name]
        self.damping = 1 if damping_type == 'damp_bj' else 0
        self.par = dict(s6=p['s6'], s8=p['s18'], a1=p['rs6'], a2=p['rs18'], alp6=p['alp'], alp8=p['alp'] + 2.0)
        self.rthr, self.cnthr = float(vdw_cutoff), float(cn_cutoff)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.s7b_d3_create(ctypes.byref(self._h)))
        self._numbers = None
        self.n = 0

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None and self._h.value:
                self.lib.s7b_d3_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def set_system(self, numbers, positions, cell, pbc=(True, True, True)):
        """numbers [n] atomic numbers, positions [n,3] and cell rows in Angstrom."""
        numbers = np.asarray(numbers, dtype=np.int64)
        uniq = list(dict.fromkeys(numbers.tolist()))            # order of first appearance, as calculator.py:484-492
        if self._numbers != uniq:
            T = d3_tables()
            z = np.array(uniq) - 1
            f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
            rcov, r2r4 = f8(T['rcov'][z]), f8(T['r2r4'][z])
            r0, c6 = f8(T['r0ab'][np.ix_(z, z)]), f8(T['c6ref'][np.ix_(z, z)])
            cr, mxc = f8(T['cnref'][z]), np.ascontiguousarray(T['mxc'][z], dtype=np.int32)
            with self.torch.cuda.device(self.device):
                check(self.lib.s7b_d3_set_params(self._h, len(uniq), rcov.ctypes.data, r2r4.ctypes.data, r0.ctypes.data,
                                                 c6.ctypes.data, cr.ctypes.data, mxc.ctypes.data))
                check(self.lib.s7b_d3_set_damping(self._h, self.damping, self.par['s6'], self.par['s8'], self.par['a1'],
                                                  self.par['a2'], self.par['alp6'], self.par['alp8'], self.rthr, self.cnthr))
            self._numbers = uniq
        lut = {zz: i for i, zz in enumerate(uniq)}
        types = np.ascontiguousarray([lut[int(a)] for a in numbers], dtype=np.int32)
        pos = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        c = np.ascontiguousarray(cell, dtype=np.float64).reshape(3, 3)
        pb = np.ascontiguousarray(np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).astype(np.int32))
        self.n = len(types)
        with self.torch.cuda.device(self.device):
            check(self.lib.s7b_d3_set_system(self._h, self.n, types.ctypes.data, pos.ctypes.data, c.ctypes.data, pb.ctypes.data, self._stream()))
        return self

    def run_stage(self, stage: int, i_begin: int = 0, i_end: Optional[int] = None):
        with self.torch.cuda.device(self.device):
            check(self.lib.s7b_d3_run_stage(self._h, stage, i_begin, self.n if i_end is None else i_end, self._stream()))

    def buffer(self, name: str, dtype='f8', shape=None):
        from .engine import _DevView
        n = ctypes.c_size_t()
        ptr = self.lib.s7b_d3_buffer(self._h, name.encode(), ctypes.byref(n))
        shp = (n.value,) if shape is None else tuple(shape)
        return self.torch.as_tensor(_DevView(ptr, shp, '<' + dtype), device=self.device)

    def results(self):
        """(energy eV, forces [n,3] eV/A in the caller's atom order, sigma6 eV = sum f (x) r: xx,yy,zz,xy,xz,yz)"""
        e, s = np.zeros(1), np.zeros(6)
        f = np.zeros((self.n, 3))
        with self.torch.cuda.device(self.device):
            check(self.lib.s7b_d3_results_host(self._h, e.ctypes.data, f.ctypes.data, s.ctypes.data, self._stream()))
        return float(e[0]), f, s

    def compute(self, numbers, positions, cell, pbc=(True, True, True)):
        self.set_system(numbers, positions, cell, pbc)
        for stage in (1, 2, 3):
            self.run_stage(stage)
        return self.results()


def distributed_d3(engine: D3Engine, numbers, positions, cell, pbc=(True, True, True), group=None):
    """The same system on every rank (positions replicated: the 50 A interaction range is of the order of the
    box), each rank evaluating a contiguous slice of the bin-sorted atoms; ``cn`` and ``dc6i`` are
    all-gathered between the stages (NCCL), energy / virial all-reduced, forces all-gathered.
    Returns (energy, forces [n,3], sigma6) on every rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    engine.set_system(numbers, positions, cell, pbc)
    n = engine.n
    chunk = (n + world - 1) // world
    lo, hi = min(rank * chunk, n), min((rank + 1) * chunk, n)

    def gather(name, width):
        buf = engine.buffer(name, shape=(n, width) if width > 1 else (n,))
        padded = torch.zeros((world * chunk,) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
        mine = torch.zeros((chunk,) + tuple(buf.shape[1:]), dtype=buf.dtype, device=buf.device)
        mine[:hi - lo] = buf[lo:hi]
        dist.all_gather_into_tensor(padded, mine, group=group)
        buf.copy_(padded[:n])

    engine.run_stage(1, lo, hi)
    gather('cn', 1)
    engine.run_stage(2, lo, hi)
    gather('dc6i', 1)
    engine.run_stage(3, lo, hi)
    gather('force', 3)
    for name in ('energy', 'sigma'):
        dist.all_reduce(engine.buffer(name), group=group)
    return engine.results()


try:
    from ase.calculators.calculator import Calculator as _Base, all_changes as _all_changes
except Exception:   # ASE absent (as in the build container): duck-typed base, as sevenn_b200.calculator does
    _all_changes = ['positions', 'numbers', 'cell', 'pbc']

    class _Base:
        implemented_properties: list = []

        def __init__(self, **kwargs):
            self.results, self.atoms = {}, None

        def calculate(self, atoms=None, properties=None, system_changes=_all_changes):
            self.atoms = atoms


class D3Calculator(_Base):
    """ASE-style calculator of the D3 correction; constructor and ``results`` keys of
    ``sevenn/calculator.py:387-618`` (``free_energy, energy, forces, stress``)."""
    implemented_properties = ['free_energy', 'energy', 'forces', 'stress']

    def __init__(self, damping_type: str = 'damp_bj', functional_name: str = 'pbe', vdw_cutoff: float = 9000,
                 cn_cutoff: float = 1600, **kwargs):
        device = kwargs.pop('device', None)
        super().__init__(**kwargs)
        self.rthr, self.cnthr = vdw_cutoff, cn_cutoff
        self.engine = D3Engine(damping_type, functional_name, vdw_cutoff, cn_cutoff, device=device)

    def calculate(self, atoms=None, properties=None, system_changes=_all_changes):
        super().calculate(atoms, properties, system_changes)
        if atoms is None:
            raise ValueError('No atoms to evaluate')
        cell = np.asarray(atoms.get_cell(), dtype=np.float64).reshape(3, 3)
        pbc = np.asarray(atoms.get_pbc(), dtype=bool)
        pos = np.asarray(atoms.get_positions(), dtype=np.float64)
        if cell.sum() == 0:       # calculator.py:534-547: an orthogonal cell large enough, periodic "for minus positions"
            print('Warning: D3Calculator requires a cell.\nWarning: An orthogonal cell large enough is generated.')
            max_cutoff = np.sqrt(max(self.rthr, self.cnthr)) * AU_TO_ANG
            cell = np.eye(3) * (pos.max(axis=0) - pos.min(axis=0) + max_cutoff + 1.0)
            pbc = np.array([True, True, True])
            atoms.set_cell(cell)
            atoms.set_pbc(pbc)
        energy, forces, s = self.engine.compute(np.asarray(atoms.get_atomic_numbers()), pos, cell, pbc)
        vol = abs(np.linalg.det(cell))
        stress = -np.array([s[0], s[1], s[2], s[5], s[4], s[3]]) / vol        # calculator.py:515-526 + /volume (:608)
        self.results = {'free_energy': energy, 'energy': energy, 'forces': forces, 'stress': stress}
        return self.results


class SevenNetD3Calculator(_Base):
    """``SevenNetCalculator`` + ``D3Calculator`` summed (the reference builds an ASE SumCalculator,
    ``sevenn/calculator.py:236-314``; without ASE the two result dicts are added here)."""
    implemented_properties = ['free_energy', 'energy', 'energies', 'forces', 'stress']

    def __init__(self, model='7net-0', file_type: str = 'checkpoint', device='auto', modal=None, enable_cueq=False,
                 enable_flash=False, enable_oeq=False, sevennet_config=None, damping_type: str = 'damp_bj',
                 functional_name: str = 'pbe', vdw_cutoff: float = 9000, cn_cutoff: float = 1600, **kwargs):
        from .calculator import SevenNetCalculator
        super().__init__()
        self.d3_calc = D3Calculator(damping_type=damping_type, functional_name=functional_name, vdw_cutoff=vdw_cutoff,
                                    cn_cutoff=cn_cutoff)
        self.sevennet_calc = SevenNetCalculator(model=model, file_type=file_type, device=device, modal=modal,
                                                enable_cueq=enable_cueq, enable_flash=enable_flash, enable_oeq=enable_oeq,
                                                sevennet_config=sevennet_config, **kwargs)

    def calculate(self, atoms=None, properties=None, system_changes=_all_changes):
        super().calculate(atoms, properties, system_changes)
        a = self.sevennet_calc.calculate(atoms, properties, system_changes)
        b = self.d3_calc.calculate(atoms, properties, system_changes)
        out = dict(a)
        for k in ('free_energy', 'energy', 'forces'):
            out[k] = a[k] + b[k]
        if 'stress' in a:
            out['stress'] = a['stress'] + b['stress']
        self.results = out
        return out
