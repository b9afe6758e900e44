"""ASE-style calculator with the reference's ``SevenNetCalculator`` surface
(``sevenn/calculator.py:20-233``): same constructor keywords, same ``results`` keys
(``free_energy, energy, energies, forces, stress, num_edges``), same stress convention
(ASE Voigt order ``-inferred_stress[[0,1,2,4,5,3]]``, ``calculator.py:198-203``).

ASE itself is optional: when importable the class derives from ``ase.calculators.calculator.
Calculator``; otherwise it is a duck-typed object with ``calculate(atoms, properties,
system_changes)`` and ``get_*`` helpers, where ``atoms`` needs ``get_positions()``,
``get_cell()``, ``get_pbc()`` and ``get_atomic_numbers()``.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .checkpoint import convert_reference_checkpoint, load_weights
from .engine import B200Engine

_WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'weights')
# names the reference resolves in sevenn/util.py:264-312
_ALIASES = {
    '7net-0': 'sevennet_0', '7net-0_11july2024': 'sevennet_0', 'sevennet-0': 'sevennet_0',
    'sevennet_0': 'sevennet_0', '7net-l3i5': 'sevennet_l3i5', 'sevennet-l3i5': 'sevennet_l3i5',
    'sevennet_l3i5': 'sevennet_l3i5',
}

try:  # pragma: no cover - ASE is not installed in the build container
    from ase.calculators.calculator import Calculator as _Base, all_changes as _all_changes
except Exception:  # noqa: BLE001
    _all_changes = ['positions', 'numbers', 'cell', 'pbc', 'initial_charges', 'initial_magmoms']

    class _Base:  # minimal stand-in for ase.calculators.calculator.Calculator
        def __init__(self, **kwargs):
            self.results = {}
            self.atoms = None

        def calculate(self, atoms=None, properties=None, system_changes=None):
            self.atoms = atoms

        def get_potential_energy(self, atoms=None, force_consistent=False):
            self.calculate(atoms)
            return self.results['free_energy' if force_consistent else 'energy']

        def get_forces(self, atoms=None):
            self.calculate(atoms)
            return self.results['forces']

        def get_stress(self, atoms=None):
            self.calculate(atoms)
            return self.results['stress']

        def get_potential_energies(self, atoms=None):
            self.calculate(atoms)
            return self.results['energies']


def resolve_model(model: str):
    """name | path to ``.npz`` (this repo's format) | path to a reference ``.pth`` checkpoint."""
    key = str(model).lower()
    if key in _ALIASES:
        return load_weights(os.path.join(_WEIGHTS_DIR, _ALIASES[key] + '.npz'))
    if os.path.isfile(model):
        if str(model).endswith('.npz'):
            return load_weights(model)
        return convert_reference_checkpoint(model, os.path.splitext(os.path.basename(model))[0])
    raise ValueError(f'unknown model {model!r}: expected one of {sorted(_ALIASES)} or a file path')


class SevenNetCalculator(_Base):
    implemented_properties = ['free_energy', 'energy', 'forces', 'stress', 'energies']

    def __init__(self, model: str = '7net-0', file_type: str = 'checkpoint', device='cuda',
                 modal: Optional[str] = None, enable_cueq: bool = False, enable_flash: bool = False,
                 enable_oeq: bool = False, compute_atomic_virial: bool = False,
                 sevennet_config: Optional[dict] = None, radial: str = 'table', **kwargs):
        super().__init__(**kwargs)
        if file_type != 'checkpoint':
            raise NotImplementedError("sevenn_b200 loads checkpoints only (file_type='checkpoint')")
        if modal is not None:
            raise NotImplementedError('multi-fidelity models are out of scope')
        if enable_cueq or enable_flash or enable_oeq:
            raise ValueError('enable_cueq/flash/oeq select other accelerators; this calculator '
                             'always runs the sevenn_b200 CUDA engine')
        import torch
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError('sevenn_b200 has no CPU path; pass a CUDA device')
        self.meta, self.arrays = resolve_model(model) if isinstance(model, str) else model
        self.engine = B200Engine(self.meta, self.arrays, radial=radial,
                                 device=dev.index if dev.index is not None else None,
                                 atomic_virial=compute_atomic_virial)
        self.cutoff = self.engine.spec.cutoff
        self.type_map = self.engine.spec.type_map
        self.compute_atomic_virial = compute_atomic_virial
        self.sevennet_config = sevennet_config or dict(self.meta)

    def calculate(self, atoms=None, properties=None, system_changes=_all_changes):
        super().calculate(atoms, properties, system_changes)
        if atoms is None:
            raise ValueError('No atoms to evaluate')
        pos = np.asarray(atoms.get_positions(), dtype=np.float64)
        cell = np.asarray(atoms.get_cell(), dtype=np.float64).reshape(3, 3)
        pbc = np.asarray(atoms.get_pbc(), dtype=bool)
        numbers = np.asarray(atoms.get_atomic_numbers())
        try:
            species = np.array([self.type_map[int(z)] for z in numbers], dtype=np.int32)
        except KeyError as e:  # same failure mode as sequential.py:131-137 for unknown elements
            raise ValueError(f'atomic number {e} is not known to this model') from None
        # neighbour list, graph build, model and force path all run on the GPU (one C-ABI call);
        # the reference builds the graph on the CPU every step (calculator.py:224-226)
        energy, energies, forces, virial, n_edges = self.engine.compute_positions(species, pos, cell, pbc)
        # the reference divides by atoms.cell.volume whatever the pbc flags are (dataload.py:121,
        # force_output.py:227-228); a missing cell (volume 0) has no stress
        vol = abs(np.linalg.det(cell))
        self.results = {
            'free_energy': energy, 'energy': energy,
            'energies': energies.astype(np.float64),
            'forces': forces.astype(np.float64),
            'num_edges': n_edges,
        }
        if vol > 0:
            inferred_stress = virial / vol
            self.results['stress'] = -inferred_stress[[0, 1, 2, 4, 5, 3]]
        if self.compute_atomic_virial:   # calculator.py:211-216: 'stresses' = the per-atom virial as the model gives it
            self.results['stresses'] = self.engine.buffer('atomic_virial', shape=(len(numbers), 6)).cpu().numpy().astype(np.float64)
        return self.results
