"""LAMMPS ML-IAP (unified) front-end on the engine's stage API.

Mirrors the reference adapter ``sevenn/mliap.py:76-253`` (``SevenNetMLIAPWrapper``): LAMMPS hands over
``nlocal`` owned atoms plus ghosts, the pair list ``pair_i, pair_j, rij`` and two communication
callbacks; the wrapper returns atomic energies, the total energy and the *pair* forces
``dE/d rij`` (``update_pair_forces_gpu``).  The reference wraps every convolution in a
``forward_exchange`` / autograd ``reverse_exchange`` pair (``sevenn/nn/_ghost_exchange.py:12-47``); here
the same two callbacks are placed between the engine's stages exactly where
``sevenn_b200.parallel.DistributedRunner`` places its NCCL exchanges:

    FWD_BEGIN, then per layer t: FWD_LAYER(t)  -> forward_exchange(x[t+1])      (t + 1 < T)
    FWD_END,   then per layer t (reversed): BWD_LAYER_A(t) -> reverse_exchange(dx[t]) -> BWD_LAYER_B(t)   (t > 0)
    BWD_END -> edge forces

Layer 0 needs no exchange: ghost species are known locally.  ``lammps.mliap`` is not installed in this
image; the base class is ``MLIAPUnified`` when importable and ``object`` otherwise, and the protocol is
tested on the CPU with a stand-in engine and a stand-in ``lmp_data`` (``tests/test_mliap_cpu.py``).
"""
from __future__ import annotations

from typing import Any

from .engine import (STAGE_BWD_END, STAGE_BWD_LAYER_A, STAGE_BWD_LAYER_B, STAGE_FWD_BEGIN, STAGE_FWD_END,
                     STAGE_FWD_LAYER)

try:  # pragma: no cover - LAMMPS python package is not in the build image
    from lammps.mliap.mliap_unified_abc import MLIAPUnified as _Base
except Exception:  # noqa: BLE001
    _Base = object

_SYMBOLS = ('X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr '
            'Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu '
            'Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu').split()


class SevenNetMLIAPWrapper(_Base):
    """``model_path``: pretrained name, ``.npz`` of this repo, or a reference checkpoint.  As in the
    reference, ``element_types`` is the periodic table indexed by atomic number with ``'X'`` for elements
    the model does not know, so ``lmp_data.elems`` are atomic numbers (``mliap.py:131-137``)."""

    def __init__(self, model_path: str, engine=None, **kwargs: Any):
        if _Base is not object:
            super().__init__()
        if kwargs.get('modal') is not None:
            raise NotImplementedError('multi-fidelity models are out of scope')
        if kwargs.get('use_cueq') or kwargs.get('use_flash') or kwargs.get('use_oeq'):
            raise ValueError('use_cueq/flash/oeq select other accelerators; this wrapper always runs the sevenn_b200 engine')
        self.model_path = model_path
        self.engine = engine                      # lazily built on the first compute_forces (mliap.py:144-149)
        self._radial = kwargs.get('radial', 'table')
        if engine is None:
            from .calculator import resolve_model
            self._model = resolve_model(model_path)
            meta = self._model[0]
            self.cutoff = float(meta['cutoff'])
            known = {int(z) for z in meta['type_map']}
        else:
            self._model = None
            self.cutoff = float(getattr(engine.spec, 'cutoff', 0.0))
            known = set(getattr(engine.spec, 'type_map', {}) or range(len(_SYMBOLS)))
        self.rcutfac = 0.5 * self.cutoff
        self.element_types = [s if z in known else 'X' for z, s in enumerate(_SYMBOLS)]
        self.ndescriptors = int(kwargs.get('ndescriptors', 1))
        self.nparams = int(kwargs.get('nparams', 1))

    def _ensure_engine(self):
        if self.engine is None:
            from .engine import B200Engine
            self.engine = B200Engine(self._model[0], self._model[1], radial=self._radial)
        return self.engine

    def compute_forces(self, lmp_data):
        eng = self._ensure_engine()
        torch = eng.torch if hasattr(eng, 'torch') else __import__('torch')
        nlocal, ntotal = int(lmp_data.nlocal), int(lmp_data.ntotal)
        if nlocal == 0:
            return
        dev = eng.device
        no_pairs = int(lmp_data.npairs) <= 1           # same guard as mliap.py:186
        if no_pairs:
            edge_index = torch.zeros(2, 0, dtype=torch.int64, device=dev)
            rij = torch.zeros(0, 3, dtype=torch.float32, device=dev)
        else:
            edge_index = torch.stack([torch.as_tensor(lmp_data.pair_i, device=dev).long(),
                                      torch.as_tensor(lmp_data.pair_j, device=dev).long()])
            rij = torch.as_tensor(lmp_data.rij, device=dev).to(torch.float32)
        elems = torch.as_tensor(lmp_data.elems, device=dev).long()
        tm = getattr(eng.spec, 'type_map', None)
        if tm:
            lut = torch.full((len(_SYMBOLS),), -1, dtype=torch.int64, device=dev)
            for z, idx in tm.items():
                lut[int(z)] = int(idx)
            species = lut[elems]
            if bool((species < 0).any()):
                raise ValueError('an element of the LAMMPS system is not known to this model')
        else:
            species = elems
        graph = eng.set_graph(species.to(torch.int32), edge_index, rij, n_local=nlocal)
        perm = graph.get('perm') if isinstance(graph, dict) else None

        spec, T = eng.spec, eng.spec.n_layers

        def exchanged(name, t, width, call):
            buf = eng.buffer(name, t, shape=(ntotal, width))
            out = torch.empty_like(buf)
            call(buf, out, width)
            return buf, out

        eng.run_stage(STAGE_FWD_BEGIN)
        for t in range(T):
            eng.run_stage(STAGE_FWD_LAYER, t)
            if t + 1 < T:
                # called on EVERY rank and layer, as the reference does (mliap.py:176-247 wraps every
                # convolution): forward_exchange is LAMMPS forward_comm, a collective over the ranks -- a
                # rank without ghosts may still own atoms that are ghosts elsewhere and must post its sends
                buf, out = exchanged('x', t + 1, spec.layers[t + 1].dim_x, lmp_data.forward_exchange)
                if ntotal > nlocal:
                    buf[nlocal:] = out[nlocal:]        # ghost rows <- their owners' features
        eng.run_stage(STAGE_FWD_END)
        for t in range(T - 1, -1, -1):
            eng.run_stage(STAGE_BWD_LAYER_A, t)
            if t > 0:
                buf, out = exchanged('dx', t, spec.layers[t].dim_x, lmp_data.reverse_exchange)   # reverse_comm: collective too
                buf[:nlocal] = out[:nlocal]            # owners <- own + ghost-row contributions
                eng.run_stage(STAGE_BWD_LAYER_B, t)
        eng.run_stage(STAGE_BWD_END)

        e_atoms = eng.buffer('atomic_energy', shape=(nlocal,))
        eatoms = torch.as_tensor(lmp_data.eatoms)
        eatoms.copy_(e_atoms.to(eatoms.device, eatoms.dtype))
        lmp_data.energy = e_atoms.double().sum().detach()
        if no_pairs:
            fij = torch.zeros(0, 3, dtype=torch.float64, device=dev)
        else:
            fe = eng.buffer('edge_force', shape=(edge_index.shape[1], 3))
            if perm is not None:                       # the engine sorted the pairs by centre: undo
                fij = torch.empty_like(fe)
                fij[perm] = fe
            else:
                fij = fe
        lmp_data.update_pair_forces_gpu(fij.to(torch.float64))   # upcast as the reference does (mliap.py:245)

    def compute_descriptors(self, lmp_data):
        pass

    def compute_gradients(self, lmp_data):
        pass
