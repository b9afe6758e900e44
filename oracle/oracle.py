"""CPU ORACLE -- test infrastructure, NOT product code.

A plain-torch (CPU, fp32 or fp64) restatement of the reference's per-MD-step energy/force
path, module by module, in the reference's own data layout (e3nn ``mul_ir``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file; nothing under ``sevenn_b200/`` does.

Parity is PINNED: ``tests/test_oracle_golden.py`` checks this oracle against every golden
vector the reference's own tests hold for this path (``tests/unit_tests/test_pretrained.py:75-164``,
``tests/unit_tests/test_calculator.py:56-104,240-266``, ``tests/data/inferences/snet0_on_hfo2``).
The third-party arithmetic the reference delegates to ``e3nn>=0.5.0`` (un-vendored, unpinned:
reference ``pyproject.toml:24``) is restated from its published definitions in
``sevenn_b200/cg.py`` (Wigner 3j) and ``sevenn_b200/sh.py`` (spherical harmonics).

Reference order of modules (``sevenn/model_build.py:448-616``,
``sevenn/nn/interaction_blocks.py:41-76``):
  edge_embedding -> onehot -> onehot_to_feature_x ->
  for t: {t}_self_connection_intro, {t}_self_interaction_1, {t}_convolution,
         {t}_self_interaction_2, {t}_self_connection_outro, {t}_equivariant_gate
  -> reduce_input_to_hidden -> reduce_hidden_to_energy -> rescale_atomic_energy
  -> reduce_total_enegy -> force_output
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from sevenn_b200.cg import tp_path_coefficients
from sevenn_b200.sh import sh_polynomials, X, Y, Z
from sevenn_b200.spec import SILU_NORM, ModelSpec, build_spec, irreps_offsets


def _sh_torch(lmax: int, u: torch.Tensor) -> torch.Tensor:
    """e3nn SphericalHarmonics(normalize=True, 'component') on unit vectors u [E,3]
    (``sevenn/nn/edge_embedding.py:184-185``).  Evaluates the sympy polynomial table."""
    x, y, z = u[:, 0], u[:, 1], u[:, 2]
    cols = []
    for expr in sh_polynomials(lmax):
        acc = torch.zeros_like(x)
        for (a, b, c), coef in expr.as_poly(X, Y, Z).terms():
            acc = acc + float(coef) * x ** a * y ** b * z ** c
        cols.append(acc)
    return torch.stack(cols, dim=-1)


class Oracle:
    def __init__(self, meta: dict, arrays: Dict[str, np.ndarray], dtype=torch.float64, device='cpu'):
        self.meta = meta
        self.device = torch.device(device)
        self.spec: ModelSpec = build_spec(meta)
        self.dtype = dtype
        self.w = {k: torch.as_tensor(np.asarray(v), dtype=dtype, device=self.device) for k, v in arrays.items()}
        self.cg = {}
        for L in self.spec.layers:
            for p in L.paths:
                key = (p.l1, p.l2, p.l3)
                if key not in self.cg:
                    self.cg[key] = torch.as_tensor(tp_path_coefficients(*key), dtype=dtype, device=self.device)

    # ---- pieces -------------------------------------------------------------------------
    def edge_embedding(self, edge_vec: torch.Tensor):
        """``EdgeEmbedding.forward`` (edge_embedding.py:207-217): Bessel (:101-103) x cutoff
        (XPLOR :150-160 or polynomial :125-132) and spherical harmonics."""
        s = self.spec
        r = torch.linalg.norm(edge_vec, dim=-1)
        ur = r.unsqueeze(-1)
        bessel = (2.0 / s.cutoff) * torch.sin(self.w['bessel_coeffs'] * ur) / ur
        if s.cutoff_fn == 'XPLOR':
            r_on, r_cut = s.cutoff_on, s.cutoff
            r2, on2, cut2 = r * r, r_on * r_on, r_cut * r_cut
            env = torch.where(
                r < r_on, torch.ones_like(r),
                (cut2 - r2) ** 2 * (cut2 + 2 * r2 - 3 * on2) / (cut2 - on2) ** 3)
        else:
            p = float(s.poly_p)
            x = r / s.cutoff
            env = (1 - (p + 1.0) * (p + 2.0) / 2.0 * torch.pow(x, p)
                   + p * (p + 2.0) * torch.pow(x, p + 1.0)
                   - p * (p + 1.0) / 2.0 * torch.pow(x, p + 2.0))
        emb = bessel * env.unsqueeze(-1)
        sh = _sh_torch(s.lmax_filter, edge_vec / ur)
        return r, emb, sh

    def linear(self, x: torch.Tensor, flat_w: torch.Tensor, in_blocks, out_muls) -> torch.Tensor:
        """e3nn ``o3.Linear`` without bias (``sevenn/nn/linear.py:94-100``).
        ``in_blocks``: list of (l, mul, offset) input entries in order; ``out_muls[l]`` output
        multiplicity.  Weight blocks are stored i_in-major, each (mul_in, mul_out) row-major;
        out = sum over input entries of x_blk^T W_blk / sqrt(total fan-in of that output)."""
        n = x.shape[0]
        fan = [0] * len(out_muls)
        for (l, mul, _) in in_blocks:
            if l < len(out_muls):
                fan[l] += mul
        outs = [torch.zeros(n, out_muls[l], 2 * l + 1, dtype=x.dtype, device=x.device) for l in range(len(out_muls))]
        woff = 0
        for (l, mul, off) in in_blocks:
            if l >= len(out_muls) or out_muls[l] == 0:
                continue
            d = 2 * l + 1
            W = flat_w[woff:woff + mul * out_muls[l]].reshape(mul, out_muls[l])
            woff += mul * out_muls[l]
            xb = x[:, off:off + mul * d].reshape(n, mul, d)
            outs[l] = outs[l] + torch.einsum('uw,nui->nwi', W, xb) / math.sqrt(fan[l])
        assert woff == flat_w.numel(), (woff, flat_w.numel())
        return torch.cat([o.reshape(n, o.shape[1] * o.shape[2]) for o in outs], dim=1)

    @staticmethod
    def _blocks(muls) -> List[tuple]:
        return [(l, m, o) for l, (m, o) in enumerate(zip(muls, irreps_offsets(list(muls))))]

    def radial_mlp(self, t: int, emb: torch.Tensor) -> torch.Tensor:
        """e3nn ``FullyConnectedNet([8,64,64,W], silu)`` (``convolution.py:93-95,121``;
        SURVEY A.7): h = c*silu(h W / sqrt(fan_in)); last layer linear."""
        n_mlp = len(self.spec.radial_hidden) + 1
        h = emb
        for j in range(n_mlp):
            W = self.w[f'{t}.mlp{j}']
            h = h @ W / math.sqrt(W.shape[0])
            if j < n_mlp - 1:
                h = SILU_NORM * torch.nn.functional.silu(h)
        return h

    def tensor_product(self, L, x_src: torch.Tensor, sh: torch.Tensor, weight: torch.Tensor):
        """e3nn ``TensorProduct`` with 'uvu' instructions and per-edge weights
        (``convolution.py:84-91,131``).  x_src [E,dim_x] mul_ir, sh [E,n_sh],
        weight [E,W] -> message [E,dim_mid] mul_ir, slots sorted by l3."""
        E = x_src.shape[0]
        xoff = irreps_offsets(list(L.x_muls))
        out = []
        for p in L.paths:
            d1, d2 = 2 * p.l1 + 1, 2 * p.l2 + 1
            xb = x_src[:, xoff[p.l1]:xoff[p.l1] + p.mul * d1].reshape(E, p.mul, d1)
            yb = sh[:, p.l2 * p.l2:p.l2 * p.l2 + d2]
            m = torch.einsum('ijk,eui,ej->euk', self.cg[(p.l1, p.l2, p.l3)], xb, yb)
            m = m * weight[:, p.w_off:p.w_off + p.mul].unsqueeze(-1)
            out.append(m.reshape(E, p.mul * (2 * p.l3 + 1)))
        return torch.cat(out, dim=1)

    def gate(self, L, g: torch.Tensor) -> torch.Tensor:
        """e3nn ``Gate`` (``equivariant_gate.py:57-59``): input is
        (scalars | gates) x0e + gated irreps; silu is rescaled by SILU_NORM."""
        n = g.shape[0]
        ns = L.n_scalars
        act = SILU_NORM * torch.nn.functional.silu(g[:, :L.gate_muls[0]])
        outs = [act[:, :ns]]
        goff = irreps_offsets(list(L.gate_muls))
        gate_off = ns
        for l in range(1, len(L.out_muls)):
            mul, d = L.out_muls[l], 2 * l + 1
            blk = g[:, goff[l]:goff[l] + mul * d].reshape(n, mul, d)
            outs.append((blk * act[:, gate_off:gate_off + mul].unsqueeze(-1)).reshape(n, mul * d))
            gate_off += mul
        return torch.cat(outs, dim=1)

    # ---- full step ----------------------------------------------------------------------
    def forward(self, species: np.ndarray, edge_index: np.ndarray, edge_vec: np.ndarray,
                volume: Optional[float] = None, keep: bool = False,
                x_ghost_map: Optional[np.ndarray] = None,
                edge_chunk: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """species [N] species indices; edge_index [2,E] with [0]=centre i (aggregation
        target), [1]=neighbour j; edge_vec [E,3] = r_j - r_i + shift.

        ``edge_chunk``: evaluate each convolution in chunks of that many edges under
        ``torch.utils.checkpoint`` (same arithmetic, the per-edge intermediates are recomputed in the
        backward instead of stored) so that large cells -- the 12 000-atom benchmark cell in fp64 --
        fit in memory; results equal the unchunked evaluation to rounding (tests/test_oracle_golden.py)."""
        s, dt = self.spec, self.dtype
        dev = self.device
        species_t = torch.as_tensor(np.asarray(species), dtype=torch.long, device=dev)
        dst = torch.as_tensor(np.asarray(edge_index[0]), dtype=torch.long, device=dev)
        src = torch.as_tensor(np.asarray(edge_index[1]), dtype=torch.long, device=dev)
        ev = torch.as_tensor(np.asarray(edge_vec), dtype=dt, device=dev).clone().requires_grad_(True)
        n = species_t.shape[0]
        saved: Dict[str, torch.Tensor] = {}

        r, emb, sh = self.edge_embedding(ev)
        # onehot_to_feature_x: Linear(S x0e -> mul0 x0e) on a one-hot = row lookup / sqrt(S)
        W_e = self.w['embed'].reshape(s.num_species, -1)
        x = W_e[species_t] / math.sqrt(s.num_species)
        if keep:
            saved.update(edge_embedding=emb, edge_attr=sh, x_embed=x)

        for L in s.layers:
            t = L.t
            xb = self._blocks(L.x_muls)
            sc = self.linear(x, self.w[f'{t}.sc'], xb, list(L.gate_muls))         # self_connection_intro
            x = self.linear(x, self.w[f'{t}.si1'], xb, list(L.x_muls))            # self_interaction_1
            if keep:
                saved[f'{t}.x_si1'] = x
            if edge_chunk is None or keep:
                weight = self.radial_mlp(t, emb)
                msg = self.tensor_product(L, x[src], sh, weight)                  # convolution
                agg = torch.zeros(n, L.dim_mid, dtype=dt, device=dev).index_add_(0, dst, msg)
            else:
                from torch.utils.checkpoint import checkpoint

                def conv_chunk(x_, emb_c, sh_c, src_c, dst_c, L=L, t=t):
                    w_c = self.radial_mlp(t, emb_c)
                    m_c = self.tensor_product(L, x_[src_c], sh_c, w_c)
                    return torch.zeros(n, L.dim_mid, dtype=dt, device=dev).index_add_(0, dst_c, m_c)

                agg = torch.zeros(n, L.dim_mid, dtype=dt, device=dev)
                for c0 in range(0, src.shape[0], int(edge_chunk)):
                    c1 = min(c0 + int(edge_chunk), src.shape[0])
                    agg = agg + checkpoint(conv_chunk, x, emb[c0:c1], sh[c0:c1], src[c0:c1], dst[c0:c1],
                                           use_reentrant=False)
                weight = None
            agg = agg / self.w[f'{t}.den']
            if keep:
                saved[f'{t}.weight'] = weight
                saved[f'{t}.mid'] = agg
            mid_blocks, off = [], 0
            for p in L.paths:
                mid_blocks.append((p.l3, p.mul, off))
                off += p.mul * (2 * p.l3 + 1)
            g = self.linear(agg, self.w[f'{t}.si2'], mid_blocks, list(L.gate_muls))  # self_interaction_2
            g = g + sc                                                             # self_connection_outro
            x = self.gate(L, g)                                                    # equivariant_gate
            if keep:
                saved[f'{t}.gate_in'] = g
                saved[f'{t}.x_out'] = x

        Lz = s.layers[-1]
        hb = self._blocks(Lz.out_muls)
        h = self.linear(x, self.w['readout1'], hb, [s.readout_hidden])
        e = self.linear(h, self.w['readout2'], [(0, s.readout_hidden, 0)], [1])
        atomic_e = e[:, 0] * self.w['scale'][species_t] + self.w['shift'][species_t]  # scale.py:155-162
        total = atomic_e.sum()                                                        # linear.py:127-141

        # ForceStressOutputFromEdge (force_output.py:171-230)
        if ev.shape[0] > 0:
            (fij,) = torch.autograd.grad(total, ev, allow_unused=True)
            if fij is None:
                fij = torch.zeros_like(ev)
        else:
            fij = torch.zeros_like(ev)
        evd = ev.detach()
        forces = torch.zeros(n, 3, dtype=dt, device=dev).index_add_(0, dst, fij) \
            - torch.zeros(n, 3, dtype=dt, device=dev).index_add_(0, src, fij)
        vir = torch.stack([evd[:, 0] * fij[:, 0], evd[:, 1] * fij[:, 1], evd[:, 2] * fij[:, 2],
                           evd[:, 0] * fij[:, 1], evd[:, 1] * fij[:, 2], evd[:, 2] * fij[:, 0]],
                          dim=-1)
        atomic_virial = -torch.zeros(n, 6, dtype=dt, device=dev).index_add_(0, src, vir)
        out = dict(energy=total.detach(), atomic_energy=atomic_e.detach(), forces=forces,
                   edge_force=fij, atomic_virial=atomic_virial,
                   virial=-vir.sum(0))
        if volume is not None and volume > 0:
            out['stress'] = -vir.sum(0) / volume   # 'inferred_stress' xx,yy,zz,xy,yz,zx
        if keep:
            out['saved'] = {k: v.detach() for k, v in saved.items()}
        return out


def ase_voigt_stress(inferred_stress: np.ndarray) -> np.ndarray:
    """``SevenNetCalculator.output_to_results`` (calculator.py:198-203)."""
    return -np.asarray(inferred_stress)[[0, 1, 2, 4, 5, 3]]
