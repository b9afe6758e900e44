"""CPU tests of the multi-rank path (world_size 2 and 4, gloo): brick decomposition, ghost maps
and the stage/exchange protocol of sevenn_b200.parallel.DistributedRunner, driven with a small
CPU stand-in for the CUDA engine that has the same stage interface (a linear-plus-tanh message
passing model whose serial result is the oracle for the distributed one)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sevenn_b200.engine import (STAGE_BWD_END, STAGE_BWD_LAYER_A, STAGE_BWD_LAYER_A1, STAGE_BWD_LAYER_A2,
                                STAGE_BWD_LAYER_B, STAGE_BWD_LAYER_B1, STAGE_BWD_LAYER_B2, STAGE_FWD_BEGIN,
                                STAGE_FWD_CONV_INTERIOR, STAGE_FWD_END, STAGE_FWD_LAYER, STAGE_FWD_LAYER_A,
                                STAGE_FWD_LAYER_A2, STAGE_FWD_LAYER_SC)
from sevenn_b200.neighbors import build_graph, diamond_si, rocksalt_nacl
from sevenn_b200.parallel import DistributedRunner, brick_decompose


class _L:
    def __init__(self, d):
        self.dim_x = d


class _Spec:
    def __init__(self, n_layers, d):
        self.n_layers = n_layers
        self.layers = [_L(d) for _ in range(n_layers)]


class FakeEngine:
    """Same stage protocol as B200Engine, trivial physics:
       x_0 = (species+1); a_t[i] = sum_{e->i} |v_e| x_t[src_e]; h = tanh(a); x_{t+1} = c_t h; E = sum h_T."""
    D, T = 4, 3

    def __init__(self):
        self.spec = _Spec(self.T, self.D)
        self.device = torch.device('cpu')
        self.coef = [0.3, 0.2, 0.1]

    def set_graph(self, species, edge_index, edge_vec, n_local=None):
        self.species = torch.as_tensor(species).double()
        ei = torch.as_tensor(edge_index).long()
        self.dst, self.src = ei[0], ei[1]
        self.vec = torch.as_tensor(edge_vec).double()
        self.n_nodes = len(self.species)
        self.n_local = self.n_nodes if n_local is None else n_local
        self.w = self.vec.norm(dim=1)
        self.x = [torch.zeros(self.n_nodes, self.D, dtype=torch.float64) for _ in range(self.T)]
        self.a = [None] * self.T
        self.dx = torch.zeros(self.n_nodes, self.D, dtype=torch.float64)
        self.forces = torch.zeros(self.n_nodes, 3, dtype=torch.float64)
        self.energy = torch.zeros(1, dtype=torch.float64)
        self.virial = torch.zeros(6, dtype=torch.float64)
        self.dEdw = torch.zeros(len(self.w), dtype=torch.float64)
        self.n_interior = self.n_local

    def set_interior(self, n_interior):
        self.n_interior = int(n_interior)

    def buffer(self, name, t=0, dtype='f4', shape=None):
        return {'x': lambda: self.x[t], 'dx': lambda: self.dx, 'forces': lambda: self.forces,
                'energy': lambda: self.energy, 'virial': lambda: self.virial,
                'atomic_energy': lambda: self.h.sum(1), 'edge_force': lambda: self.fedge}[name]()

    def run_stage(self, stage, t=0):
        nl = self.n_local
        if stage == STAGE_FWD_BEGIN:
            self.x[0][:] = (self.species + 1.0)[:, None]
            self.dEdw.zero_()
        elif stage in (STAGE_FWD_LAYER_SC, STAGE_BWD_LAYER_B1):
            pass        # the stand-in has no self-connection term
        elif stage in (STAGE_FWD_LAYER, STAGE_FWD_LAYER_A, STAGE_FWD_CONV_INTERIOR, STAGE_FWD_LAYER_A2):
            # centre ranges as in the CUDA engine: interior = [0, n_interior), boundary = the rest
            m = torch.ones_like(self.dst, dtype=torch.bool)
            if stage == STAGE_FWD_CONV_INTERIOR:
                m = self.dst < self.n_interior
                assert bool((self.src[m] < nl).all())       # interior atoms never read a ghost row
            elif stage == STAGE_FWD_LAYER_A2:
                m = self.dst >= self.n_interior
            part = torch.zeros(nl, self.D, dtype=torch.float64).index_add_(0, self.dst[m], self.w[m, None] * self.x[t][self.src[m]])
            if stage == STAGE_FWD_CONV_INTERIOR:
                self.a[t] = part
                return
            a = part if stage != STAGE_FWD_LAYER_A2 else self.a[t] + part
            self.a[t] = a
            self.h = torch.tanh(a)
            if t + 1 < self.T:
                self.x[t + 1][:nl] = self.coef[t] * self.h
        elif stage == STAGE_FWD_END:
            self.energy[0] = self.h.sum()
            self.dh = torch.ones(nl, self.D, dtype=torch.float64)
        elif stage in (STAGE_BWD_LAYER_A, STAGE_BWD_LAYER_A1, STAGE_BWD_LAYER_A2):
            m = torch.ones_like(self.dst, dtype=torch.bool)
            if stage == STAGE_BWD_LAYER_A1:
                m = self.dst >= self.n_interior
            elif stage == STAGE_BWD_LAYER_A2:
                m = self.dst < self.n_interior
            if stage != STAGE_BWD_LAYER_A2:
                self.da = self.dh * (1 - torch.tanh(self.a[t]) ** 2)
                self.dx.zero_()
            da = self.da
            self.dEdw[m] += (da[self.dst[m]] * self.x[t][self.src[m]]).sum(1)
            if t > 0:
                self.dx.index_add_(0, self.src[m], self.w[m, None] * da[self.dst[m]])
        elif stage in (STAGE_BWD_LAYER_B, STAGE_BWD_LAYER_B2):
            self.dh = self.coef[t - 1] * self.dx[:nl]
        elif stage == STAGE_BWD_END:
            f = self.dEdw[:, None] * self.vec / self.w[:, None]
            self.fedge = f
            self.forces.zero_()
            self.forces.index_add_(0, self.dst, f)
            self.forces.index_add_(0, self.src, -f)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _system(kind):
    if kind == 'si':
        pos, cell, z = diamond_si(3, 2, 2, seed=4)
    elif kind == 'si_long':                                  # bricks wide enough to have interior atoms
        pos, cell, z = diamond_si(6, 2, 2, seed=5)
    else:
        pos, cell, z = rocksalt_nacl(2, 2, 2, sigma=0.08, seed=7)
    species = (z == z.min()).astype(np.int32)
    return pos, cell, species


def _worker(rank, world, port, grid, kind, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pos, cell, species = _system(kind)
        part = brick_decompose(pos, cell, species, grid, rank, 5.0)
        assert 0 <= part['n_interior'] <= part['n_local']
        nl, ni = part['n_local'], part['n_interior']
        ghost_edge = part['edge_index'][1] >= nl
        assert (part['edge_index'][0][ghost_edge] >= ni).all()       # only boundary atoms have ghost neighbours
        assert set(np.unique(part['edge_index'][0][ghost_edge])) == set(range(ni, nl)) or ni == nl
        eng = FakeEngine()
        run = DistributedRunner(eng, part)
        assert run.split
        run.compute()
        if kind == 'si_long':
            assert 0 < ni < nl
        q.put((rank, part['global_ids'][:part['n_local']], eng.forces[:part['n_local']].numpy().copy(),
               float(eng.energy[0]), part['edge_index'].shape[1], part['n_nodes'] - part['n_local']))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,grid,kind', [(2, (2, 1, 1), 'si'), (2, (1, 1, 2), 'nacl'), (4, (2, 2, 1), 'si'),
                                             (2, (2, 1, 1), 'si_long')])
def test_distributed_protocol_matches_serial(world, grid, kind):
    pos, cell, species = _system(kind)
    ei, ev = build_graph(pos, cell, True, 5.0)
    ser = FakeEngine()
    ser.set_graph(species, ei, ev)
    for st, ts in [(STAGE_FWD_BEGIN, [0]), (STAGE_FWD_LAYER, range(ser.T)), (STAGE_FWD_END, [0])]:
        for t in ts:
            ser.run_stage(st, t)
    for t in range(ser.T - 1, -1, -1):
        ser.run_stage(STAGE_BWD_LAYER_A, t)
        if t > 0:
            ser.run_stage(STAGE_BWD_LAYER_B, t)
    ser.run_stage(STAGE_BWD_END)

    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    forces = np.zeros((len(pos), 3))
    seen = np.zeros(len(pos), dtype=int)
    n_edges = 0
    for rank, gids, f, energy, e_loc, n_ghost in res:
        forces[gids] = f
        seen[gids] += 1
        n_edges += e_loc
        assert abs(energy - float(ser.energy[0])) < 1e-9 * abs(float(ser.energy[0]))   # all-reduced
        assert n_ghost > 0
    assert (seen == 1).all()                       # every atom owned exactly once
    assert n_edges == ei.shape[1]                  # every directed edge appears on exactly one rank
    assert np.allclose(forces, ser.forces.numpy(), atol=1e-9)


def test_brick_decompose_invariants():
    pos, cell, species = _system('si')
    ei, ev = build_graph(pos, cell, True, 5.0)
    parts = [brick_decompose(pos, cell, species, (2, 1, 1), r, 5.0) for r in range(2)]
    assert sum(p['n_local'] for p in parts) == len(pos)
    for r, p in enumerate(parts):
        nl = p['n_local']
        assert (np.diff(p['edge_index'][0]) >= 0).all() and p['edge_index'][0].max() < nl
        assert (np.diff(p['ghost_owner']) >= 0).all() and (p['ghost_owner'] != r).all()
        gids = p['global_ids']
        assert len(np.unique(gids)) == len(gids)
        g_edges = {(int(gids[a]), int(gids[b]), tuple(np.round(v, 6))) for a, b, v in zip(p['edge_index'][0], p['edge_index'][1], p['edge_vec'])}
        mask = np.isin(ei[0], gids[:nl])
        ref = {(int(a), int(b), tuple(np.round(v, 6))) for a, b, v in zip(ei[0][mask], ei[1][mask], ev[mask])}
        assert g_edges == ref


class _RowsEngine:
    """CPU stand-in for B200Engine.neighbor_rows (the device cell-list kernels): numpy neighbour list
    restricted to the requested centre atoms."""
    device = torch.device('cpu')

    def neighbor_rows(self, species, positions, cell, pbc, centres):
        ei, ev = build_graph(np.asarray(positions), np.asarray(cell), True, 5.0)
        centres = np.asarray(centres)
        pos_of = -np.ones(len(positions), dtype=np.int64)
        pos_of[centres] = np.arange(len(centres))
        keep = pos_of[ei[0]] >= 0
        c, s, v = pos_of[ei[0][keep]], ei[1][keep], ev[keep]
        o = np.argsort(c, kind='stable')
        rowptr = np.zeros(len(centres) + 1, dtype=np.int64)
        np.cumsum(np.bincount(c, minlength=len(centres)), out=rowptr[1:])
        return (torch.as_tensor(rowptr, dtype=torch.int32), torch.as_tensor(s[o], dtype=torch.int32),
                torch.as_tensor(v[o], dtype=torch.float32))


@pytest.mark.parametrize('grid,kind', [((2, 1, 1), 'si_long'), ((2, 2, 1), 'si'), ((1, 1, 2), 'nacl')])
def test_device_partition_equals_host_partition(grid, kind):
    """device_brick_partition (per-step, from positions, send lists derived locally) describes the same local
    systems as brick_decompose (global numpy neighbour list) -- same atoms in the same rows, same edges, and
    send lists that match what the peers expect in their ghost rows."""
    from sevenn_b200.parallel import device_brick_partition
    pos, cell, species = _system(kind)
    world = int(np.prod(grid))
    host = [brick_decompose(pos, cell, species, grid, r, 5.0) for r in range(world)]
    devp = [device_brick_partition(_RowsEngine(), pos, cell, species, grid, r) for r in range(world)]
    for r in range(world):
        h, d = host[r], devp[r]
        assert (h['n_local'], h['n_nodes'], h['n_interior']) == (d['n_local'], d['n_nodes'], d['n_interior'])
        assert np.array_equal(h['global_ids'], d['global_ids'])
        assert np.array_equal(h['species'], d['species'].numpy())
        assert np.array_equal(h['ghost_owner'], d['ghost_owner'])
        rp, src, vec = d['rowptr'].numpy(), d['src'].numpy(), d['edge_vec'].numpy()
        assert rp[-1] == h['edge_index'].shape[1]
        cen = np.repeat(np.arange(d['n_local']), np.diff(rp))

        def canon(c, s, v):
            o = np.lexsort((np.round(v[:, 2], 3), np.round(v[:, 1], 3), np.round(v[:, 0], 3), s, c))
            return c[o], s[o], v[o]
        hc, hs, hv = canon(h['edge_index'][0], h['edge_index'][1], np.asarray(h['edge_vec'], dtype=np.float64))
        dc, ds, dv = canon(cen, src, vec.astype(np.float64))
        assert np.array_equal(hc, dc) and np.array_equal(hs, ds) and np.allclose(hv, dv, atol=1e-5)
        # what rank r sends to q is exactly what q holds as ghosts owned by r, in q's ghost-row order
        for q in range(world):
            if q == r:
                assert d['send_lists'][q].numel() == 0
                continue
            sent_gids = d['global_ids'][d['send_lists'][q].numpy()]
            dq = devp[q]
            ghosts_q = dq['global_ids'][dq['n_local']:][dq['ghost_owner'] == r]
            assert np.array_equal(sent_gids, ghosts_q)
            assert dq['recv_counts'][r] == len(sent_gids)


class _PosEngine(FakeEngine, _RowsEngine):
    """stand-in with the two extra entry points the positions-in runner uses"""

    def set_graph_csr(self, species, rowptr, src, edge_vec, n_local):
        rp = torch.as_tensor(rowptr).long()
        centre = torch.repeat_interleave(torch.arange(int(n_local)), rp[1:] - rp[:-1])
        self.set_graph(species, torch.stack([centre, torch.as_tensor(src).long()]), edge_vec, n_local=int(n_local))


def _serial(pos, cell, species):
    ei, ev = build_graph(pos, cell, True, 5.0)
    ser = FakeEngine()
    ser.set_graph(species, ei, ev)
    for st, ts in [(STAGE_FWD_BEGIN, [0]), (STAGE_FWD_LAYER, range(ser.T)), (STAGE_FWD_END, [0])]:
        for t in ts:
            ser.run_stage(st, t)
    for t in range(ser.T - 1, -1, -1):
        ser.run_stage(STAGE_BWD_LAYER_A, t)
        if t > 0:
            ser.run_stage(STAGE_BWD_LAYER_B, t)
    ser.run_stage(STAGE_BWD_END)
    return float(ser.energy[0]), ser.forces.numpy().copy()


def _pos_worker(rank, world, port, grid, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pos, cell, species = _system('si_long')
        eng = _PosEngine()
        with DistributedRunner.from_positions(eng, pos, cell, species, grid) as run:
            run.compute()
            first = (run.part['global_ids'][:run.n_local].copy(), eng.forces[:run.n_local].numpy().copy(), float(eng.energy[0]))
            moved = pos + np.random.RandomState(3).normal(scale=0.3, size=pos.shape)     # far enough for atoms to change owner
            run.update_positions(moved)
            run.compute()
            second = (run.part['global_ids'][:run.n_local].copy(), eng.forces[:run.n_local].numpy().copy(), float(eng.energy[0]))
        q.put((rank, first, second))
    finally:
        dist.destroy_process_group()


def test_positions_in_runner_repartitions_every_step():
    """DistributedRunner.from_positions / update_positions (device_brick_partition + GhostExchange.from_lists: send
    lists derived locally, no handshake) under gloo with a stand-in engine: after the atoms move -- some change
    owner -- energy and forces still equal the serial evaluation of the moved system"""
    world, grid = 2, (2, 1, 1)
    pos, cell, species = _system('si_long')
    moved = pos + np.random.RandomState(3).normal(scale=0.3, size=pos.shape)
    refs = [_serial(pos, cell, species), _serial(moved, cell, species)]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pos_worker, args=(r, world, port, grid, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owners = []
    for k in range(2):
        forces = np.zeros((len(pos), 3))
        seen = np.zeros(len(pos), dtype=int)
        for rank, first, second in res:
            gids, f, energy = (first, second)[k]
            forces[gids] = f
            seen[gids] += 1
            assert abs(energy - refs[k][0]) < 1e-6 * abs(refs[k][0])          # edge vectors travel as fp32 here
        assert (seen == 1).all()
        assert np.allclose(forces, refs[k][1], atol=1e-5)
        owners.append({int(g): rank for rank, first, second in res for g in (first, second)[k][0]})
    assert any(owners[0][g] != owners[1][g] for g in owners[0])               # somebody really changed owner
