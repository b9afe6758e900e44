"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the brick-decomposed, NCCL-exchanged
evaluation must reproduce the single-GPU engine and the CPU oracle on the same system."""
import os
import socket

import numpy as np
import pytest

from helpers import model_weights, oracle, species_of

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, grid, model, q):
    import torch
    import torch.distributed as dist
    from sevenn_b200.engine import B200Engine
    from sevenn_b200.neighbors import diamond_si
    from sevenn_b200.parallel import DistributedRunner, brick_decompose
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    run = rp = None
    try:
        meta, arrays = model_weights(model)
        pos, cell, z = diamond_si(4, 3, 3, seed=2)
        part = brick_decompose(pos, cell, species_of(meta, z), grid, rank, 5.0)
        run = DistributedRunner(B200Engine(meta, arrays, device=rank), part)
        run.set_cuda_graph(False)
        run.compute()                       # direct launches: split convolutions, overlapped exchanges
        torch.cuda.synchronize()
        r_eager = run.results()
        run.set_cuda_graph(True)
        for _ in range(3):                  # capture, then replays of the whole step incl. the NCCL calls
            run.compute()
        torch.cuda.synchronize()
        r = run.results()
        assert run.graph_error is None, run.graph_error
        assert run.graph_captures == 1 and run.graph_replays == 3
        assert abs(float(r['energy'].cpu()[0]) - float(r_eager['energy'].cpu()[0])) < 1e-9
        assert torch.allclose(r['forces'], r_eager['forces'], atol=2e-6)
        h = run.compute_host()
        # positions in: partition, ghost lists and graph built on the device, send lists derived locally
        rp = DistributedRunner.from_positions(B200Engine(meta, arrays, device=rank), pos, cell, species_of(meta, z), grid)
        rp.compute()
        torch.cuda.synchronize()
        r2 = rp.results()
        assert abs(float(r2['energy'].cpu()[0]) - float(r['energy'].cpu()[0])) < 2e-5
        assert np.array_equal(r2['global_ids'], r['global_ids'])
        assert torch.allclose(r2['forces'], r['forces'], atol=5e-6)
        moved = pos + np.random.RandomState(9).normal(scale=0.03, size=pos.shape)
        rp.update_positions(moved)
        rp.compute()
        torch.cuda.synchronize()
        r3 = rp.results()
        q.put((rank, r['global_ids'], r['forces'].cpu().numpy(), float(r['energy'].cpu()[0]),
               r['atomic_energy'].cpu().numpy(), r['virial'].cpu().numpy(), h['energy'], h['forces'].copy(),
               r3['global_ids'], r3['forces'].cpu().numpy(), float(r3['energy'].cpu()[0])))
    finally:
        for rr in (run, rp):                # the captured graph pins the communicator: release it before the teardown
            if rr is not None:
                rr.close()
        dist.destroy_process_group()


@pytest.mark.parametrize('world,grid,model', [(2, (2, 1, 1), 'sevennet_0'), (2, (1, 2, 1), 'sevennet_l3i5'),
                                              (4, (2, 2, 1), 'sevennet_0'), (8, (2, 2, 2), 'sevennet_0')])
def test_multi_gpu_matches_oracle(world, grid, model):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    from sevenn_b200.neighbors import build_graph, diamond_si
    meta, _ = model_weights(model)
    pos, cell, z = diamond_si(4, 3, 3, seed=2)
    ei, ev = build_graph(pos, cell, True, 5.0)
    ref = oracle(model).forward(species_of(meta, z), ei, ev, volume=abs(np.linalg.det(cell)))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, model, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:                     # never leave a rank behind (a hung teardown would hang pytest's exit)
            if p.is_alive():
                p.kill()
    forces = np.zeros((len(pos), 3))
    forces_h = np.zeros((len(pos), 3))
    forces_m = np.zeros((len(pos), 3))
    ae = np.zeros(len(pos))
    moved = pos + np.random.RandomState(9).normal(scale=0.03, size=pos.shape)
    ei_m, ev_m = build_graph(moved, cell, True, 5.0)
    ref_m = oracle(model).forward(species_of(meta, z), ei_m, ev_m)
    for rank, gids, f, energy, a, virial, e_h, f_h, gids_m, f_m, e_m in res:
        forces[gids], forces_h[gids], ae[gids] = f, f_h, a
        forces_m[gids_m] = f_m
        assert abs(e_m - float(ref_m['energy'])) < 1e-4      # after update_positions: re-partitioned on the device
        assert abs(energy - float(ref['energy'])) < 1e-4
        assert abs(e_h - energy) < 1e-6
        # virial = sum over ~4000 edges of r (x) f in fp32 products: fp32 edge-force noise (~5e-6 eV/A)
        # accumulates to ~2e-3 eV on components of magnitude ~90 eV
        assert np.allclose(virial, ref['virial'].numpy(), atol=5e-3, rtol=1e-5)
    assert np.allclose(forces, ref['forces'].numpy(), atol=5e-5)
    assert np.allclose(forces_h, forces, atol=5e-6)
    assert np.allclose(forces_m, ref_m['forces'].numpy(), atol=5e-5)
    assert np.allclose(ae, ref['atomic_energy'].numpy(), atol=2e-5)


def _d3_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from sevenn_b200.d3 import D3Engine, distributed_d3
    from sevenn_b200.neighbors import rocksalt_nacl
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        pos, cell, z = rocksalt_nacl(3, 3, 2, sigma=0.05, seed=2)
        e, f, s = distributed_d3(D3Engine(device=rank), z, pos, cell)
        q.put((rank, e, f, s))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_multi_gpu_d3_matches_single_gpu(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    from sevenn_b200.d3 import D3Engine
    from sevenn_b200.neighbors import rocksalt_nacl
    pos, cell, z = rocksalt_nacl(3, 3, 2, sigma=0.05, seed=2)
    e1, f1, s1 = D3Engine(device=0).compute(z, pos, cell)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_d3_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:                     # never leave a rank behind (a hung teardown would hang pytest's exit)
            if p.is_alive():
                p.kill()
    for rank, e, f, s in res:
        assert abs(e - e1) < 1e-9 * abs(e1)
        assert np.abs(f - f1).max() < 1e-10
        assert np.abs(s - s1).max() < 1e-9 * np.abs(s1).max()
