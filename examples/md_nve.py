"""Velocity-Verlet NVE molecular dynamics driven by the engine's positions-in entry point
(neighbour list, model and forces on the GPU each step).  Usage:
    python examples/md_nve.py [cells_per_side=3] [steps=500] [temperature_K=600]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 examples/md_nve.py 8 200     # N GPUs
Under torchrun every rank integrates the same (replicated) positions and evaluates its own brick: partition,
ghost lists and graph are rebuilt on the device every step (sevenn_b200.parallel.device_brick_partition), ghost
features travel over NCCL, forces are all-gathered.  Prints total-energy conservation and MD steps per second."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

AMU_A2_FS2_IN_EV = 103.642696      # 1 amu A^2 / fs^2 in eV
KB = 8.617333e-5                   # eV / K


def run_nve(engine, species, positions, cell, masses, steps, dt_fs=1.0, temperature=600.0, seed=0):
    """-> array [steps, 2] of (potential, kinetic) energy in eV; positions are advanced in place."""
    rng = np.random.RandomState(seed)
    m = np.asarray(masses, dtype=np.float64)[:, None] * AMU_A2_FS2_IN_EV
    v = rng.normal(size=positions.shape) * np.sqrt(KB * temperature / m)
    v -= (v * m).sum(0) / m.sum()
    e, _, f, _, _ = engine.compute_positions(species, positions, cell, True)
    hist = np.zeros((steps, 2))
    for s in range(steps):
        v += 0.5 * dt_fs * f / m
        positions += dt_fs * v
        e, _, f, _, _ = engine.compute_positions(species, positions, cell, True)
        v += 0.5 * dt_fs * f / m
        hist[s] = e, 0.5 * (m * v * v).sum()
    return hist


class DistributedForces:
    """compute_positions() of a multi-GPU runner: (energy, None, forces of ALL atoms, None, None)"""

    def __init__(self, engine, species, pos, cell, grid):
        import torch
        import torch.distributed as dist
        from sevenn_b200.parallel import DistributedRunner
        self.torch, self.dist = torch, dist
        self.runner = DistributedRunner.from_positions(engine, pos, cell, species, grid)
        self.n = len(species)

    def compute_positions(self, species, positions, cell, pbc):
        torch, dist, run = self.torch, self.dist, self.runner
        run.update_positions(positions, cell)
        run.compute()
        r = run.results()
        f_all = torch.zeros(self.n, 3, dtype=torch.float32, device=run.device)
        f_all[torch.as_tensor(np.asarray(r['global_ids']), dtype=torch.long, device=run.device)] = r['forces']
        dist.all_reduce(f_all)
        return float(r['energy'].cpu()[0]), None, f_all.cpu().numpy().astype(np.float64), None, None


def main():
    from sevenn_b200.calculator import resolve_model
    from sevenn_b200.engine import B200Engine
    from sevenn_b200.neighbors import diamond_si
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    temp = float(sys.argv[3]) if len(sys.argv) > 3 else 600.0
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    meta, arrays = resolve_model('7net-0')
    pos, cell, z = diamond_si(n, n, n, sigma=0.0)
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        eng = B200Engine(meta, arrays, device=local)
        grid = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
        species = np.full(len(pos), eng.spec.type_map[14], dtype=np.int32)
        driver = DistributedForces(eng, species, pos, cell, grid)
    else:
        eng = B200Engine(meta, arrays)
        species = np.full(len(pos), eng.spec.type_map[14], dtype=np.int32)
        driver = eng
    t0 = time.perf_counter()
    hist = run_nve(driver, species, pos, cell, np.full(len(pos), 28.0855), steps, temperature=temp)
    dt = time.perf_counter() - t0
    tot = hist.sum(1)
    if rank == 0:
        print(f'{len(pos)} atoms on {world} GPU(s), {steps} steps of 1 fs: {steps / dt:.0f} steps/s; total energy range '
              f'{(tot.max() - tot.min()) / len(pos):.2e} eV/atom, kinetic range {(hist[:, 1].max() - hist[:, 1].min()) / len(pos):.2e} eV/atom; '
              f'CUDA graph (captures, replays) = {eng.graph_stats()}', flush=True)
    if world > 1:
        driver.runner.close()           # a captured step graph would pin the NCCL communicator
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
