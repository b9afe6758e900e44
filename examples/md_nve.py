"""Velocity-Verlet NVE molecular dynamics driven by the engine's positions-in entry point
(neighbour list, model and forces on the GPU each step).  Usage:
    python examples/md_nve.py [cells_per_side=3] [steps=500] [temperature_K=600]
Prints total-energy conservation and MD steps per second."""
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


def main():
    from sevenn_b200.calculator import resolve_model
    from sevenn_b200.engine import B200Engine
    from sevenn_b200.neighbors import diamond_si
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    temp = float(sys.argv[3]) if len(sys.argv) > 3 else 600.0
    meta, arrays = resolve_model('7net-0')
    eng = B200Engine(meta, arrays)
    pos, cell, z = diamond_si(n, n, n, sigma=0.0)
    species = np.full(len(pos), eng.spec.type_map[14], dtype=np.int32)
    t0 = time.perf_counter()
    hist = run_nve(eng, species, pos, cell, np.full(len(pos), 28.0855), steps, temperature=temp)
    dt = time.perf_counter() - t0
    tot = hist.sum(1)
    print(f'{len(pos)} atoms, {steps} steps of 1 fs: {steps / dt:.0f} steps/s; total energy range '
          f'{(tot.max() - tot.min()) / len(pos):.2e} eV/atom, kinetic range {(hist[:, 1].max() - hist[:, 1].min()) / len(pos):.2e} eV/atom; '
          f'CUDA graph (captures, replays) = {eng.graph_stats()}')


if __name__ == '__main__':
    main()
