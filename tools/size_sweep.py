"""Step time vs system size on one GPU (device-resident graph, CUDA events), from the launch-bound
regime (64 atoms) to the largest cell that fits.  Usage: python tools/size_sweep.py [model] [max_atoms]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sevenn_b200.checkpoint import load_weights  # noqa: E402
from sevenn_b200.engine import B200Engine, set_option  # noqa: E402
from sevenn_b200.neighbors import diamond_si  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else 'sevennet_0'
    max_atoms = int(sys.argv[2]) if len(sys.argv) > 2 else 1_100_000
    meta, arrays = load_weights(os.path.join(ROOT, 'weights', f'{model}.npz'))
    set_option('cuda_graph', int(os.environ.get('S7B_CUDA_GRAPH', '1')))
    eng = B200Engine(meta, arrays)
    si = eng.spec.type_map[14]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    rows = []
    for reps in [(2, 2, 2), (4, 4, 4), (6, 6, 6), (8, 8, 8), (10, 10, 15), (16, 16, 16), (25, 25, 20), (32, 32, 32), (40, 40, 40), (50, 50, 50)]:
        n = 8 * reps[0] * reps[1] * reps[2]
        if n > max_atoms:
            break
        pos, cell, _ = diamond_si(*reps)
        species = np.full(n, si, dtype=np.int32)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); eng.set_positions(species, pos, cell, True); t1.record(); torch.cuda.synchronize()
        nl_ms = t0.elapsed_time(t1)
        for _ in range(3):
            eng.compute()
        steps = 20 if n < 100_000 else 5
        ms = []
        for _ in range(steps):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); eng.compute(); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        med = float(np.median(ms))
        row = dict(atoms=n, edges=eng.n_edges, ms_per_step=round(med, 4), atom_updates_per_s=round(n / med * 1e3),
                   nl_ms_incl_h2d=round(nl_ms, 3), mem_gb=round((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30, 2),
                   energy_per_atom=float(eng.buffer('energy', dtype='f8')[0]) / n)
        print(json.dumps(row), flush=True)
        rows.append(row)
    return rows


if __name__ == '__main__':
    main()
