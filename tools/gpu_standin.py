"""Reference-GPU stand-in: the oracle restatement of the reference torch/e3nn
path executed with stock torch CUDA ops (unfused gather -> einsum -> index_add_, torch autograd
backward) in fp32 on the benchmark cell.  The real reference GPU path needs e3nn, which is not
installable here; this measures the same unfused computation.  Test/benchmark infrastructure."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from sevenn_b200.checkpoint import load_weights
from sevenn_b200.neighbors import build_graph, diamond_si
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else 'sevennet_0'
cells = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (10, 10, 15)
as_json = '--json' in sys.argv
meta, arrays = load_weights(os.path.join(ROOT, 'weights', f'{name}.npz'))
tm = {int(k): int(v) for k, v in meta['type_map'].items()}
pos, cell, z = diamond_si(*cells)
ei, ev = build_graph(pos, cell, True, 5.0)
sp = np.array([tm[int(a)] for a in z])
o = Oracle(meta, arrays, dtype=torch.float32, device='cuda')
for _ in range(2):
    out = o.forward(sp, ei, ev)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    out = o.forward(sp, ei, ev)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
if as_json:
    import json
    print(json.dumps({'what': 'torch-CUDA fp32 unfused stand-in of the reference GPU path (oracle restatement executed with stock '
                              'torch CUDA ops + autograd; the e3nn path itself cannot be installed here)',
                      'model': name, 'atoms': len(z), 'edges': int(ei.shape[1]), 'ms_per_step': 1e3 * min(ts),
                      'value': len(z) / min(ts), 'unit': 'atom-updates/s', 'best_of': len(ts),
                      'peak_mem_GiB': torch.cuda.max_memory_allocated() / 2**30, 'energy_eV': float(out['energy'])}))
    sys.exit(0)
print(f'{name} {len(z)} atoms {ei.shape[1]} edges: torch-CUDA fp32 unfused stand-in {1e3*min(ts):.1f} ms/step '
      f'(best of 5) = {len(z)/min(ts):.0f} atom-updates/s; E = {float(out["energy"]):.3f} eV; '
      f'peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB')
