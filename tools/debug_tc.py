"""Localise a tensor-core / SIMT GEMM disagreement inside the engine: run the stage sequence twice on the
same graph (tc_gemm = 0 and 1) and print, per layer, the largest relative difference of the GEMM outputs."""
import os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from sevenn_b200.checkpoint import load_weights
from sevenn_b200.engine import B200Engine, set_option, STAGE_FWD_BEGIN, STAGE_FWD_LAYER, STAGE_FWD_END, STAGE_BWD_LAYER_A, STAGE_BWD_LAYER_B, STAGE_BWD_END
from sevenn_b200.neighbors import build_graph, diamond_si
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = sys.argv[2] if len(sys.argv) > 2 else 'sevennet_0'
meta, arrays = load_weights(os.path.join(ROOT, 'weights', f'{model}.npz'))
tm = {int(k): int(v) for k, v in meta['type_map'].items()}
pos, cell, z = diamond_si(nc, nc, nc)
ei, ev = build_graph(pos, cell, True, 5.0)
sp = np.array([tm[int(a)] for a in z])
set_option('cuda_graph', 0)
snaps = {}
for tc in (0, 1):
    set_option('tc_gemm', tc)
    e = B200Engine(meta, arrays)
    e.set_graph(sp, ei, ev)
    T = e.spec.n_layers
    s = {}
    e.run_stage(STAGE_FWD_BEGIN)
    for t in range(T):
        e.run_stage(STAGE_FWD_LAYER, t)
        torch.cuda.synchronize()
        L = e.spec.layers[t]
        s[f'fwd{t}.gate_in'] = e.buffer('gate_in', t, shape=(e.n_local, L.dim_gate)).clone()
        if t + 1 < T:
            s[f'fwd{t}.x_next'] = e.buffer('x', t + 1, shape=(e.n_nodes, e.spec.layers[t + 1].dim_x)).clone()
    e.run_stage(STAGE_FWD_END)
    for t in range(T - 1, -1, -1):
        e.run_stage(STAGE_BWD_LAYER_A, t)
        torch.cuda.synchronize()
        s[f'bwd{t}.dx'] = e.buffer('dx', t, shape=(e.n_nodes, e.spec.layers[t].dim_x)).clone()
        if t > 0:
            e.run_stage(STAGE_BWD_LAYER_B, t)
            torch.cuda.synchronize()
            s[f'bwd{t}.dh'] = e.buffer('dh', t, shape=(e.n_local, e.spec.layers[t].dim_x)).clone()
    e.run_stage(STAGE_BWD_END)
    torch.cuda.synchronize()
    s['energy'] = e.buffer('energy', dtype='f8').clone()
    s['forces'] = e.buffer('forces', shape=(e.n_nodes, 3)).clone()
    snaps[tc] = s
for k in snaps[0]:
    a, b = snaps[0][k].double(), snaps[1][k].double()
    d = (a - b).abs()
    print(f'{k:16s} max|simt| {float(a.abs().max()):.4e}  max|diff| {float(d.max()):.3e}  rel {float(d.max() / (a.abs().max() + 1e-30)):.3e}', flush=True)
