"""dh after B1 (= dg x sc^T) per irrep block against fp64, tensor-core and SIMT, every backward layer."""
import os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from sevenn_b200.checkpoint import load_weights
from sevenn_b200 import engine as E
from sevenn_b200.neighbors import build_graph, diamond_si
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model = sys.argv[2] if len(sys.argv) > 2 else 'sevennet_l3i5'
meta, arrays = load_weights(os.path.join(ROOT, 'weights', f'{model}.npz'))
tm = {int(k): int(v) for k, v in meta['type_map'].items()}
pos, cell, z = diamond_si(nc, nc, nc)
ei, ev = build_graph(pos, cell, True, 5.0)
sp = np.array([tm[int(a)] for a in z])
E.set_option('cuda_graph', 0)
for tc in (1, 0):
    E.set_option('tc_gemm', tc)
    e = E.B200Engine(meta, arrays)
    e.set_graph(sp, ei, ev)
    spec = e.spec
    params = E.prepare_params(spec, arrays, 'table', e.knots)
    T = spec.n_layers
    e.run_stage(E.STAGE_FWD_BEGIN)
    for t in range(T):
        e.run_stage(E.STAGE_FWD_LAYER, t)
    e.run_stage(E.STAGE_FWD_END)
    for t in range(T - 1, 0, -1):
        L = spec.layers[t]
        e.run_stage(E.STAGE_BWD_LAYER_A, t)
        torch.cuda.synchronize()
        dg = e.buffer('dg', t, shape=(e.n_local, L.dim_gate)).clone()
        e.run_stage(E.STAGE_BWD_LAYER_B1, t)
        torch.cuda.synchronize()
        dh1 = e.buffer('dh', t, shape=(e.n_local, L.dim_x)).clone()
        e.run_stage(E.STAGE_BWD_LAYER_B2, t)
        n_sc = min(len(L.x_muls), len(L.gate_muls))
        flat = params[('scT', t)]
        goff = xoff = woff = 0
        for l in range(n_sc):
            d, K, N = 2 * l + 1, L.gate_muls[l], L.x_muls[l]
            W = torch.tensor(flat[woff:woff + K * N].reshape(K, N), dtype=torch.float64, device='cuda')
            a = dg[:, goff:goff + d * K].reshape(-1, d, K).double()
            ref = (a @ W)
            got = dh1[:, xoff:xoff + d * N].reshape(-1, d, N).double()
            err = (got - ref).abs()
            rows_bad = (err.amax((1, 2)) > 1e-5 * ref.abs().max()).nonzero().flatten()
            print(f'tc={tc} bwd{t} scT l={l} K={K} N={N}: max|ref| {float(ref.abs().max()):.3e} max err {float(err.max()):.3e}; bad nodes {rows_bad.numel()} '
                  f'{rows_bad[:6].tolist()}; |dg| row max range [{float(a.abs().amax(2).min()):.2e}, {float(a.abs().amax(2).max()):.2e}]', flush=True)
            goff += d * K; xoff += d * N; woff += K * N
    del e
