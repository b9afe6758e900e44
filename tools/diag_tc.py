"""Diagnose wrong tensor-core linear outputs: for each failing configuration say WHICH tiles are wrong and
what the wrong values look like (old C missing / doubled, a prefix of the K chunks, another tile's data)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from sevenn_b200.engine import check, load_library, set_option
lib = load_library()


def run(n_nodes, a_K, c_N, acc, swz, seed=0):
    set_option('tc_swizzle', swz)
    rng = np.random.RandomState(seed)
    n_l = len(a_K)
    a_off, c_off, lda, ldc = [], [], 0, 0
    for l in range(n_l):
        a_off.append(lda); lda += (2 * l + 1) * a_K[l]
        c_off.append(ldc); ldc += (2 * l + 1) * c_N[l]
    A = rng.normal(size=(n_nodes, lda)).astype(np.float32)
    C0 = rng.normal(size=(n_nodes, ldc)).astype(np.float32)
    Ws = [(rng.normal(size=(a_K[l], c_N[l])) / np.sqrt(a_K[l])).astype(np.float32) for l in range(n_l)]
    W = np.ascontiguousarray(np.concatenate([w.ravel() for w in Ws]))
    a_t, c_t = torch.tensor(A, device='cuda'), torch.tensor(C0, device='cuda')
    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    ao, ak, co, cn = i32(a_off), i32(a_K), i32(c_off), i32(c_N)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.s7b_block_linear(a_t.data_ptr(), lda, n_nodes, n_l, ao.ctypes.data, ak.ctypes.data, W.ctypes.data,
                               c_t.data_ptr(), ldc, co.ctypes.data, cn.ctypes.data, int(acc), 1, st))
    torch.cuda.synchronize()
    got = c_t.cpu().numpy().astype(np.float64)
    print(f'== nodes {n_nodes} K {a_K} N {c_N} acc {acc} swz {swz}')
    for l in range(n_l):
        d, K, N = 2 * l + 1, a_K[l], c_N[l]
        a = A[:, a_off[l]:a_off[l] + d * K].reshape(n_nodes, d, K).astype(np.float64)
        g = got[:, c_off[l]:c_off[l] + d * N].reshape(n_nodes, d, N)
        c0 = C0[:, c_off[l]:c_off[l] + d * N].reshape(n_nodes, d, N).astype(np.float64)
        full = a @ Ws[l].astype(np.float64)
        ref = full + (c0 if acc else 0)
        err = np.abs(g - ref)
        bad = err > 1e-3
        print(f' block l={l}: bad fraction {bad.mean():.4f}, max err {err.max():.3e}')
        if not bad.any():
            continue
        NT = N if N <= 128 else next(N // c for c in range(2, 17) if N % c == 0 and (N // c) % 16 == 0 and N // c <= 128)
        for mt in range((n_nodes + 127) // 128):
            rows = slice(mt * 128, min(n_nodes, mt * 128 + 128))
            for ci in range(d):
                for nt in range(N // NT):
                    cols = slice(nt * NT, nt * NT + NT)
                    gb, rb = g[rows, ci, cols], ref[rows, ci, cols]
                    if np.abs(gb - rb).max() < 1e-3:
                        continue
                    hyp = {}
                    hyp['no_old'] = np.abs(gb - full[rows, ci, cols]).max()
                    hyp['old_twice'] = np.abs(gb - (full[rows, ci, cols] + 2 * c0[rows, ci, cols])).max()
                    hyp['only_old'] = np.abs(gb - c0[rows, ci, cols]).max()
                    for kk in range(32, K, 32):       # prefix / suffix of the K chunks
                        part = a[rows, ci, :kk] @ Ws[l][:kk, cols].astype(np.float64) + (c0[rows, ci, cols] if acc else 0)
                        hyp[f'prefix{kk}'] = np.abs(gb - part).max()
                        part = a[rows, ci, kk:] @ Ws[l][kk:, cols].astype(np.float64) + (c0[rows, ci, cols] if acc else 0)
                        hyp[f'suffix{kk}'] = np.abs(gb - part).max()
                    for ck in range(K // 32):          # one chunk missing / one chunk doubled
                        sl = slice(ck * 32, ck * 32 + 32)
                        one = a[rows, ci, sl] @ Ws[l][sl, cols].astype(np.float64)
                        hyp[f'missing{ck}'] = np.abs(gb - (rb - one)).max()
                        hyp[f'doubled{ck}'] = np.abs(gb - (rb + one)).max()
                    best = min(hyp, key=hyp.get)
                    rowbad = (np.abs(gb - rb) > 1e-3).any(axis=1)
                    colbad = (np.abs(gb - rb) > 1e-3).any(axis=0)
                    print(f'   tile mt={mt} ci={ci} nt={nt}: err {np.abs(gb - rb).max():.3e}; best hypothesis {best} ({hyp[best]:.2e}); '
                          f'bad rows {rowbad.sum()}/{rowbad.size} bad cols {colbad.sum()}/{colbad.size} first bad row {int(np.argmax(rowbad))}')


for cfg in [(777, [224], [112], False, 0), (1000, [224, 384, 352], [224, 64, 32], True, 1),
            (515, [256, 64, 32, 32], [256, 480, 416, 352], False, 1), (515, [256, 64, 32, 32], [256, 480, 416, 352], True, 1),
            (515, [128, 64, 32, 32], [256, 64, 32, 32], False, 1), (515, [256, 64, 32, 32], [128, 64, 32, 32], False, 1),
            (515, [128, 64, 32, 32], [128, 64, 32, 32], True, 1), (515, [256, 480, 416, 352], [256, 64, 32, 32], True, 1),
            (3, [256, 64, 32, 32], [256, 480, 416, 352], False, 1), (64, [256, 64, 32, 32], [256, 480, 416, 352], False, 1)]:
    run(*cfg)
