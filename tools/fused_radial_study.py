"""Design study for round 2 (DESIGN.md section 6, item ii): evaluate the radial MLP's last layer
w = h2(r) @ W2 on tensor cores inside the convolution kernel, with h2(r) (64 values) from a small
cubic-Hermite table instead of the 960-wide w table.  CPU emulation of the arithmetic only:
how accurate is w (and dw/dr) for (a) the h2 table alone, (b) the product in fp32, single TF32,
3xTF32 (hi/lo split, three products) and bf16x3 splits?  Reference: fp64 MLP.
Usage: python tools/fused_radial_study.py [model] [layer]"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sevenn_b200.checkpoint import load_weights  # noqa: E402
from sevenn_b200.engine import _dsilu, _silu, default_table_knots, radial_embedding, radial_weights  # noqa: E402
from sevenn_b200.spec import SILU_NORM, build_spec  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hidden(spec, arrays, t, r):
    """h2(r) [n, 64] and dh2/dr after the two hidden layers (float64)."""
    h, dh = radial_embedding(spec, arrays['bessel_coeffs'], r)
    for j in range(len(spec.radial_hidden)):
        W = arrays[f'{t}.mlp{j}'].astype(np.float64) / math.sqrt(arrays[f'{t}.mlp{j}'].shape[0])
        z, dz = h @ W, dh @ W
        h, dh = SILU_NORM * _silu(z), SILU_NORM * _dsilu(z) * dz
    return h, dh


def round_mantissa(x, bits):
    """round-to-nearest-even of float32 values to `bits` explicit mantissa bits (tf32: 10, bf16: 7)."""
    x = np.asarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    u = (u + (1 << (drop - 1)) - 1 + ((u >> drop) & 1)) >> drop << drop
    return u.astype(np.uint32).view(np.float32)


def split(x, bits, terms):
    out, rem = [], np.asarray(x, dtype=np.float32)
    for _ in range(terms):
        hi = round_mantissa(rem, bits)
        out.append(hi)
        rem = (rem - hi).astype(np.float32)
    return out


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else 'sevennet_0'
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    meta, arrays = load_weights(os.path.join(ROOT, 'weights', f'{model}.npz'))
    spec = build_spec(meta)
    knots = default_table_knots(spec)
    hstep = spec.cutoff / knots
    rk = np.arange(knots + 1) * hstep
    H, dH = hidden(spec, arrays, t, rk)
    f0, f1, d0, d1 = H[:-1], H[1:], dH[:-1] * hstep, dH[1:] * hstep
    tab = np.stack([f0, d0, 3 * (f1 - f0) - (2 * d0 + d1), 2 * (f0 - f1) + d0 + d1], -1).astype(np.float32).astype(np.float64)
    rng = np.random.RandomState(0)
    r = rng.uniform(1.5, spec.cutoff - 1e-6, size=20000)
    k = np.minimum((r / hstep).astype(int), knots - 1)
    s = (r / hstep - k)[:, None]
    c = tab[k]
    h_tab = c[..., 0] + s * (c[..., 1] + s * (c[..., 2] + s * c[..., 3]))
    dh_tab = (c[..., 1] + s * (2 * c[..., 2] + 3 * s * c[..., 3])) / hstep
    W2 = arrays[f'{t}.mlp2'].astype(np.float64) / math.sqrt(arrays[f'{t}.mlp2'].shape[0])
    w_ref, dw_ref = radial_weights(spec, arrays, t, r)
    print(f'{model} layer {t}: W = {W2.shape[1]}, max|w| = {np.abs(w_ref).max():.2f}, max|dw/dr| = {np.abs(dw_ref).max():.2f}, '
          f'h2 table = {tab.nbytes / 2 / 2**20:.2f} MiB fp32 (w table today: {knots * W2.shape[1] * 12 / 2**20:.1f} MiB)')

    def report(name, w, dw=None):
        e = np.abs(w - w_ref)
        line = f'  {name:44s} max|dw| {e.max():.2e}  rms {np.sqrt((e ** 2).mean()):.2e}'
        if dw is not None:
            line += f'   d/dr: max {np.abs(dw - dw_ref).max():.2e}'
        print(line)

    report('h2 table (fp32 coeffs), product in fp64', h_tab @ W2, dh_tab @ W2)
    h32, W32 = h_tab.astype(np.float32), W2.astype(np.float32)
    acc = np.zeros((len(r), W2.shape[1]), np.float32)
    for kk in range(h32.shape[1]):          # fp32 FMA chain (what the SIMT kernels do)
        acc = (acc.astype(np.float64) + h32[:, kk:kk + 1].astype(np.float64) * W32[kk].astype(np.float64)).astype(np.float32)
    report('fp32 inputs, fp32 sequential accumulation', acc.astype(np.float64))
    for name, bits, terms, keep in [('single TF32', 10, 1, 1), ('3xTF32 (hi*hi + hi*lo + lo*hi)', 10, 2, 3),
                                    ('bf16x3 (6 products up to 2nd order)', 7, 3, 6), ('bf16x2 (3 products)', 7, 2, 3)]:
        hs, ws = split(h32, bits, terms), split(W32, bits, terms)
        prods = sorted(((i + j, i, j) for i in range(terms) for j in range(terms)))[:keep]
        tot = np.zeros_like(acc, dtype=np.float64)
        for _, i, j in prods:               # each product exact in fp32; ideal (fp64) accumulation = best case
            tot += hs[i].astype(np.float64) @ ws[j].astype(np.float64)
        report(f'{name}, exact accumulation', tot)


if __name__ == '__main__':
    main()
