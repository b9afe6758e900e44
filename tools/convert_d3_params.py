"""Convert the DFT-D3 parameter tables the reference ships (Grimme's published reference C6 / R0 / r2r4 / rcov
values and the per-functional damping parameters) into ``weights/d3_params.npz``.

Run in the build container (needs /root/reference); the output is committed so that the GPU box -- which
has no /root/reference -- can load the parameters.  Same role as tools/convert_checkpoints.py: data
conversion, no reference code is copied.  Sources parsed:
  sevenn/pair_e3gnn/pair_d3_pars.h         R0AB_TABLE [94x94], C6AB_TABLE [32385x5]
  sevenn/pair_e3gnn/pair_d3_for_ase.cu     r2r4_ref / rcov_ref (:653-741), setfuncpar_zero / _bj (:394-556)
Checked here: the reference coordination numbers of the C6 table depend only on (element, reference index),
which is what makes C6_ij(CN_i, CN_j) separable (sevenn_b200/csrc/d3_kernels.cuh)."""
import json
import os
import re
import sys

import numpy as np

REF = '/root/reference/sevenn/pair_e3gnn'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'weights', 'd3_params.npz')


def numbers(text):
    return np.array([float(v) for v in re.findall(r'[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?', text)])


def main():
    pars = open(os.path.join(REF, 'pair_d3_pars.h')).read()
    i0, i1 = pars.index('#define R0AB_TABLE'), pars.index('#define C6AB_TABLE')
    r0ab = numbers(pars[i0 + len('#define R0AB_TABLE'):i1]).reshape(94, 94)
    c6tab = numbers(pars[i1 + len('#define C6AB_TABLE'):]).reshape(32385, 5)
    cu = open(os.path.join(REF, 'pair_d3_for_ase.cu')).read()

    def array_after(name):
        m = re.search(name + r'\[94\]\s*=\s*\{(.*?)\};', cu, re.S)
        body = re.sub(r'/\*.*?\*/', '', m.group(1), flags=re.S)
        v = numbers(body)
        assert v.shape == (94,), (name, v.shape)
        return v
    r2r4, rcov = array_after('r2r4_ref'), array_after('rcov_ref')

    # compact C6 reference table: c6ref[Zi, Zj, a, b], cnref[Z, a], mxc[Z]
    c6ref = np.zeros((94, 94, 5, 5))
    cnref = np.full((94, 5), np.nan)
    mxc = np.zeros(94, dtype=np.int32)
    for c6, z1, z2, cn1, cn2 in c6tab:
        z1, z2 = int(z1), int(z2)
        a, b = (z1 - 1) // 100, (z2 - 1) // 100
        e1, e2 = (z1 - 1) % 100, (z2 - 1) % 100
        for (ei, ai, ci) in ((e1, a, cn1), (e2, b, cn2)):
            assert np.isnan(cnref[ei, ai]) or cnref[ei, ai] == ci, 'reference CN is not a function of (element, index)'
            cnref[ei, ai] = ci
            mxc[ei] = max(mxc[ei], ai + 1)
        c6ref[e1, e2, a, b] = c6
        c6ref[e2, e1, b, a] = c6
    for zi in range(94):
        for zj in range(94):
            blk = c6ref[zi, zj] > 0
            want = np.zeros((5, 5), dtype=bool)
            want[:mxc[zi], :mxc[zj]] = True
            assert (blk == want).all(), ('valid block is not the mxc rectangle', zi, zj)
    cnref = np.nan_to_num(cnref, nan=0.0)

    def funcpars(fn):
        body = cu[cu.index(f'void PairD3::{fn}()'):]
        body = body[:body.index('\n}\n')]
        names = dict((n, int(c)) for n, c in re.findall(r'\{\s*"([^"]+)"\s*,\s*(\d+)\s*\}', body))
        defaults = dict((k, float(v)) for k, v in re.findall(r'^\s*(s6|alp|rs18)\s*=\s*([-\d.]+);', body[:body.index('commandMap')], re.M))
        cases = {}
        for code, rest in re.findall(r'case\s+(\d+):(.*?)break;', body, re.S):
            cases[int(code)] = dict((k, float(v)) for k, v in re.findall(r'(rs6|s18|rs18|s6|alp)\s*=\s*([-\d.eE+]+);', rest))
        out = {}
        for n, c in names.items():
            if c in cases:
                p = dict(defaults)
                p.update(cases[c])
                out[n] = p
        return out
    func = {'damp_zero': funcpars('setfuncpar_zero'), 'damp_bj': funcpars('setfuncpar_bj')}
    assert abs(func['damp_bj']['pbe']['rs6'] - 0.4289) < 1e-12 and abs(func['damp_zero']['pbe']['rs6'] - 1.217) < 1e-12
    np.savez_compressed(OUT, r0ab=r0ab, c6ref=c6ref.astype(np.float64), cnref=cnref, mxc=mxc, r2r4=r2r4, rcov=rcov,
                        functionals=np.frombuffer(json.dumps(func).encode(), dtype=np.uint8))
    # flat binary for the reference-named C entry points (pair_run_coeff in sevenn_b200/csrc/d3.cu)
    binp = OUT.replace('.npz', '.bin')
    with open(binp, 'wb') as f:
        for arr in (r0ab, c6ref, cnref, mxc.astype(np.float64), r2r4, rcov):
            f.write(np.ascontiguousarray(arr, dtype=np.float64).tobytes())
        for damp, table in func.items():
            for name, p in table.items():
                f.write(f"{damp} {name} {p['s6']!r} {p['rs6']!r} {p['s18']!r} {p['rs18']!r} {p['alp']!r}\n".encode())
    print('wrote', binp, os.path.getsize(binp), 'bytes')
    print('wrote', OUT, os.path.getsize(OUT), 'bytes;', len(func['damp_bj']), 'bj /', len(func['damp_zero']), 'zero functionals')


if __name__ == '__main__':
    sys.exit(main())
