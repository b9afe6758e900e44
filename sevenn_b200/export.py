"""Flat binary export of a prepared model for non-Python hosts (LAMMPS-style C++ callers).

File layout (little endian), consumed by ``examples/host_entry.cpp``:
    magic    8 bytes  b'S7BMODEL'
    version  int32    1
    desc     sizeof(S7bModelDesc) bytes, exactly the C struct of include/sevenn_b200.h
    n_arrays int32
    n_types  int32, then n_types x (atomic number int32, species index int32)
    then per array: name char[32] (NUL padded), layer int32, numel int64, numel x float32
The arrays are the ones ``sevenn_b200.engine.prepare_params`` produces (normalisations folded in,
radial tables packed), i.e. exactly what ``s7b_engine_set_param`` takes.
"""
from __future__ import annotations

import ctypes
import struct

import numpy as np

from .engine import S7bModelDesc, default_table_knots, prepare_params
from .spec import build_spec


def model_desc(spec, knots: int) -> S7bModelDesc:
    d = S7bModelDesc()
    d.n_layers, d.lmax_filter, d.num_species, d.n_basis = spec.n_layers, spec.lmax_filter, spec.num_species, spec.n_basis
    d.cutoff, d.cutoff_fn = spec.cutoff, 0 if spec.cutoff_fn == 'XPLOR' else 1
    d.cutoff_on, d.poly_p = spec.cutoff_on, spec.poly_p
    d.radial_hidden[0], d.radial_hidden[1] = spec.radial_hidden
    irreps = [list(L.x_muls) for L in spec.layers] + [list(spec.layers[-1].out_muls)]
    for t, muls in enumerate(irreps):
        d.n_l[t] = len(muls)
        for l, m in enumerate(muls):
            d.muls[t][l] = m
    d.table_knots = knots
    return d


def export_flat(path: str, meta: dict, arrays, radial: str = 'table', knots=None) -> None:
    spec = build_spec(meta)
    knots = (knots or default_table_knots(spec)) if radial == 'table' else 0
    params = prepare_params(spec, arrays, radial, knots)
    d = model_desc(spec, knots)
    with open(path, 'wb') as f:
        f.write(b'S7BMODEL')
        f.write(struct.pack('<i', 1))
        f.write(bytes(d))
        f.write(struct.pack('<i', len(params)))
        f.write(struct.pack('<i', len(spec.type_map)))
        for z, idx in sorted(spec.type_map.items()):
            f.write(struct.pack('<ii', int(z), int(idx)))
        for (name, layer), arr in params.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            f.write(name.encode().ljust(32, b'\0'))
            f.write(struct.pack('<iq', int(layer), int(a.size)))
            f.write(a.tobytes())
