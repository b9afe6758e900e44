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

import struct

import numpy as np

from .engine import default_table_knots, model_desc, prepare_params
from .spec import build_spec


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
