"""Shared helpers for tests (systems from tests/golden, graph building, oracle cache)."""
import functools
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


@functools.lru_cache(maxsize=None)
def golden_vectors():
    return json.load(open(os.path.join(GOLDEN, 'reference_vectors.json')))


@functools.lru_cache(maxsize=None)
def model_weights(name):
    from sevenn_b200.checkpoint import load_weights
    return load_weights(os.path.join(ROOT, 'weights', f'{name}.npz'))


@functools.lru_cache(maxsize=None)
def oracle(name, dtype_name='float64'):
    import torch
    from oracle.oracle import Oracle
    meta, arrays = model_weights(name)
    return Oracle(meta, arrays, dtype=getattr(torch, dtype_name))


def system_graph(system, cutoff):
    from sevenn_b200.neighbors import build_graph
    pos = np.asarray(system['positions'], dtype=np.float64)
    cell = np.zeros((3, 3)) if system['cell'] is None else np.asarray(system['cell'], dtype=np.float64)
    ei, ev = build_graph(pos, cell, bool(system['pbc']), cutoff)
    vol = abs(np.linalg.det(cell)) if system['pbc'] else 0.0
    return ei, ev, vol


def species_of(meta, numbers):
    tm = {int(k): int(v) for k, v in meta['type_map'].items()}
    return np.array([tm[int(z)] for z in numbers], dtype=np.int64)
