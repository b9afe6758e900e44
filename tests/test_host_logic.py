"""CPU tests of the host logic: model description, layouts, neighbour lists, parameter
preparation and the C-ABI library's export table (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from helpers import ROOT, model_weights
from sevenn_b200.spec import build_spec, perm_cm_from_mulir


def test_sevennet0_layer_tables_match_survey_appendix_b():
    meta, _ = model_weights('sevennet_0')
    s = build_spec(meta)
    assert s.n_layers == 5 and s.lmax_filter == 2 and s.num_species == 89
    dims = [(L.dim_x, len(L.paths), L.weight_numel, L.dim_mid, L.dim_gate, L.dim_out) for L in s.layers]
    assert dims[0] == (128, 3, 384, 1152, 576, 480)
    assert dims[1] == dims[2] == dims[3] == (480, 15, 960, 3136, 576, 480)
    assert dims[4] == (480, 3, 224, 224, 128, 128)
    # slot, (l1 l2 l3), mul, w off, out off, old block
    table = [(0, 0, 0, 128, 0, 0, 0), (1, 1, 0, 64, 128, 128, 4), (2, 2, 0, 32, 192, 192, 12),
             (0, 1, 1, 128, 224, 224, 1), (1, 0, 1, 64, 352, 608, 3), (1, 1, 1, 64, 416, 800, 5),
             (1, 2, 1, 64, 480, 992, 7), (2, 1, 1, 32, 544, 1184, 10), (2, 2, 1, 32, 576, 1280, 13),
             (0, 2, 2, 128, 608, 1376, 2), (1, 1, 2, 64, 736, 2016, 6), (1, 2, 2, 64, 800, 2336, 8),
             (2, 0, 2, 32, 864, 2656, 9), (2, 1, 2, 32, 896, 2816, 11), (2, 2, 2, 32, 928, 2976, 14)]
    got = [(p.l1, p.l2, p.l3, p.mul, p.w_off, p.out_off, p.created) for p in s.layers[1].paths]
    assert got == table


def test_l3i5_layer_tables():
    meta, _ = model_weights('sevennet_l3i5')
    s = build_spec(meta)
    L = s.layers[2]
    assert (L.dim_x, len(L.paths), L.weight_numel, L.dim_mid, L.dim_gate, L.dim_out) == (704, 34, 1760, 7776, 832, 704)
    p33 = L.paths[33]
    assert (p33.l1, p33.l2, p33.l3, p33.w_off, p33.out_off, p33.created) == (3, 3, 3, 1728, 7552, 33)


def test_layout_permutations_are_bijections():
    meta, _ = model_weights('sevennet_0')
    L = build_spec(meta).layers[1]
    for p, n in [(perm_cm_from_mulir(list(L.x_muls)), L.dim_x), (L.mid_perm_cm_from_mulir(), L.dim_mid),
                 (perm_cm_from_mulir(list(L.gate_muls)), L.dim_gate)]:
        assert sorted(p.tolist()) == list(range(n))
    # component-major: first l=1 channel's three components are mul apart
    px = perm_cm_from_mulir([128, 64, 32])
    assert px[128] == 128 and px[129] == 131 and px[128 + 64] == 129


def test_neighbor_builders_agree():
    from sevenn_b200.neighbors import diamond_si, neighbor_list_brute, neighbor_list_cells
    pos, cell, _ = diamond_si(3, 3, 3, seed=1)
    ei_b, ev_b, _ = neighbor_list_brute(pos, cell, True, 5.0)
    ei_c, ev_c = neighbor_list_cells(pos, cell, 5.0)
    assert ei_b.shape == ei_c.shape == (2, 216 * 28)
    assert (ei_b == ei_c).all()
    assert np.allclose(ev_b, ev_c, atol=1e-9)
    assert (np.diff(ei_c[0]) >= 0).all()


def test_neighbor_small_cell_counts_images():
    from helpers import golden_vectors, system_graph
    ei, ev, vol = system_graph(golden_vectors()['7net0_nacl']['system'], 5.0)
    assert ei.shape[1] == 58 and abs(vol - 36.689) < 1e-2      # calculator test expects num_edges 58
    assert (np.linalg.norm(ev, axis=1) < 5.0).all()


def test_neighbor_counts_held_by_the_reference_tests():
    """tests/unit_tests/test_data.py:26-48,80-96 of the reference: directed edge counts at cutoff 4.0 of ASE's
    bulk('NaCl', 'rocksalt', a=5.63) primitive cell (36), molecule('H2O') (6), molecule('H') (0)"""
    from sevenn_b200.neighbors import neighbor_list_brute
    a = 5.63
    cell = np.array([[0, a / 2, a / 2], [a / 2, 0, a / 2], [a / 2, a / 2, 0]])
    ei, ev, _ = neighbor_list_brute(np.array([[0, 0, 0], [a / 2, a / 2, a / 2]]), cell, True, 4.0)
    assert ei.shape == (2, 36) and ev.shape == (36, 3)
    h2o = np.array([[0, 0, 0.119262], [0, 0.763239, -0.477047], [0, -0.763239, -0.477047]])
    ei, ev, _ = neighbor_list_brute(h2o, np.zeros((3, 3)), False, 4.0)
    assert ei.shape == (2, 6)
    ei, ev, _ = neighbor_list_brute(np.zeros((1, 3)), np.zeros((3, 3)), False, 4.0)
    assert ei.shape == (2, 0)


def test_neighbor_builders_against_the_definition():
    """both numpy builders against a direct enumeration of the definition (sevenn/train/dataload.py:32-129: every
    pair and every lattice translation with |r_j - r_i + S.cell| < cutoff, self-images at S != 0 included),
    written independently here: a sheared cell thinner than the cutoff, so several images of one pair count"""
    import itertools
    from sevenn_b200.neighbors import neighbor_list_brute, neighbor_list_cells
    rng = np.random.RandomState(5)
    cell = np.array([[4.1, 0.3, 0.0], [1.2, 4.6, 0.2], [0.4, -0.8, 9.5]])
    pos = rng.uniform(0, 1, size=(7, 3)) @ cell
    cutoff = 5.0
    want = []
    for i in range(len(pos)):
        for j in range(len(pos)):
            for S in itertools.product(range(-3, 4), repeat=3):
                if i == j and S == (0, 0, 0):
                    continue
                d = pos[j] - pos[i] + np.array(S, dtype=float) @ cell
                if np.linalg.norm(d) < cutoff:
                    want.append((i, j) + tuple(np.round(d, 9)))
    want.sort()
    for build in (lambda: neighbor_list_brute(pos, cell, True, cutoff)[:2], lambda: neighbor_list_cells(pos, cell, cutoff)):
        ei, ev = build()
        got = sorted((int(a), int(b)) + tuple(np.round(v, 9)) for a, b, v in zip(ei[0], ei[1], ev))
        assert len(got) == len(want) > 7 * 20
        assert np.allclose(np.array(got), np.array(want), atol=1e-8)


def test_prepare_params_shapes_and_table_accuracy():
    from sevenn_b200.engine import prepare_params, radial_weights
    meta, arrays = model_weights('sevennet_0')
    spec = build_spec(meta)
    P = prepare_params(spec, arrays, 'table', 500)
    assert P[('embed_x0', -1)].shape == (89, 128) and P[('embed_g0', -1)].shape == (89, 576)
    assert P[('si2', 1)].size == 86016 and P[('si1', 1)].size == 21504 and P[('sc', 4)].size == 128 * 128
    assert P[('table', 1)].shape == (500, 480, 4) and P[('table23', 1)].shape == (500, 480, 2)
    assert P[('readout', -1)].shape == (128,)
    # spline vs exact radial MLP (undo the channel-pair packing and the fp16 storage of a2, a3 first)
    r = np.random.RandomState(0).uniform(0.5, 4.999, 300)
    t01 = P[('table', 2)].astype(np.float64).reshape(500, 480, 2, 2)             # [k, pair, coef, parity]
    t23 = P[('table23', 2)].view(np.float16).astype(np.float64).reshape(500, 480, 2, 2)
    tab = np.concatenate([t01, t23], axis=2).transpose(0, 1, 3, 2).reshape(500, 960, 4)
    h = spec.cutoff / 500
    k = np.minimum((r / h).astype(int), 499)
    s = (r / h - k)[:, None]
    w = tab[k, :, 0] + s * (tab[k, :, 1] + s * (tab[k, :, 2] + s * tab[k, :, 3]))
    f, _ = radial_weights(spec, arrays, 2, r)
    assert np.abs(w - f).max() < 1e-4 * np.abs(f).max()
    Pm = prepare_params(spec, arrays, 'mlp', 0)
    assert Pm[('mlp2', 1)].shape == (64, 960) and Pm[('mlp2T', 1)].shape == (960, 64)


def test_library_exports_every_declared_symbol():
    """The built C-ABI library loads and exports exactly what include/sevenn_b200.h declares."""
    lib_path = os.path.join(ROOT, 'sevenn_b200', 'lib', 'libsevenn_b200.so')
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    header = open(os.path.join(ROOT, 'include', 'sevenn_b200.h')).read()
    declared = re.findall(r'S7B_API\s+[\w\s\*]+?\b((?:s7b|pair)_\w+)\s*\(', header)
    assert len(declared) >= 16
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), sym
    from sevenn_b200.engine import EXPORTS
    assert sorted(EXPORTS) == sorted(declared)
    lib.s7b_version.restype = ctypes.c_int
    assert lib.s7b_version() == 1


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from sevenn_b200.engine import B200Engine
    meta, arrays = model_weights('sevennet_0')
    with pytest.raises(RuntimeError, match='no CPU path'):
        B200Engine(meta, arrays)


def test_missing_library_fails_loudly(monkeypatch):
    import sevenn_b200.engine as eng
    monkeypatch.setattr(eng, '_lib', None)
    monkeypatch.setattr(eng, '_LIB_PATH', '/nonexistent/libsevenn_b200.so')
    with pytest.raises(ImportError, match='no CPU or PyTorch fallback'):
        eng.load_library()


def test_cpp_examples_compile(tmp_path):
    """examples/host_entry.cpp links against the built library; the LAMMPS pair style (no LAMMPS in the
    image) at least passes a syntax check against minimal stand-in declarations."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, 'sevenn_b200', 'lib')
    if not os.path.exists(os.path.join(lib_dir, 'libsevenn_b200.so')):
        import __graft_entry__
        __graft_entry__.build()
    subprocess.run(['g++', '-std=c++17', '-Wall', '-Werror', os.path.join(root, 'examples', 'host_entry.cpp'), '-o',
                    str(tmp_path / 'host_entry'), f'-L{lib_dir}', '-lsevenn_b200', f'-Wl,-rpath,{lib_dir}'], check=True)
    subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-Wall', '-Werror', '-I', os.path.join(root, 'tests', 'mock_lammps'),
                    os.path.join(root, 'examples', 'lammps', 'pair_e3gnn_b200.cpp')], check=True)
    subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-Wall', '-Werror', '-I', os.path.join(root, 'tests', 'mock_lammps'),
                    os.path.join(root, 'examples', 'lammps', 'pair_e3gnn_b200_parallel.cpp')], check=True)


def test_bench_graph_phase_guard_prints_the_direct_launch_line():
    """bench.py at N > 1 times the direct-launch schedule first; if the whole-step graph phase does not finish,
    a timer prints that phase's line and the process leaves with exit code 0 (never a run without a line)"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import argparse, time, bench\n"
        "a = argparse.Namespace(gpus=8, steps=20, warmup=5, model='sevennet_0')\n"
        "d = {'ms_per_step': 6.0, 'value': 100000 / 6.0e-3, 'gpu_launches': 1234}\n"
        "bench.graph_phase_guard(a, 0, 100000, 2800000, (25, 25, 20), d, {'sm_mhz': 1965.0, 'reasons': []}, deadline_s=0.3)\n"
        "time.sleep(30)\n"
        "print('not reached')\n")
    p = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1 and 'not reached' not in p.stdout
    line = json.loads(lines[0])
    assert line['metric'] == 'atom_updates_per_sec' and line['n_gpus'] == 8 and line['ms_per_step'] == 6.0
    assert line['config']['cuda_graph'] is False and 'did not finish' in line['config']['cuda_graph_note']
    for key in ('value', 'unit', 'steps', 'warmup', 'higher_is_better', 'scaling', 'dtype', 'data', 'clocks', 'gpu_launches'):
        assert key in line


def test_lammps_pair_styles_in_the_mock_harness(tmp_path):
    """examples/lammps/pair_e3gnn_b200{,_parallel}.cpp run inside tests/mock_lammps/harness_parallel.cpp: one rank, periodic
    image ghosts, a full neighbour list with skin, stock Comm forward/reverse through the pair style's own pack/unpack
    hooks, against a CPU double of the stage protocol (stub_s7b.cpp).  Graph with ghost rows + exchanges between the
    stages must equal the ghost-free evaluation: energy, per-atom energies, forces after the newton reverse sum, virial
    in LAMMPS component order."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mock, ex = os.path.join(root, 'tests', 'mock_lammps'), os.path.join(root, 'examples', 'lammps')
    exe = str(tmp_path / 'harness_parallel')
    subprocess.run(['g++', '-std=c++17', '-O1', '-Wall', '-Werror', '-I', mock, '-I', ex, os.path.join(mock, 'harness_parallel.cpp'),
                    os.path.join(ex, 'pair_e3gnn_b200_parallel.cpp'), os.path.join(ex, 'pair_e3gnn_b200.cpp'),
                    os.path.join(mock, 'stub_s7b.cpp'), '-o', exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip().endswith('OK'), p.stdout + p.stderr
