"""The C ABI used from C++ without Python or torch (the LAMMPS-style binding of INTEGRATION.md):
examples/host_entry.cpp is compiled with g++, linked against libsevenn_b200.so and run on the GPU;
its energies / forces must equal the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import ROOT, golden_vectors, model_weights, oracle, species_of, system_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def host(tmp_path_factory):
    from sevenn_b200.export import export_flat
    d = tmp_path_factory.mktemp('cpp_host')
    exe = str(d / 'host_entry')
    lib_dir = os.path.join(ROOT, 'sevenn_b200', 'lib')
    subprocess.check_call(['g++', '-O1', '-std=c++17', os.path.join(ROOT, 'examples', 'host_entry.cpp'), '-o', exe,
                           f'-L{lib_dir}', '-lsevenn_b200', f'-Wl,-rpath,{lib_dir}'])
    meta, arrays = model_weights('sevennet_0')
    model = str(d / 'sevennet_0.s7b')
    export_flat(model, meta, arrays)
    return exe, model, d


def _parse(out, n):
    lines = out.strip().splitlines()
    return float(lines[0]), np.array([[float(v) for v in l.split()] for l in lines[1:1 + n]]), np.array([float(v) for v in lines[1 + n].split()])


def test_cpp_host_graph_entry(host):
    exe, model, d = host
    g = golden_vectors()['7net0_hfo2_1']
    meta, _ = model_weights('sevennet_0')
    ei, ev, vol = system_graph(g['system'], 5.0)
    z = np.array(g['system']['numbers'], dtype=np.int32)
    path = str(d / 'graph.bin')
    with open(path, 'wb') as f:
        f.write(struct.pack('<iq', len(z), ei.shape[1]))
        f.write(z.tobytes() + ei[0].astype(np.int32).tobytes() + ei[1].astype(np.int32).tobytes() + ev.astype(np.float32).tobytes())
    out = subprocess.run([exe, model, path], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    energy, forces, virial = _parse(out.stdout, len(z))
    ref = oracle('sevennet_0').forward(species_of(meta, z), ei, ev, volume=vol)
    assert abs(energy - float(ref['energy'])) < 2e-5
    assert np.allclose(forces, ref['forces'].numpy(), atol=5e-5)
    assert np.allclose(forces, g['forces'], atol=max(g['atol']['forces'], 5e-5))   # the reference's own CSV
    assert np.allclose(virial, ref['virial'].numpy(), atol=5e-4)


def test_cpp_host_positions_entry_and_errors(host):
    exe, model, d = host
    g = golden_vectors()['7net0_nacl_rattled']
    sysd = g['system']
    z = np.array(sysd['numbers'], dtype=np.int32)
    path = str(d / 'pos.bin')
    with open(path, 'wb') as f:
        f.write(struct.pack('<i3i', len(z), 1, 1, 1))
        f.write(np.array(sysd['cell'], dtype=np.float64).tobytes() + z.tobytes() + np.array(sysd['positions'], dtype=np.float64).tobytes())
    out = subprocess.run([exe, model, path, 'pos'], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    energy, forces, virial = _parse(out.stdout, len(z))
    ei, _, _ = system_graph(sysd, 5.0)
    assert f'edges: {ei.shape[1]}' in out.stderr
    assert abs(energy - g['energy']) < 5e-5 and np.allclose(forces, g['forces'], atol=5e-5)
    # unsorted edges must be rejected by the library, with a message
    bad = str(d / 'bad.bin')
    with open(bad, 'wb') as f:
        f.write(struct.pack('<iq', 2, 2))
        f.write(z[:2].tobytes() + np.array([1, 0], np.int32).tobytes() + np.array([0, 1], np.int32).tobytes() + np.ones((2, 3), np.float32).tobytes())
    out = subprocess.run([exe, model, bad], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and 'sorted by centre' in out.stderr
