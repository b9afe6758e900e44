"""Generate tests/golden/* from the reference tree (run in the build container only).

* ``reference_vectors.json``: the literal golden numbers the reference's own tests assert for
  this path (file:line cited per entry), plus the HfO2 inference CSVs parsed to arrays.
* ``w3j_reference.npz``: the Wigner-3j buffers carried by the shipped checkpoints
  (e3nn-generated), used to pin ``sevenn_b200/cg.py``.
* ``silu_norm.json``: e3nn's normalize2mom(silu) Monte-Carlo constant recomputed with this torch.
"""
import csv
import json
import os
import re

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden')

NACL = dict(cell=[[1.0, 2.815, 2.815], [2.815, 0.0, 2.815], [2.815, 2.815, 0.0]],
            positions=[[0.0, 0.0, 0.0], [2.815, 0.0, 0.0]], numbers=[11, 17], pbc=True)
H2O = dict(cell=None, positions=[[0.0, 0.2, 0.12], [0.0, 0.76, -0.48], [0.0, -0.76, -0.48]],
           numbers=[8, 1, 1], pbc=False)


def parse_extxyz(path):
    frames, lines, i = [], open(path).read().splitlines(), 0
    sym2z = {'Hf': 72, 'O': 8}
    while i < len(lines):
        n = int(lines[i])
        hdr = lines[i + 1]
        lat = [float(x) for x in re.search(r'Lattice="([^"]+)"', hdr).group(1).split()]
        rows = [l.split() for l in lines[i + 2:i + 2 + n]]
        frames.append(dict(cell=np.array(lat).reshape(3, 3).tolist(),
                           numbers=[sym2z[r[0]] for r in rows],
                           positions=[[float(v) for v in r[1:4]] for r in rows], pbc=True))
        i += 2 + n
    return frames


def main():
    os.makedirs(OUT, exist_ok=True)
    vec = {}
    # tests/unit_tests/test_pretrained.py:75-118 (7net-0 11Jul2024), atol 1e-6
    vec['7net0_nacl'] = dict(model='sevennet_0', system=NACL, source='test_pretrained.py:75-118',
                             energy=-3.779199,
                             forces=[[12.666697, 0.04726403, 0.04775861],
                                     [-12.666697, -0.04726403, -0.04775861]],
                             inferred_stress=[0.6439122, 0.03643947, 0.03643981,
                                              -0.04543639, -0.00599139, -0.04544507],
                             atol=dict(energy=1e-6, forces=1e-6, stress=1e-6))
    vec['7net0_h2o'] = dict(model='sevennet_0', system=H2O, source='test_pretrained.py:104-118',
                            energy=-12.782808303833008,
                            forces=[[0.0, -1.3619621e01, 7.5937047e00],
                                    [0.0, 9.3918495e00, -1.0172190e01],
                                    [0.0, 4.2277718e00, 2.5784855e00]],
                            atol=dict(energy=1e-6, forces=1e-6))
    # test_pretrained.py:121-164 (7net-l3i5), F/stress atol 1e-5
    vec['l3i5_nacl'] = dict(model='sevennet_l3i5', system=NACL, source='test_pretrained.py:121-164',
                            energy=-3.611131191253662,
                            forces=[[13.430887, 0.08655541, 0.08754013],
                                    [-13.430886, -0.08655544, -0.08754011]],
                            inferred_stress=[0.6818918, 0.04104544, 0.04107663,
                                             -0.04794561, -0.00565416, -0.04793138],
                            atol=dict(energy=1e-6, forces=1e-5, stress=1e-5))
    vec['l3i5_h2o'] = dict(model='sevennet_l3i5', system=H2O, source='test_pretrained.py:150-164',
                           energy=-12.700481414794922,
                           forces=[[0.0, -1.4547814e01, 8.1347866],
                                   [0.0, 1.0308369e01, -1.0880318e01],
                                   [0.0, 4.2394452, 2.7455316]],
                           atol=dict(energy=1e-6, forces=1e-6))
    # tests/unit_tests/test_calculator.py:56-84: rattle(stdev=0.01, seed=42) = pos + RandomState(42).normal(scale=0.01)
    rs = np.random.RandomState(42)
    nacl_r = dict(NACL, positions=(np.array(NACL['positions']) + rs.normal(scale=0.01, size=(2, 3))).tolist())
    vec['7net0_nacl_rattled'] = dict(
        model='sevennet_0', system=nacl_r, source='test_calculator.py:56-84',
        energy=-3.647711753845215, energies=[-1.7780534029006958, -1.8696582317352295],
        forces=[[13.095220565795898, 0.05549357831478119, 0.10542003065347672],
                [-13.095221519470215, -0.055493563413619995, -0.1054200679063797]],
        ase_stress=[-0.6614749431610107, -0.03719595819711685, -0.03681188449263573,
                    0.005672863684594631, 0.04221367835998535, 0.04504658654332161],
        atol=dict(energy=1e-5, forces=1e-5, stress=1e-5, energies=1e-5))
    rs = np.random.RandomState(42)
    h2o_r = dict(H2O, positions=(np.array(H2O['positions']) + rs.normal(scale=0.01, size=(3, 3))).tolist())
    vec['7net0_h2o_rattled'] = dict(
        model='sevennet_0', system=h2o_r, source='test_calculator.py:87-104',
        energy=-12.870156288146973,
        energies=[-6.2914958000183105, -3.1829171180725098, -3.3957436084747314],
        forces=[[-0.11430990695953369, -12.89616584777832, 6.915047645568848],
                [0.16116246581077576, 8.810967445373535, -9.560930252075195],
                [-0.04685257002711296, 4.085198402404785, 2.6458816528320312]],
        atol=dict(energy=1e-5, forces=1e-5, energies=1e-5))
    # test_calculator.py:240-312 disconnected / isolated oxygen atoms
    box = np.diag([20.0, 20.0, 20.0]).tolist()
    vec['7net0_single_o'] = dict(model='sevennet_0', source='test_calculator.py:241-247',
                                 system=dict(cell=box, positions=[[10, 10, 10]], numbers=[8], pbc=True),
                                 energy=-1.9413528442382812, energies=[-1.9413528442382812],
                                 forces=[[0.0, 0.0, 0.0]], ase_stress=[0.0] * 6,
                                 atol=dict(energy=1e-5, forces=1e-5, stress=1e-5, energies=1e-5))
    vec['7net0_two_o'] = dict(model='sevennet_0', source='test_calculator.py:248-254',
                              system=dict(cell=box, positions=[[5, 10, 10], [15, 10, 10]], numbers=[8, 8], pbc=True),
                              energy=-3.882704734802246, energies=[-1.941352367401123] * 2,
                              forces=[[0.0] * 3] * 2, ase_stress=[0.0] * 6,
                              atol=dict(energy=1e-5, forces=1e-5, stress=1e-5, energies=1e-5))
    vec['7net0_three_o'] = dict(model='sevennet_0', source='test_calculator.py:255-266',
                                system=dict(cell=box, positions=[[10, 10, 10], [12, 10, 10], [10, 10, 20]],
                                            numbers=[8, 8, 8], pbc=True),
                                energy=-6.802117824554443,
                                energies=[-2.43038272857666, -2.43038272857666, -1.941352367401123],
                                forces=[[3.8830623626708984, 0, 0], [-3.8830623626708984, 0, 0], [0, 0, 0]],
                                ase_stress=[0.0009707655990496278, 0, 0, 0, 0, 0],
                                atol=dict(energy=1e-5, forces=1e-5, stress=1e-5, energies=1e-5))
    # tests/data/inferences/snet0_on_hfo2 (7net-0 on tests/data/systems/hfo2.extxyz; test_cli.py:84-136)
    frames = parse_extxyz(f'{REF}/tests/data/systems/hfo2.extxyz')
    rows = list(csv.DictReader(open(f'{REF}/tests/data/inferences/snet0_on_hfo2/per_atom.csv')))
    graph = list(csv.DictReader(open(f'{REF}/tests/data/inferences/snet0_on_hfo2/per_graph.csv')))
    for f, fr in enumerate(frames):
        mine = [r for r in rows if int(r['stct_id']) == f]
        vec[f'7net0_hfo2_{f}'] = dict(
            model='sevennet_0', system=fr, source='tests/data/inferences/snet0_on_hfo2/per_atom.csv',
            energies=[float(r['atomic_energy']) for r in mine],
            forces=[[float(r['inferred_force_x']), float(r['inferred_force_y']), float(r['inferred_force_z'])] for r in mine],
            energy=float(graph[f]['inferred_total_energy']),
            stress_kbar=[float(graph[f][f'inferred_stress_{c}']) for c in ('xx', 'yy', 'zz', 'xy', 'yz', 'zx')],
            atol=dict(energy=2e-4, forces=2e-5, energies=2e-5, stress_kbar=2e-2))
    json.dump(vec, open(os.path.join(OUT, 'reference_vectors.json'), 'w'), indent=1)

    # Wigner-3j buffers from the checkpoints
    w3j = {}
    for name, path in [('7net0', f'{REF}/sevenn/pretrained_potentials/SevenNet_0__11Jul2024/checkpoint_sevennet_0.pth'),
                       ('l3i5', f'{REF}/sevenn/pretrained_potentials/SevenNet_l3i5/checkpoint_l3i5.pth')]:
        sd = torch.load(path, map_location='cpu', weights_only=False)['model_state_dict']
        for k, v in sd.items():
            if k.startswith('1_convolution') and '_w3j_' in k:
                w3j[f"{name}_{k.split('_w3j_')[1]}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, 'w3j_reference.npz'), **w3j)

    z = torch.randn(1_000_000, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    c = torch.nn.functional.silu(z).pow(2).mean().pow(-0.5).item()
    json.dump(dict(silu_norm=c), open(os.path.join(OUT, 'silu_norm.json'), 'w'))
    print('silu norm', c, '| vectors', len(vec), '| w3j', len(w3j))


if __name__ == '__main__':
    main()
