#!/bin/bash
# round-2 call 4: TC kernel (TMEM fix) + block-linear tests, per-layer TC/SIMT diff, D3 tests, full suite, A/B bench, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_tc_gemm_gpu.py -q > gpurun_out/c4_tc_pytest.log 2>&1; echo "tc pytest rc=$?" | tee -a gpurun_out/c4_tc_pytest.log
tail -15 gpurun_out/c4_tc_pytest.log | cut -c1-300
timeout -k 10 300 python tools/debug_tc.py 4 > gpurun_out/c4_debug_tc.txt 2>&1; cat gpurun_out/c4_debug_tc.txt | tail -30
timeout -k 10 900 python -m pytest tests/test_d3_gpu.py -q > gpurun_out/c4_d3_pytest.log 2>&1; echo "d3 pytest rc=$?" | tee -a gpurun_out/c4_d3_pytest.log
tail -30 gpurun_out/c4_d3_pytest.log | cut -c1-300
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c4_pytest.log
tail -12 gpurun_out/c4_pytest.log | cut -c1-300
for tc in 0 1; do
  S7B_TC_GEMM=$tc timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c4_bench_tc$tc.json 2> gpurun_out/c4_bench_tc$tc.err; echo "bench tc=$tc rc=$?"
done
python - <<'PY'
import json
for f in ('c4_bench_tc0', 'c4_bench_tc1'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        kb = d['kernel_breakdown_ms']
        gem = sum(v for k, v in kb.items() if 'gemm' in k)
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'gemm ms', round(gem, 3), 'e2e', round(d['e2e']['value']), 'parity', {k: d['parity'][k] for k in ('dE_eV', 'max_dF_eV_per_A', 'ok')})
        print('   ', {k: round(v, 3) for k, v in kb.items() if 'gemm' in k})
    except Exception as e:
        print(f, 'ERR', e)
PY
S7B_CUDA_GRAPH=0 S7B_CONCURRENT_CONV=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blocklin_tc_kernel --launch-skip 29 --launch-count 14 -o gpurun_out/c4_tc_gemm \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c4_ncu_tc.log 2>&1
S7B_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/c4_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c4_ncu_launches.log 2>&1
# D3 timing: rocksalt NaCl up to 25x25x10 = 50 000 atoms on one GPU (configs[4] system), default cutoffs
timeout 600 python - > gpurun_out/c4_d3_timing.txt 2>&1 <<'PY'
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from sevenn_b200.d3 import D3Engine
from sevenn_b200.neighbors import rocksalt_nacl
for cells in ((6, 6, 4), (12, 12, 10), (25, 25, 10)):
    pos, cell, z = rocksalt_nacl(*cells, sigma=0.05, seed=1)
    eng = D3Engine()
    eng.compute(z, pos, cell)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e, f, s = eng.compute(z, pos, cell)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.set_system(z, pos, cell)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for st in (1, 2, 3):
        eng.run_stage(st)
        ev[st].record()
    torch.cuda.synchronize()
    print(len(z), 'atoms: compute', round(dt * 1e3, 2), 'ms; stages (ms)', [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(3)], 'E', e, 'sumF', np.abs(f.sum(0)).max(), flush=True)
PY
cat gpurun_out/c4_d3_timing.txt | tail -5
ls -la gpurun_out | grep c4_
