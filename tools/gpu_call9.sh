#!/bin/bash
# round-2 call 9: TC kernel v3 (8 transform warps, packed math), l3i5 localisation, A/B bench, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 600 python tools/diag_tc.py > gpurun_out/c9_diag.txt 2>&1; grep -c "bad fraction 0.0000" gpurun_out/c9_diag.txt; grep -B1 -A6 "bad fraction" gpurun_out/c9_diag.txt | grep -v "0.0000" | head -40 | cut -c1-250
timeout -k 10 600 python -m pytest tests/test_tc_gemm_gpu.py -q > gpurun_out/c9_tc_pytest.log 2>&1; echo "tc pytest rc=$?" | tee -a gpurun_out/c9_tc_pytest.log
tail -8 gpurun_out/c9_tc_pytest.log | cut -c1-300
timeout -k 10 300 python tools/debug_tc3.py 3 sevennet_l3i5 > gpurun_out/c9_debug3_l3i5.txt 2>&1; cat gpurun_out/c9_debug3_l3i5.txt | tail -40 | cut -c1-260
timeout -k 10 1500 python -m pytest tests -m gpu -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c9_pytest.log
tail -12 gpurun_out/c9_pytest.log | cut -c1-300
for tc in 0 1; do
  S7B_TC_GEMM=$tc timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c9_bench_tc$tc.json 2> gpurun_out/c9_bench_tc$tc.err; echo "bench tc=$tc rc=$?"
done
python - <<'PY'
import json
for f in ('c9_bench_tc0', 'c9_bench_tc1'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        kb = d['kernel_breakdown_ms']
        gem = sum(v for k, v in kb.items() if 'gemm' in k)
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'gemm ms', round(gem, 3), 'e2e', round(d['e2e']['value']), 'parity', {k: d['parity'][k] for k in ('dE_eV', 'max_dF_eV_per_A', 'ok')})
        print('   ', {k: round(v, 3) for k, v in kb.items() if 'gemm' in k})
    except Exception as e:
        print(f, 'ERR', e)
PY
S7B_CUDA_GRAPH=0 S7B_CONCURRENT_CONV=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blocklin_tc_kernel|row_exponent" --launch-skip 58 --launch-count 12 -o gpurun_out/c9_tc_gemm \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c9_ncu_tc.log 2>&1
ls -la gpurun_out | grep c9_
