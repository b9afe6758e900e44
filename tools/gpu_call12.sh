#!/bin/bash
# round-2 call 12: TC kernel with the vectorised epilogue: correctness, timeline, A/B bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_tc_gemm_gpu.py -q -x > gpurun_out/c12_tc_pytest.log 2>&1; echo "tc pytest rc=$?" | tee -a gpurun_out/c12_tc_pytest.log
tail -6 gpurun_out/c12_tc_pytest.log | cut -c1-300
timeout -k 10 300 python tools/tc_trace.py > gpurun_out/c12_trace.txt 2>&1; grep -E "^==|prodA|raw_landed|ops_ready|acc_f|tile_done|per chunk" gpurun_out/c12_trace.txt | cut -c1-200
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c12_pytest.log
tail -5 gpurun_out/c12_pytest.log | cut -c1-300
for tc in 0 1; do
  S7B_TC_GEMM=$tc timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c12_bench_tc$tc.json 2> gpurun_out/c12_bench_tc$tc.err; echo "bench tc=$tc rc=$?"
done
python - <<'PY'
import json
for f in ('c12_bench_tc0', 'c12_bench_tc1'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        kb = d['kernel_breakdown_ms']
        gem = sum(v for k, v in kb.items() if 'gemm' in k)
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'gemm ms', round(gem, 3), 'e2e', round(d['e2e']['value']), 'parity', {k: d['parity'][k] for k in ('dE_eV', 'max_dF_eV_per_A', 'ok')})
        print('   ', {k: round(v, 3) for k, v in kb.items() if 'gemm' in k})
    except Exception as e:
        print(f, 'ERR', e)
PY
S7B_CUDA_GRAPH=0 S7B_CONCURRENT_CONV=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blocklin_tc_kernel|row_exponent" --launch-skip 42 --launch-count 10 -o gpurun_out/c12_tc_gemm \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c12_ncu_tc.log 2>&1
S7B_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c12_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c12_ncu_launches.log 2>&1
ls -la gpurun_out | grep c12_
