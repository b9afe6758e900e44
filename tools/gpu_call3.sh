#!/bin/bash
# round-2 call 3: the TMA/tcgen05 node-linear kernel -- unit tests, full GPU suite, A/B bench, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_tc_gemm_gpu.py -x -q > gpurun_out/c3_tc_pytest.log 2>&1; echo "tc pytest rc=$?" | tee -a gpurun_out/c3_tc_pytest.log
tail -25 gpurun_out/c3_tc_pytest.log
if ! grep -q "rc=0" gpurun_out/c3_tc_pytest.log; then
  # diagnose: tiny problem, both tile modes, values printed
  timeout -k 10 200 python - > gpurun_out/c3_tc_diag.txt 2>&1 <<'PY'
import ctypes, numpy as np, torch, sys
sys.path.insert(0, '.')
from sevenn_b200.engine import check, load_library, set_option
lib = load_library()
for swz in (1, 0):
    set_option('tc_swizzle', swz)
    for rows, K, N in ((128, 32, 32), (200, 64, 16), (300, 224, 112)):
        rng = np.random.RandomState(0)
        A = rng.normal(size=(rows, K)).astype(np.float32); W = (rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)
        a, w = torch.tensor(A, device='cuda'), torch.tensor(W, device='cuda')
        c = torch.full((rows, N), float('nan'), device='cuda')
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = lib.s7b_dense_linear(a.data_ptr(), w.data_ptr(), c.data_ptr(), rows, K, N, 1, st)
        torch.cuda.synchronize()
        ref = A.astype(np.float64) @ W.astype(np.float64)
        got = c.cpu().numpy()
        print('swz', swz, rows, K, N, 'rc', rc, 'maxerr', np.nanmax(np.abs(got - ref)), 'nan', np.isnan(got).sum(), 'ratio00', got[0, 0] / ref[0, 0], flush=True)
        print(' got', got[0, :4], got[1, :4], '\n ref', ref[0, :4], ref[1, :4], flush=True)
PY
  cat gpurun_out/c3_tc_diag.txt | tail -40
fi
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c3_pytest.log
tail -5 gpurun_out/c3_pytest.log
for tc in 0 1; do
  S7B_TC_GEMM=$tc timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c3_bench_tc$tc.json 2> gpurun_out/c3_bench_tc$tc.err; echo "bench tc=$tc rc=$?"
done
S7B_TC_SWIZZLE=0 timeout 600 python bench.py --no-extras --no-cpu-baseline --parity off > gpurun_out/c3_bench_tc1_noswz.json 2> gpurun_out/c3_bench_tc1_noswz.err
timeout 600 python bench.py --radial mlp --no-extras --no-cpu-baseline > gpurun_out/c3_bench_mlp.json 2> gpurun_out/c3_bench_mlp.err
S7B_TC_GEMM=0 timeout 600 python bench.py --radial mlp --no-extras --no-cpu-baseline > gpurun_out/c3_bench_mlp_tc0.json 2> gpurun_out/c3_bench_mlp_tc0.err
python - <<'PY'
import json
for f in ('c3_bench_tc0', 'c3_bench_tc1', 'c3_bench_tc1_noswz', 'c3_bench_mlp', 'c3_bench_mlp_tc0'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        kb = d['kernel_breakdown_ms']
        gem = sum(v for k, v in kb.items() if 'gemm' in k)
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'gemm ms', round(gem, 3), 'parity', d.get('parity'))
    except Exception as e:
        print(f, 'ERR', e)
PY
S7B_CUDA_GRAPH=0 S7B_CONCURRENT_CONV=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:blocklin_tc_kernel --launch-skip 29 --launch-count 14 -o gpurun_out/c3_tc_gemm \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c3_ncu_tc.log 2>&1
S7B_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/c3_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c3_ncu_launches.log 2>&1
ls -la gpurun_out | grep c3_
