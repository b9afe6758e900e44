# A/B within one box: prints ms/step and the GEMM kernel breakdown
timeout 200 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in "S7B_TC_GEMM=0" "S7B_TC_GEMM=1"; do
  env $v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2>gpurun_out/bench_ab.err
  python -c "
import json,collections
d=json.loads(open('gpurun_out/bench_ab.json').read().strip().splitlines()[-1])
bd=d['kernel_breakdown_ms']; g=collections.defaultdict(float)
for k,v in bd.items(): g[k.split('.')[0]]+=v
print('$v', round(d['ms_per_step'],3), 'ms;', {k:round(v,3) for k,v in sorted(g.items(), key=lambda kv:-kv[1]) if 'gemm' in k})
"
done
