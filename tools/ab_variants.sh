# A/B within one box: prints ms/step and the grouped kernel breakdown
for v in "S7B_X=1" "S7B_LIB=$PWD/sevenn_b200/lib/libsevenn_b200_old.so"; do
  env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ab.json 2>gpurun_out/bench_ab.err
  python -c "
import json,collections
d=json.loads(open('gpurun_out/bench_ab.json').read().strip().splitlines()[-1])
bd=d['kernel_breakdown_ms']; g=collections.defaultdict(float)
for k,v in bd.items(): g[k.split('.')[0]]+=v
print('$v'[:40], round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), {k:round(v,3) for k,v in sorted(g.items(), key=lambda kv:-kv[1])[:8]})
"
done
timeout 250 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
