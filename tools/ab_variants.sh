for v in "" "_nv1"; do
  S7B_LIB=$PWD/sevenn_b200/lib/libsevenn_b200$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab$v.json 2>gpurun_out/bench_ab$v.err
  python -c "
import json,collections
d=json.loads(open('gpurun_out/bench_ab$v.json').read().strip().splitlines()[-1])
bd=d['kernel_breakdown_ms']; g=collections.defaultdict(float)
for k,v in bd.items(): g[k.split('.')[0]]+=v
print('variant [$v]', round(d['ms_per_step'],3), 'ms;', {k:round(v,3) for k,v in sorted(g.items(), key=lambda kv:-kv[1])[:6]})
print('   ', {k:round(v,4) for k,v in bd.items() if k.endswith('.t2.l0') or k.endswith('.t2.l1') or k.endswith('.t2.l2') or k=='si2_gemm.t2'})
"
done
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
