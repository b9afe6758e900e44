# A/B of several builds of the library within one box: tools/ab_libs.sh <libdir>... 
# (box-to-box variance is 5-8 %, so variants are only comparable inside one gpurun call)
for rep in 1 2; do
  for lib in "$@"; do
    S7B_LIB=$PWD/sevenn_b200/$lib/libsevenn_b200.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python -c "
import json,collections
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
g=collections.defaultdict(float)
for k,v in d['kernel_breakdown_ms'].items(): g[k.split('.')[0]]+=v
print('$lib', round(d['ms_per_step'],3), 'ms;', {k:round(v,3) for k,v in sorted(g.items(), key=lambda kv:-kv[1])[:2]}, {k.replace('conv_',''):round(v,3) for k,v in sorted(d['kernel_breakdown_ms'].items()) if k.startswith('conv') and ('.t1.' in k or '.t0.' in k)})
"
  done
done
