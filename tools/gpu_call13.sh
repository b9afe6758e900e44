#!/bin/bash
# round-2 call 13 (1 GPU): full GPU test suite (stage graphs, gate_bwd row maxima), default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 420 python -m pytest tests -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/c13_pytest.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/c13_bench.json').read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'], 3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity'))
    kb = d['kernel_breakdown_ms']
    import collections
    g = collections.defaultdict(float)
    for k, v in kb.items():
        g[k.split('.')[0]] += v
    print({k: round(v, 3) for k, v in sorted(g.items(), key=lambda x: -x[1])})
    print({k: d[k] for k in d if k.startswith('extra') or k in ('l3i5', 'gpu_standin')})
except Exception as e:
    print('ERR', e)
PY
tail -3 gpurun_out/c13_bench.err | cut -c1-300
