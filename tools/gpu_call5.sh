#!/bin/bash
# round-2 call 5 (2 GPUs): multi-GPU parity tests (log kept under profiles/), bench N = 2 for both workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/c5_smi.txt
timeout -k 10 1200 python -m pytest tests/test_parallel_gpu.py -v > gpurun_out/c5_parallel_pytest.log 2>&1; echo "parallel pytest rc=$?" | tee -a gpurun_out/c5_parallel_pytest.log
tail -25 gpurun_out/c5_parallel_pytest.log | cut -c1-250
for g in 1 0; do
  S7B_CUDA_GRAPH=$g timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    > gpurun_out/c5_bench_n2_graph$g.json 2> gpurun_out/c5_bench_n2_graph$g.err; echo "bench n2 graph=$g rc=$?"
done
timeout 600 python bench.py --no-extras --no-cpu-baseline --parity off > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload nacl_d3 --steps 5 \
    > gpurun_out/c5_bench_d3_n2.json 2> gpurun_out/c5_bench_d3_n2.err; echo "bench d3 n2 rc=$?"
timeout 900 python bench.py --workload nacl_d3 --steps 5 > gpurun_out/c5_bench_d3_n1.json 2> gpurun_out/c5_bench_d3_n1.err; echo "bench d3 n1 rc=$?"
python - <<'PY'
import json
for f in ('c5_bench_n1', 'c5_bench_n2_graph1', 'c5_bench_n2_graph0', 'c5_bench_d3_n1', 'c5_bench_d3_n2'):
    try:
        d = json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'value', round(d['value']), 'graph', d['config'].get('cuda_graph'), d['config'].get('cuda_graph_note'), 'parity', d.get('parity'), {k: d[k] for k in ('ms_network', 'ms_d3') if k in d})
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -5 gpurun_out/c5_bench_n2_graph1.err gpurun_out/c5_bench_d3_n2.err | cut -c1-300
