#!/bin/bash
# round-2 call 1: baseline of the current tree on one B200 (tests, bench with parity/extras, l3i5 ncu)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_smi.txt
free -g > gpurun_out/c1_free.txt; nproc >> gpurun_out/c1_free.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_pytest.log
timeout 900 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c1_bench.err
# CPU fp64 chunked oracle on the 12k cell: how long / how much memory (decides the default of --parity)
timeout 600 python - > gpurun_out/c1_cpu_oracle.txt 2>&1 <<'PY'
import os, sys, time, resource
import numpy as np, torch
sys.path.insert(0, '.')
from oracle.oracle import Oracle
from sevenn_b200.checkpoint import load_weights
from sevenn_b200.neighbors import build_graph, diamond_si
meta, arrays = load_weights('weights/sevennet_0.npz')
tm = {int(k): int(v) for k, v in meta['type_map'].items()}
pos, cell, z = diamond_si(10, 10, 15)
ei, ev = build_graph(pos, cell, True, 5.0)
sp = np.array([tm[int(a)] for a in z])
for th in (32, 64):
    torch.set_num_threads(th)
    o = Oracle(meta, arrays, dtype=torch.float64)
    t0 = time.perf_counter(); a = o.forward(sp, ei, ev, edge_chunk=32768)
    print('threads', th, 'cpu fp64 chunked 12k atoms', time.perf_counter() - t0, 's maxrss GB', resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, float(a['energy']), flush=True)
PY
cat gpurun_out/c1_cpu_oracle.txt | tail -3
# l3i5: launch list + full captures of the mid-layer conv kernels (lmax 3 kinds)
S7B_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c1_l3i5_launches.csv \
  python bench.py --model sevennet_l3i5 --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c1_l3i5_ncu_bench.log 2>&1
S7B_CUDA_GRAPH=0 S7B_CONCURRENT_CONV=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_kernel --launch-skip 18 --launch-count 4 -o gpurun_out/c1_l3i5_fwd \
  python bench.py --model sevennet_l3i5 --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c1_l3i5_ncu_fwd.log 2>&1
S7B_CUDA_GRAPH=0 S7B_CONCURRENT_CONV=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_bwd_kernel --launch-skip 21 --launch-count 4 -o gpurun_out/c1_l3i5_bwd \
  python bench.py --model sevennet_l3i5 --steps 1 --warmup 1 --no-cpu-baseline --parity off --no-extras > gpurun_out/c1_l3i5_ncu_bwd.log 2>&1
ls -la gpurun_out | tail -20
