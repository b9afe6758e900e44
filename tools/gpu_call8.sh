#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tc in 1 0; do S7B_TC_GEMM=$tc timeout -k 10 300 python tools/debug_tc2.py 3 sevennet_l3i5 > gpurun_out/c8_debug2_tc$tc.txt 2>&1; echo "tc=$tc"; tail -12 gpurun_out/c8_debug2_tc$tc.txt | cut -c1-220; done
