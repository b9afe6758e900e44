#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 300 python tools/tc_trace.py > gpurun_out/c10_trace.txt 2>&1; cat gpurun_out/c10_trace.txt | cut -c1-400
