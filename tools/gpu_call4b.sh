#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -k 10 600 python tools/diag_tc.py > gpurun_out/c4b_diag.txt 2>&1; tail -80 gpurun_out/c4b_diag.txt | cut -c1-250
