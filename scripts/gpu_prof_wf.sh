#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/prof_wf.py fp64 36
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf -s 1 -c 1 -o gpurun_out/prof_wf_fp64 -f python scripts/prof_wf.py fp64 36 > gpurun_out/ncu_wf.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_wf.log
ls -la gpurun_out
