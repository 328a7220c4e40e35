#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-mt}
timeout 900 python -m pytest tests/test_gpu_persistent.py -q -m gpu -k "multi_tile or multi_cluster" > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/${tag}_tests.log
timeout 600 python scripts/exp_mt.py ${2:-128} > gpurun_out/${tag}_exp.log 2>&1; echo "exp rc=$?"; cat gpurun_out/${tag}_exp.log
