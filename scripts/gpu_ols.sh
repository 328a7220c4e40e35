#!/bin/bash
set -u
mkdir -p gpurun_out
tag=${1:-ols}
timeout 900 python -m pytest tests/test_gpu_filters.py tests/test_gpu_pd_edfa.py tests/test_gpu_cfg4_receiver.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/${tag}_tests.log
timeout 600 python scripts/exp_filt2.py > gpurun_out/${tag}_exp.log 2>&1; echo "exp rc=$?"; cat gpurun_out/${tag}_exp.log
SSFM_FILTFILT_NO_OLS=1 timeout 600 python scripts/exp_filt2.py > gpurun_out/${tag}_exp_noots.log 2>&1; echo "== without overlap-save"; cat gpurun_out/${tag}_exp_noots.log
