#!/bin/bash
# one build -> measure iteration on a B200: k_wf parity tests, team-variant timings, phase profile (clock64 build)
set -u
mkdir -p gpurun_out
tag=${1:-iter}
timeout 900 python -m pytest tests/test_gpu_persistent.py tests/test_gpu_fiber.py ${EXTRA_TESTS:-} -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python scripts/exp_wf.py 288 ${2:-fp64,fp32} > gpurun_out/${tag}_exp_wf.log 2>&1; echo "exp rc=$?"; cat gpurun_out/${tag}_exp_wf.log
if [ -f opticomlib_b200/_ssfm_b200_prof.so ]; then
  SSFM_B200_LIB=$PWD/opticomlib_b200/_ssfm_b200_prof.so timeout 300 python scripts/exp_wf_prof.py 72 > gpurun_out/${tag}_prof.log 2>&1; echo "prof rc=$?"; cat gpurun_out/${tag}_prof.log
fi
