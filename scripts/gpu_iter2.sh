#!/bin/bash
# tests given in $TESTS, bench line (no extras), all-team phase profile
set -u
mkdir -p gpurun_out
tag=${1:-iter}
if [ -n "${TESTS:-}" ]; then timeout 900 python -m pytest $TESTS -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${tag}_tests.log; fi
timeout 600 python bench.py --no-extra --steps 3 --warmup 2 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/${tag}_bench.json
if [ -f opticomlib_b200/_ssfm_b200_prof.so ]; then
  SSFM_B200_LIB=$PWD/opticomlib_b200/_ssfm_b200_prof.so timeout 300 python scripts/exp_wf_prof.py 288 > gpurun_out/${tag}_prof.log 2>&1; echo "prof rc=$?"; grep -A40 "fp64 cluster 0" gpurun_out/${tag}_prof.log | grep -B40 "fp64 cluster 1" | head -50
fi
