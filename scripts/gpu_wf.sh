#!/bin/bash
# First GPU run of the persistent kernel: parity tests, then both schedules on a reduced batch, then the full bench.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
SSFM_DEBUG=1 timeout 900 python -m pytest tests/test_gpu_persistent.py -x -q 2>&1 | tail -25
for sched in persistent multilaunch; do
  for prec in fp64 fp32; do
    echo "== $sched $prec 1152 rows"
    timeout 300 python bench.py --steps 2 --warmup 1 --rows 1152 --precision $prec --schedule $sched --no-extra 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['step']['frac'], d['e2e']['value'])
except Exception as e: print('ERR', e)"
  done
done
echo "== full bench"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_wf.json 2> gpurun_out/bench_wf.err; echo "bench rc=$?"; cat gpurun_out/bench_wf.json; tail -3 gpurun_out/bench_wf.err
