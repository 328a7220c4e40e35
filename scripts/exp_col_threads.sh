#!/bin/bash
# column-kernel CTA size: default library vs a -DSSFM_COL_THREADS=128 build (multi-launch schedule, long waveform)
for lib in "" "$PWD/opticomlib_b200/_var_t128.so"; do
  echo "== lib: ${lib:-default}"
  SSFM_B200_LIB=$lib timeout 300 python bench.py --steps 2 --warmup 1 --rows 1152 --schedule multilaunch --no-extra 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 multilaunch', d['value'], d['roofline']['kernel_ms'])"
  SSFM_B200_LIB=$lib timeout 300 python bench.py --workload cfg5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 1 GPU', d['value'], d['ms_per_step'])"
done
