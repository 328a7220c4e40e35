#!/bin/bash
# Run on the GPU box through gpurun: bench line, chunk sweep, ncu launch list and full captures.
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for c in 16 32 64 128 512; do
  echo "chunk $c"; timeout 300 python bench.py --steps 2 --warmup 1 --rows 1024 --chunk $c --no-extra 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
echo "fp32"; timeout 300 python bench.py --steps 2 --warmup 1 --rows 1024 --precision fp32 --no-extra 2>&1 | tail -1 | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --rows 256 --no-extra > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_row -s 6 -c 2 -o gpurun_out/prof_krow -f python bench.py --steps 1 --warmup 1 --rows 256 --no-extra > gpurun_out/ncu_krow.log 2>&1; echo "ncu krow rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_col -s 12 -c 2 -o gpurun_out/prof_kcol -f python bench.py --steps 1 --warmup 1 --rows 256 --no-extra > gpurun_out/ncu_kcol.log 2>&1; echo "ncu kcol rc=$?"
ls -la gpurun_out
