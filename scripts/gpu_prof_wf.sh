#!/bin/bash
# ncu evidence for the persistent kernel: launch list of a bench run + one full capture per precision
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_wf.csv python bench.py --steps 1 --warmup 1 --rows 288 --no-extra > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
for prec in fp64 fp32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf -s 1 -c 1 -o gpurun_out/prof_wf_$prec -f python scripts/prof_wf.py $prec 72 > gpurun_out/ncu_wf_$prec.log 2>&1; echo "ncu $prec rc=$?"; tail -2 gpurun_out/ncu_wf_$prec.log
done
timeout 300 python bench.py --steps 2 --warmup 1 --rows 1152 --precision fp32 --no-extra 2>&1 | tail -1 | cut -c1-400
python scripts/exp_cfg2b.py
ls -la gpurun_out
