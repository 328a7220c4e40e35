#!/bin/bash
# Round-end style run: smoke, all GPU tests, the bench line, the reference arm, ncu evidence for k_wf (cluster variant).
set -u
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --rows 288 --no-extra > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf -s 2 -c 2 -o gpurun_out/prof_wf_cluster_fp64 -f python scripts/prof_wf.py fp64 72 > gpurun_out/ncu_wf_cluster.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_wf_cluster.log
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print({k: d['roofline'][k] for k in ('kernel','achieved','frac','traffic','kernel_ms')}); print(d['cpu_baseline']); print(d['extra'])"
