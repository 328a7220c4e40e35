#!/bin/bash
# round-2 evidence: launch list of the bench command, DRAM traffic of k_wf at FULL size, full captures, sanitizer logs
set -u
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/r2_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:k_wf -c 4 --csv --log-file gpurun_out/r2_traffic_k_wf_4096rows.csv python bench.py --steps 1 --warmup 1 --no-extra > gpurun_out/r2_ncu_traffic.log 2>&1; echo "traffic fp64 rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:k_wf -c 4 --csv --log-file gpurun_out/r2_traffic_k_wf_fp32_4096rows.csv python bench.py --steps 1 --warmup 1 --no-extra --precision fp32 > gpurun_out/r2_ncu_traffic32.log 2>&1; echo "traffic fp32 rc=$?"
for prec in fp64 fp32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf -s 2 -c 2 -o /tmp/r2_prof_wf_$prec -f python scripts/prof_wf.py $prec 72 > gpurun_out/r2_ncu_wf_$prec.log 2>&1; echo "ncu k_wf $prec rc=$?"
  python scripts/ncu_summarize.py /tmp/r2_prof_wf_$prec.ncu-rep gpurun_out/r2_ncu_summary_k_wf_$prec.txt "# ncu --set full --clock-control none --import-source on -k regex:k_wf -s 2 -c 2, scripts/prof_wf.py $prec 72: the two launches (cluster teams, fill teams) of ONE propagation of 72 config-#3 waveforms"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_col|k_row|k_pd|k_unpack|k_filtfilt|k_scatter|k_save|k_welch|k_edfa" -s 14 -c 40 -o /tmp/r2_prof_filters -f python scripts/prof_filters.py 64 > gpurun_out/r2_ncu_filters.log 2>&1; echo "ncu filters rc=$?"
python scripts/ncu_summarize.py /tmp/r2_prof_filters.ncu-rep gpurun_out/r2_ncu_summary_filters.txt "# ncu --set full, scripts/prof_filters.py 64 (64 frames x 2^18, fp64): second pass of BPF, PD -> LPF -> SAMPLER, Welch (launches 15..54 of the process)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_col|k_row" -s 8 -c 8 -o /tmp/r2_prof_long -f python scripts/prof_long.py > gpurun_out/r2_ncu_long.log 2>&1; echo "ncu long rc=$?"
python scripts/ncu_summarize.py /tmp/r2_prof_long.ncu-rep gpurun_out/r2_ncu_summary_longwave.txt "# ncu --set full, scripts/prof_long.py: two split steps of a 2^24-sample waveform (64 x 2^18), outer k_col_mid + inner k_col_fwd / k_row / k_col_inv"
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize_r2.py > gpurun_out/r2_sanitizer_${tool}_r2paths.log 2>&1; echo "$tool r2 rc=$?"; tail -3 gpurun_out/r2_sanitizer_${tool}_r2paths.log
  timeout 1200 compute-sanitizer --tool $tool python scripts/sanitize_wf.py > gpurun_out/r2_sanitizer_${tool}_wf.log 2>&1; echo "$tool wf rc=$?"; tail -3 gpurun_out/r2_sanitizer_${tool}_wf.log
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_paths.py > gpurun_out/r2_sanitizer_${tool}_paths.log 2>&1; echo "$tool paths rc=$?"; tail -3 gpurun_out/r2_sanitizer_${tool}_paths.log
done
du -sh gpurun_out; ls -la gpurun_out | tail -30
