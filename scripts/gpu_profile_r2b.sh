#!/bin/bash
# evidence for the kernels added late in round 2: launch list of the bench command, DRAM traffic of k_wf at FULL size, full ncu
# captures of k_wf (16-CTA clusters + clusters of 2 with 8 tiles per CTA) and of k_ols, compute-sanitizer logs
set -u
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/r2b_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:k_wf -c 6 --csv --log-file gpurun_out/r2b_traffic_k_wf_4096rows.csv python bench.py --steps 1 --warmup 1 --no-extra > gpurun_out/r2b_ncu_traffic.log 2>&1; echo "traffic fp64 rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:k_wf -c 6 --csv --log-file gpurun_out/r2b_traffic_k_wf_fp32_4096rows.csv python bench.py --steps 1 --warmup 1 --no-extra --precision fp32 > gpurun_out/r2b_ncu_traffic32.log 2>&1; echo "traffic fp32 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wf -s 2 -c 2 -o /tmp/r2b_prof_wf_fp64 -f python scripts/prof_wf.py fp64 320 > gpurun_out/r2b_ncu_wf_fp64.log 2>&1; echo "ncu k_wf fp64 rc=$?"
python scripts/ncu_summarize.py /tmp/r2b_prof_wf_fp64.ncu-rep gpurun_out/r2b_ncu_summary_k_wf_fp64.txt "# ncu --set full --clock-control none --import-source on -k regex:k_wf -s 2 -c 2, scripts/prof_wf.py fp64 320: the two launches of ONE propagation of 320 config-#3 waveforms -- 14 clusters of 16 CTAs (TM = 1) and 36 clusters of 2 CTAs with 8 tiles per CTA (TM = 3).  ncu serialises the two launches: the first one processes every waveform alone, the second finds none left (see the next file for the small clusters)"
SSFM_MT_CS=2 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wf -s 2 -c 1 -o /tmp/r2b_prof_wf_mt2 -f python scripts/prof_wf.py fp64 320 > gpurun_out/r2b_ncu_wf_mt2.log 2>&1; echo "ncu k_wf mt2 rc=$?"
python scripts/ncu_summarize.py /tmp/r2b_prof_wf_mt2.ncu-rep gpurun_out/r2b_ncu_summary_k_wf_fp64_small_clusters.txt "# ncu --set full, SSFM_MT_CS=2 scripts/prof_wf.py fp64 320: k_wf<double,256,256,.,3> as the only team kind -- 40 clusters of 2 CTAs (80 of the 296 CTA slots), every CTA carrying 8 tiles per phase, adaptive steps (two passes per column phase)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ols|k_filtfilt|k_scatter|k_save|k_pd|k_fill_h2" -s 6 -c 12 -o /tmp/r2b_prof_ols -f python scripts/prof_ols.py 256 > gpurun_out/r2b_ncu_ols.log 2>&1; echo "ncu ols rc=$?"
python scripts/ncu_summarize.py /tmp/r2b_prof_ols.ncu-rep gpurun_out/r2b_ncu_summary_ols.txt "# ncu --set full, scripts/prof_ols.py 256 (256 frames x 2^18, fp64): second pass of BPF and PD -> LPF -> SAMPLER through the overlap-save kernel k_ols"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool python scripts/sanitize_r2b.py > gpurun_out/r2b_sanitizer_${tool}.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r2b_sanitizer_${tool}.log
done
du -sh gpurun_out; ls -la gpurun_out | tail -12
