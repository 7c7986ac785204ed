#!/bin/bash
# Run under gpurun: ncu launch list of the bench command + one --set full capture per headline kernel.
# Outputs land in gpurun_out/ (scratch); tools/summarize_profiles.py turns them into profiles/*.md + r01_traffic.json.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/full_*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
cap() {  # name kernel-regex workload voices mode
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/full_$1 python tools/prof_bank.py --workload $3 --voices $4 --mode $5 > /dev/null 2>&1
}
cap saw_svf_mix bank_kernel saw_svf 16384 mix
cap saw_svf_voices bank_kernel saw_svf 16384 voices
cap noise_svf_mix bank_kernel noise_svf 16384 mix
cap fm_mix bank_kernel fm 4096 mix
cap subdry bank_kernel subtractive_dry 1024 mix
FDSP_NO_PIPELINE=1 cap fdn fdn_kernel subtractive 1024 mix
ls -la gpurun_out
