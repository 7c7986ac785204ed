#!/bin/bash
# Run under gpurun: ncu launch list of the bench command + one --set full capture per headline kernel.
# Outputs land in gpurun_out/ (scratch); tools/summarize_profiles.py turns them into profiles/*.md + r01_traffic.json.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
for w in saw_svf noise_svf; do
  ncu --set full --clock-control none --import-source on -k regex:bank_kernel -s 2 -c 1 -o gpurun_out/full_${w}_mix python tools/prof_bank.py --workload $w --mode mix > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:bank_kernel -s 2 -c 1 -o gpurun_out/full_saw_svf_voices python tools/prof_bank.py --workload saw_svf --mode voices+mix > /dev/null 2>&1
ls -la gpurun_out
