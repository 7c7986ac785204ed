#!/bin/bash
# Round 2 profile capture (run under gpurun, 1 GPU): the ncu launch list of the default bench command and one --set full capture per
# kernel DESIGN.md §5 talks about. Outputs: gpurun_out/r02_launches_bench.csv, gpurun_out/r02_full_*.ncu-rep; `python tools/summarize_profiles.py r02`
# (run in the build container) turns them into profiles/r02_*.md and profiles/r02_traffic.json.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
cap() { local name=$1 re=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o gpurun_out/r02_full_$name "$@" > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log; }
cap saw_svf_mix "bank_kernel<" python tools/prof_bank.py --workload saw_svf --voices 16384 --mode mix --n 16384 --iters 3
cap noise_svf_mix "bank_kernel<" python tools/prof_bank.py --workload noise_svf --voices 16384 --mode mix --n 16384 --iters 3
cap fm_mix "bank_kernel<" python tools/prof_bank.py --workload fm --voices 4096 --mode mix --n 16384 --iters 3
cap subdry_st bank_kernel_st python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3
FDSP_NO_PIPELINE=1 cap fdn fdn_kernel python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
cap conv_tc conv_tc_kernel python tools/prof_convolver.py 16384 1000
ls -la gpurun_out | grep r02_
