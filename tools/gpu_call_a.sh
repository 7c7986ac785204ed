#!/bin/bash
# Round 2, GPU call A (run under gpurun from the repo root): parity of the stage-pipelined kernels and the TMA FDN kernel, their
# timings against the round-1 forms, and one ncu --set full capture of each. Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -25 gpurun_out/a_pytest.log
T=gpurun_out/a_timings.txt; : > $T
run() { echo "## $*" >> $T; env "$@" 2>&1 | tail -4 >> $T; }
run FDSP_STAGED=0 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3
run FDSP_STAGED=1 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3
run FDSP_STAGED=1 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode voices --n 16384 --iters 3
run FDSP_NO_PIPELINE=1 FDSP_STAGED=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run FDSP_STAGED=0 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
for pc in 1024 2048 4096 8192; do
  run FDSP_PIPE_CHUNK=$pc python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
done
run FDSP_PIPE_TRACE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 2
echo "## trace" >> $T; FDSP_PIPE_TRACE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 2 2>&1 | grep pipe | tail -10 >> $T
run FDSP_STAGED=0 python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 3
run FDSP_STAGED=1 python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 3
run python tools/prof_bank.py --workload saw_svf --voices 16384 --mode mix --n 16384 --iters 3
cat $T
cap() {  # name kernel-regex env... -- args
  local name=$1 re=$2; shift 2
  timeout 300 env "$@" ncu --set full --clock-control none --import-source on -k regex:$re -s 2 -c 1 -f -o gpurun_out/r02_full_$name python tools/prof_bank.py --workload ${WL} --voices 1024 --mode mix --n 16384 --iters 2 > gpurun_out/ncu_$name.log 2>&1
}
WL=subtractive cap fdn fdn_kernel FDSP_NO_PIPELINE=1
WL=subtractive_dry cap subdry_st bank_kernel_st FDSP_STAGED=1
ls -la gpurun_out | tail -8
