#!/bin/bash
# Round 2, GPU call B: parity (incl. the BASELINE-size runs), timings of the cp.async FDN kernel and the staged dry kernel, ncu captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -15 gpurun_out/b_pytest.log
T=gpurun_out/b_timings.txt; : > $T
run() { echo "## $*" >> $T; env "$@" 2>&1 | tail -2 >> $T; }
run FDSP_STAGED=1 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode voices --n 16384 --iters 3
run FDSP_STAGED=1 FDSP_TB_MIN=100000 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode voices --n 16384 --iters 3
run FDSP_NO_PIPELINE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
echo "## trace" >> $T; FDSP_PIPE_TRACE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 2 2>&1 | grep pipe | tail -4 >> $T
cat $T
cap() {  # name kernel-regex workload env...
  local name=$1 re=$2 wl=$3; shift 3
  timeout 300 env "$@" ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o gpurun_out/r02_full_$name python tools/prof_bank.py --workload $wl --voices 1024 --mode mix --n 16384 --iters 3 > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
}
cap fdn fdn_kernel subtractive FDSP_NO_PIPELINE=1
cap subdry_st bank_kernel_st subtractive_dry FDSP_STAGED=1
ls -la gpurun_out | grep r02
