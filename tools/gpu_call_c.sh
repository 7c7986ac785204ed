#!/bin/bash
# Round 2, GPU call C: the tensor-core convolver (first contact), the division-free / barrier-free FDN kernel, the reordered dry -> FDN pipeline.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -q --tb=short -rf -p no:cacheprovider -k "convol" > gpurun_out/c_pytest_conv.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest_conv.log; tail -30 gpurun_out/c_pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log; tail -12 gpurun_out/c_pytest.log
T=gpurun_out/c_timings.txt; : > $T
run() { echo "## $*" >> $T; timeout 300 env "$@" 2>&1 | tail -${TAIL:-2} >> $T; }
run FDSP_NO_PIPELINE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run FDSP_PIPE_CHUNK=4096 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
echo "## trace" >> $T; FDSP_PIPE_TRACE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 2 2>&1 | grep pipe | tail -4 >> $T
TAIL=6 run python tools/prof_convolver.py
TAIL=3 run FDSP_TC_CONV=0 python tools/prof_convolver.py 16384 1000
cat $T
cap() {  # name kernel-regex env... -- command
  local name=$1 re=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o gpurun_out/r02_full_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
}
FDSP_NO_PIPELINE=1 cap fdn2 fdn_kernel python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
cap conv_tc conv_tc_kernel python tools/prof_convolver.py 16384 1000
ls -la gpurun_out | grep r02
