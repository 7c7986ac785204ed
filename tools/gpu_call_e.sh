#!/bin/bash
# Round 2, GPU call E: tensor-core convolver after the 16-byte window alignment fix (TMA tile starts must be 16-byte aligned), FDN two-sample form.
mkdir -p gpurun_out
L=gpurun_out/e_probe.log; : > $L
for args in "4" "4 1000 300 900" "4 33 130 128" "4 4096 128 1024"; do timeout 60 tests/cpp/_probe/conv_tc_probe $args >> $L 2>&1; echo "rc=$?" >> $L; done
cat $L
timeout 600 python -m pytest tests/test_gpu_jit.py -m gpu -q --tb=short -rf -p no:cacheprovider -k "convol" > gpurun_out/e_pytest_conv.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest_conv.log; tail -15 gpurun_out/e_pytest_conv.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -rf -p no:cacheprovider -k "fdn or subtractive or reverb or full_size or mixed" > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log; tail -8 gpurun_out/e_pytest.log
T=gpurun_out/e_timings.txt; : > $T
run() { echo "## $*" >> $T; timeout 300 env "$@" 2>&1 | tail -${TAIL:-2} >> $T; }
run FDSP_NO_PIPELINE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
TAIL=6 run python tools/prof_convolver.py
cat $T
cap() { local name=$1 re=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o gpurun_out/r02_full_$name "$@" > gpurun_out/ncu_$name.log 2>&1; tail -2 gpurun_out/ncu_$name.log; }
FDSP_NO_PIPELINE=1 cap fdn3 fdn_kernel python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
cap conv_tc conv_tc_kernel python tools/prof_convolver.py 16384 1000
ls -la gpurun_out | grep r02
