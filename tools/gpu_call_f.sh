#!/bin/bash
# Round 2, GPU call F: whole GPU suite (join fix of the solo two-stage pipeline, four-accumulator convolver), probe accuracy, timings, profiles.
mkdir -p gpurun_out
L=gpurun_out/f_probe.log; : > $L
for args in "4 1000 300 900" "4 4096 128 1024"; do timeout 60 tests/cpp/_probe/conv_tc_probe $args >> $L 2>&1; echo "rc=$?" >> $L; done
cat $L
timeout 1200 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log; tail -15 gpurun_out/f_pytest.log
T=gpurun_out/f_timings.txt; : > $T
run() { echo "## $*" >> $T; timeout 300 env "$@" 2>&1 | tail -${TAIL:-2} >> $T; }
run FDSP_NO_PIPELINE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run FDSP_FDN_FLAGS=1 FDSP_NO_PIPELINE=1 python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
TAIL=6 run python tools/prof_convolver.py
cat $T
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench_saw_svf.json 2> gpurun_out/f_bench_err.log; tail -c 1500 gpurun_out/f_bench_saw_svf.json
timeout 300 python bench.py --steps 5 --warmup 3 --workload subtractive > gpurun_out/f_bench_subtractive.json 2>> gpurun_out/f_bench_err.log; tail -c 700 gpurun_out/f_bench_subtractive.json
timeout 300 python bench.py --steps 5 --warmup 3 --workload conv > gpurun_out/f_bench_conv.json 2>> gpurun_out/f_bench_err.log; tail -c 700 gpurun_out/f_bench_conv.json
tail -5 gpurun_out/f_bench_err.log
