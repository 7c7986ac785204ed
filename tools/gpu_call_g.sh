#!/bin/bash
# Round 2, GPU call G: the shortened Moog chain (unrolled heavy stage, branch-free tanhf, float-side argument reduction): parity + time.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_jit.py -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_pytest.log; tail -8 gpurun_out/g_pytest.log
T=gpurun_out/g_timings.txt; : > $T
run() { echo "## $*" >> $T; timeout 300 env "$@" 2>&1 | tail -${TAIL:-2} >> $T; }
run FDSP_STAGED=1 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode voices --n 16384 --iters 3
run FDSP_STAGED=0 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode voices --n 16384 --iters 3
run python tools/prof_bank.py --workload subtractive --voices 1024 --mode mix --n 16384 --iters 3
run FDSP_STAGED=0 python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 3
run FDSP_STAGED=1 python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 3
cat $T
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_st -s 1 -c 1 -f -o gpurun_out/r02_full_subdry_st python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3 > gpurun_out/ncu_subdry_st.log 2>&1; tail -2 gpurun_out/ncu_subdry_st.log
