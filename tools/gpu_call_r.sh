#!/bin/bash
# Round 2, call R: the Moog stage after (a) moving the coefficient recomputation out of line (73 KB -> 22 KB of code per 8-sample group),
# (b) the steady-group path (no per-sample "input changed" branch), and the level-2 fast tanh (sign off the chain, variants/tanh2.so):
# probe sweep + cycles, A/B timings, the whole GPU suite on the product, one ncu --set full capture of the staged dry kernel.
mkdir -p gpurun_out
timeout 120 tools/probe/_build/moog_chain_probe --sweep > gpurun_out/r_probe.txt 2>&1; cat gpurun_out/r_probe.txt
rm -f gpurun_out/r_ab.txt
for lib in fundsp_b200/libfundsp_b200.so fundsp_b200/variants/tanh2.so; do
  echo "== $lib" >> gpurun_out/r_ab.txt
  for w in "subtractive_dry 1024" "subtractive 1024" "net 65536"; do
    set -- $w
    FDSP_B200_LIB=$PWD/$lib timeout 120 python tools/prof_bank.py --workload $1 --voices $2 --mode mix --n 16384 --iters 3 2>&1 | tail -1 >> gpurun_out/r_ab.txt
  done
  FDSP_STAGED=0 FDSP_B200_LIB=$PWD/$lib timeout 120 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3 2>&1 | tail -1 | sed 's/^/plain kernel: /' >> gpurun_out/r_ab.txt
done
FDSP_B200_LIB=$PWD/fundsp_b200/variants/tanh2.so timeout 200 python tools/net_class_times.py 2>&1 | sed -n 2p >> gpurun_out/r_ab.txt
cat gpurun_out/r_ab.txt
timeout 540 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r_pytest.log; tail -6 gpurun_out/r_pytest.log
FDSP_B200_LIB=$PWD/fundsp_b200/variants/tanh2.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wider.py -m gpu -q --tb=short -p no:cacheprovider -k "subtractive or moog or net or staged or shaper or tanh or full_size" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_st -s 1 -c 1 -f -o gpurun_out/r02_full_subdry_st python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3 > gpurun_out/ncu_subdry_st.log 2>&1; tail -1 gpurun_out/ncu_subdry_st.log
