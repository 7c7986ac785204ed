#!/bin/bash
# Round 2, GPU call D: bring-up of the tensor-core convolver kernel, level by level (tests/cpp/conv_tc_probe.cu).
mkdir -p gpurun_out
L=gpurun_out/d_probe.log; : > $L
for s in 1 2 3 4; do timeout 60 tests/cpp/_probe/conv_tc_probe $s >> $L 2>&1; echo "rc=$?" >> $L; done
timeout 120 compute-sanitizer --tool memcheck tests/cpp/_probe/conv_tc_probe 4 2>&1 | head -60 >> $L
for alt in tests/cpp/_probe/conv_tc_probe_*; do [ -x "$alt" ] || continue; echo "== $alt" >> $L; for s in 3 4; do timeout 60 $alt $s >> $L 2>&1; echo "rc=$?" >> $L; done; done
timeout 60 tests/cpp/_probe/conv_tc_probe 4 1000 300 900 >> $L 2>&1
cat $L
