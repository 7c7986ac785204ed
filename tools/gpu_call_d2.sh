#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/d2_tma.log; : > $L
for v in 0 1 2 3 4 5; do timeout 30 tests/cpp/_probe/tma_probe $v >> $L 2>&1; echo "rc=$?" >> $L; done
nvidia-smi --query-gpu=driver_version,name --format=csv >> $L
cat $L
