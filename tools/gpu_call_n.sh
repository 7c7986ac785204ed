#!/bin/bash
# Round 2, call N: (1) where the second of the two-bank hand-over test goes; (2) config 5 (four concurrent classes) with one shared-memory
# carve-out for all class kernels, with and without the heavy class first, against class order and against serial classes.
mkdir -p gpurun_out
timeout 120 python tools/rt_handover_probe.py > gpurun_out/n_handover.txt 2>&1; cat gpurun_out/n_handover.txt
run() { echo "## $*" >> gpurun_out/n_timings.txt; env "$@" timeout 100 python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 4 2>&1 | tail -3 >> gpurun_out/n_timings.txt; }
rm -f gpurun_out/n_timings.txt
run X=1
run FDSP_NO_CONCURRENT=1
run FDSP_CARVEOUT=100
run FDSP_CARVEOUT=100 FDSP_HEAVY_FIRST=1
run FDSP_HEAVY_FIRST=1
run X=2
cat gpurun_out/n_timings.txt
for v in "X=1" "FDSP_CARVEOUT=100" "FDSP_CARVEOUT=100 FDSP_HEAVY_FIRST=1" "FDSP_NO_CONCURRENT=1"; do
  env $v timeout 200 python bench.py --steps 10 --warmup 3 --workload net 2>> gpurun_out/n_err.log | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('net [$v] value %.0f ms %.3f' % (d['value'], d['ms_per_step']))"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -k "two_banks or resident or net_of" 2>&1 | tail -4
tail -3 gpurun_out/n_err.log
