#!/bin/bash
# Round 2, last 1-GPU call: the resident kernel's fold (back to the in-order L2 fold, 8 loads in flight), the two-level Net mix, leftover captures.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -rf -p no:cacheprovider -k "resident or net_of_voices or full_size or process" > gpurun_out/last_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/last_pytest.log; tail -6 gpurun_out/last_pytest.log
timeout 200 python tools/process_latency.py > gpurun_out/last_latency.txt 2>&1; cat gpurun_out/last_latency.txt
for w in net saw_svf; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/final_bench_$w.json 2>> gpurun_out/last_err.log
  python -c "
import json
d = json.loads(open('gpurun_out/final_bench_$w.json').read().strip().splitlines()[-1])
print('$w value %.0f e2e %.0f proc %.1f us ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['process_granularity']['us_per_call'], d['ms_per_step']))"
done
FDSP_NO_DOM=1 timeout 300 python bench.py --steps 10 --warmup 3 --workload net 2>> gpurun_out/last_err.log | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('net (no dominant-kernel events) value %.0f ms %.3f' % (d['value'], d['ms_per_step']))"
cap() { local name=$1 re=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o gpurun_out/r02_full_$name "$@" > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log; }
cap saw_svf_mix bank_kernel python tools/prof_bank.py --workload saw_svf --voices 16384 --mode mix --n 16384 --iters 3
cap noise_svf_mix bank_kernel python tools/prof_bank.py --workload noise_svf --voices 16384 --mode mix --n 16384 --iters 3
cap fm_mix bank_kernel python tools/prof_bank.py --workload fm --voices 4096 --mode mix --n 16384 --iters 3
tail -3 gpurun_out/last_err.log
