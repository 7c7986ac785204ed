#!/bin/bash
# Round 2, call O: the whole GPU suite (Net::crossfade, looping sequencer, resident-slot hand-over with rt_stop at every entry point), smoke, and
# config 5 with the shared carve-out now the default for concurrent classes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/o_pytest.log; tail -8 gpurun_out/o_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/o_smoke.log 2>&1; tail -2 gpurun_out/o_smoke.log
timeout 120 python tools/rt_handover_probe.py > gpurun_out/o_handover.txt 2>&1; cat gpurun_out/o_handover.txt
for w in net saw_svf subtractive; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/o_bench_$w.json 2>> gpurun_out/o_err.log
  python -c "
import json
d = json.loads(open('gpurun_out/o_bench_$w.json').read().strip().splitlines()[-1])
print('$w value %.0f e2e %.0f proc %.1f us ms %.3f dom %.3f' % (d['value'], d['e2e']['value'], d['e2e']['process_granularity']['us_per_call'], d['ms_per_step'], d['roofline']['kernel_ms_per_step']))"
done
FDSP_CARVEOUT=-1 timeout 300 python bench.py --steps 10 --warmup 3 --workload net 2>> gpurun_out/o_err.log | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('net (driver-chosen carve-outs) value %.0f ms %.3f' % (d['value'], d['ms_per_step']))"
tail -3 gpurun_out/o_err.log
