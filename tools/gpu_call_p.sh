#!/bin/bash
# Round 2, call P (after the container was re-created): the whole GPU suite on HEAD, smoke, the default bench line and config 5.
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider --durations=8 > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log; tail -14 gpurun_out/p_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/p_smoke.log
for w in saw_svf net; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/p_bench_$w.json 2>> gpurun_out/p_err.log
  python -c "
import json
d = json.loads(open('gpurun_out/p_bench_$w.json').read().strip().splitlines()[-1])
print('$w value %.0f e2e %.0f proc %.1f us ms %.3f dom %.3f' % (d['value'], d['e2e']['value'], d['e2e']['process_granularity']['us_per_call'], d['ms_per_step'], d['roofline']['kernel_ms_per_step']))"
done
tail -3 gpurun_out/p_err.log
