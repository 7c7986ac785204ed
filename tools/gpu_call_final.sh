#!/bin/bash
# Round 2, final 1-GPU call: the whole GPU suite, smoke, one bench line per workload, process() latency, then the profile captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest.log; tail -8 gpurun_out/final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
for w in saw_svf subtractive net conv noise_svf fm; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/final_bench_$w.json 2>> gpurun_out/final_bench_err.log
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/final_bench_$w.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$w value %.0f e2e %.0f proc %.1f us roof %s %.2f frac %.4f cpu %s" % (d["value"], d["e2e"]["value"], d["e2e"]["process_granularity"]["us_per_call"], r["bound"], r["achieved"], r["frac"], d["cpu_baseline"]["value"]))
except Exception as e:
    print("$w FAILED", e)
PY
done
timeout 200 python tools/process_latency.py > gpurun_out/final_latency.txt 2>&1; cat gpurun_out/final_latency.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench_err.log; tail -c 600 gpurun_out/final_bench_reference.json
tail -3 gpurun_out/final_bench_err.log
bash tools/capture_profiles_r02.sh > gpurun_out/final_capture.log 2>&1; tail -8 gpurun_out/final_capture.log
