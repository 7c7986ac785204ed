#!/bin/bash
# Round 2, call T (last): the whole GPU suite, smoke and the default bench line on the final code (parked `latest` edits for Slot and Net::crossfade).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log; tail -5 gpurun_out/t_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/t_smoke.log
timeout 120 python bench.py > gpurun_out/t_bench_default.json 2>> gpurun_out/t_err.log; tail -c 300 gpurun_out/t_bench_default.json; echo; tail -2 gpurun_out/t_err.log
