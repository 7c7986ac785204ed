#!/bin/bash
# First GPU call of a round (run under gpurun from the repo root): the whole GPU suite WITHOUT -x and with the failure
# text kept (the round-1 call that stopped at the first failure returned only pytest's last four lines), then smoke, one bench
# line and the sequencer workload timing. Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_line.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_line.json
timeout 200 python tools/prof_sequencer.py > gpurun_out/sequencer.log 2>&1; tail -5 gpurun_out/sequencer.log
timeout 300 python bench.py --workload saw_svf_events > gpurun_out/bench_events.json 2>> gpurun_out/bench_err.log; tail -c 400 gpurun_out/bench_events.json
