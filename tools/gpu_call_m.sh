#!/bin/bash
# Round 2, call M: whole GPU suite after the Event<X> layout change (ReplayMode::Loop) and the resident-slot registry; config 5 with the heavy
# class launched first (A/B: FDSP_NO_HEAVY_FIRST, FDSP_NO_DOM); the bench launch list without the resident kernel (ncu serialises kernels).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/m_pytest.log; tail -8 gpurun_out/m_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; tail -2 gpurun_out/m_smoke.log
line() { python -c "
import json,sys
d = json.loads(open('$1').read().strip().splitlines()[-1])
print('$2 value %.0f e2e %.0f ms %.3f kernel_ms %.3f dom_ms %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_all_per_step'], d['roofline']['kernel_ms_per_step'], d['gpu_launches']))"; }
timeout 200 python bench.py --steps 10 --warmup 3 --workload net > gpurun_out/m_bench_net.json 2>> gpurun_out/m_err.log; line gpurun_out/m_bench_net.json "net (heavy class first)"
FDSP_NO_HEAVY_FIRST=1 timeout 200 python bench.py --steps 10 --warmup 3 --workload net > gpurun_out/m_bench_net_nohf.json 2>> gpurun_out/m_err.log; line gpurun_out/m_bench_net_nohf.json "net (class order)"
FDSP_NO_DOM=1 timeout 200 python bench.py --steps 10 --warmup 3 --workload net > gpurun_out/m_bench_net_nodom.json 2>> gpurun_out/m_err.log; line gpurun_out/m_bench_net_nodom.json "net (no dominant-kernel events)"
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/m_bench_saw_svf.json 2>> gpurun_out/m_err.log; line gpurun_out/m_bench_saw_svf.json "saw_svf"
timeout 100 python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 3 2>&1 | tail -2
FDSP_RT=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/r02_bench_under_ncu.log | head -3; wc -l gpurun_out/r02_launches_bench.csv
FDSP_RT=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_net.csv python tools/prof_bank.py --workload net --voices 65536 --mode mix --n 16384 --iters 2 > gpurun_out/m_net_under_ncu.log 2>&1; wc -l gpurun_out/r02_launches_net.csv
tail -3 gpurun_out/m_err.log
