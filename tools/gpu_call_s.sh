#!/bin/bash
# Round 2, call S (final): the whole GPU suite and smoke on the final code (tanh level 2 default, out-of-line Moog coefficients, steady group, reference-shaped
# reset), the bench lines the docs quote, the reference arm, the ncu launch list of the default bench command and one --set full capture of the staged dry kernel.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s_pytest.log; tail -5 gpurun_out/s_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s_smoke.log
for w in saw_svf subtractive net; do
  timeout 200 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/s_bench_$w.json 2>> gpurun_out/s_err.log
  python -c "
import json
d = json.loads(open('gpurun_out/s_bench_$w.json').read().strip().splitlines()[-1])
print('$w value %.0f e2e %.0f proc %.1f us ms %.3f dom %.3f cpu %.0f' % (d['value'], d['e2e']['value'], d['e2e']['process_granularity']['us_per_call'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['cpu_baseline']['value']))"
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s_bench_reference.json 2>> gpurun_out/s_err.log; tail -c 400 gpurun_out/s_bench_reference.json; echo
FDSP_RT=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r02_launches_bench.csv | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bank_kernel_st -s 1 -c 1 -f -o gpurun_out/r02_full_subdry_st python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3 > gpurun_out/ncu_subdry_st.log 2>&1; tail -1 gpurun_out/ncu_subdry_st.log
tail -3 gpurun_out/s_err.log
