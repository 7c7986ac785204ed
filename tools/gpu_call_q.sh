#!/bin/bash
# Round 2, call Q: (1) where the Moog ladder's per-sample cycles go (tools/probe/moog_chain_probe.cu: plain vs fast tanh, diagnostics, and the
# 2^32-argument sweep fast == plain); (2) config 5 class by class; (3) the product with the fast tanh form (default) against the plain form
# (variants/tanh_plain.so) on the Moog-bound workloads; (4) the whole GPU suite on the new default; (5) bench lines.
mkdir -p gpurun_out
timeout 120 tools/probe/_build/moog_chain_probe > gpurun_out/q_probe.txt 2>&1; cat gpurun_out/q_probe.txt
timeout 200 python tools/net_class_times.py > gpurun_out/q_net_classes.txt 2>&1; cat gpurun_out/q_net_classes.txt
rm -f gpurun_out/q_ab.txt
for lib in fundsp_b200/libfundsp_b200.so fundsp_b200/variants/tanh_plain.so; do
  echo "== $lib" >> gpurun_out/q_ab.txt
  for w in "subtractive_dry 1024" "subtractive 1024" "net 65536"; do
    set -- $w
    FDSP_B200_LIB=$PWD/$lib timeout 120 python tools/prof_bank.py --workload $1 --voices $2 --mode mix --n 16384 --iters 3 2>&1 | tail -1 >> gpurun_out/q_ab.txt
  done
  FDSP_STAGED=0 FDSP_B200_LIB=$PWD/$lib timeout 120 python tools/prof_bank.py --workload subtractive_dry --voices 1024 --mode mix --n 16384 --iters 3 2>&1 | tail -1 | sed 's/^/plain kernel: /' >> gpurun_out/q_ab.txt
done
cat gpurun_out/q_ab.txt
timeout 540 python -m pytest tests -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/q_pytest.log; tail -6 gpurun_out/q_pytest.log
for w in subtractive net; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/q_bench_$w.json 2>> gpurun_out/q_err.log
  python -c "
import json
d = json.loads(open('gpurun_out/q_bench_$w.json').read().strip().splitlines()[-1])
print('$w value %.0f e2e %.0f ms %.3f dom %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step']))"
done
tail -3 gpurun_out/q_err.log
