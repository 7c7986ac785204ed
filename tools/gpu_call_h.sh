#!/bin/bash
# Round 2, GPU call H: the resident process() kernel (doorbell): parity against the launch-per-block path, latency per block.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -rf -p no:cacheprovider -k "resident or process or sequencer" > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log; tail -12 gpurun_out/h_pytest.log
timeout 300 python -m pytest tests/test_gpu_wider.py -m gpu -q --tb=short -rf -p no:cacheprovider > gpurun_out/h_pytest_wider.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest_wider.log; tail -6 gpurun_out/h_pytest_wider.log
cat > /tmp/lat.py <<'PY'
import sys, time, os
sys.path[:0] = [".", "tests"]
import numpy as np
from fundsp_b200 import workloads
from fundsp_b200.bank import GpuBank
for name, V in (("saw_svf", 16384), ("noise_svf", 16384), ("fm", 4096), ("saw_svf", 1024)):
    for rt in ("1", "0"):
        os.environ["FDSP_RT"] = rt
        b = GpuBank(workloads.build(name, V), per_voice=False, mix=True, sample_rate=48000.0)
        for _ in range(20): b.process(64)
        t = time.perf_counter()
        for _ in range(500): b.process(64)
        dt = (time.perf_counter() - t) / 500
        print(f"{name} V={V} FDSP_RT={rt}: {dt * 1e6:.1f} us per process(64)  {V * 64 / dt / 1e9:.1f} Gsamples/s", flush=True)
        del b
PY
timeout 300 python /tmp/lat.py > gpurun_out/h_latency.txt 2>&1; cat gpurun_out/h_latency.txt
