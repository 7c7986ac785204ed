"""Wall time per `fdsp_bank_process(64)` call (host buffers), resident kernel (FDSP_RT=1, default) against one launch per block (FDSP_RT=0)."""
import os, sys, time
sys.path[:0] = [".", "tests"]
from fundsp_b200 import workloads
from fundsp_b200.bank import GpuBank
for name, V in (("saw_svf", 16384), ("noise_svf", 16384), ("fm", 4096), ("saw_svf", 1024)):
    for rt in ("1", "0"):
        os.environ["FDSP_RT"] = rt
        b = GpuBank(workloads.build(name, V), per_voice=False, mix=True, sample_rate=48000.0)
        for _ in range(20):
            b.process(64)
        t = time.perf_counter()
        for _ in range(500):
            b.process(64)
        dt = (time.perf_counter() - t) / 500
        print(f"{name} V={V} FDSP_RT={rt}: {dt * 1e6:.1f} us per process(64)  {V * 64 / dt / 1e9:.1f} Gsamples/s", flush=True)
        del b
