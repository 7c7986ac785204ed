"""Where does time go when two banks of one device alternate process() calls (the resident-kernel slot changes hands)? Prints every call
that takes more than 2 ms. Run on a GPU box: python tools/rt_handover_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fundsp_b200 import workloads
from fundsp_b200.bank import GpuBank

SR = 48000.0
a = GpuBank(workloads.build("saw_svf", 600), per_voice=False, mix=True, sample_rate=SR)
b = GpuBank(workloads.build("saw_svf", 500, first=600), per_voice=False, mix=True, sample_rate=SR)
a.process(64); b.process(64)
t_all = time.perf_counter()
k = 0
for rnd in range(6):
    for name, bank in (("a", a), ("b", b)):
        for j in range(4):
            t = time.perf_counter(); bank.process(64); dt = time.perf_counter() - t
            if dt > 2e-3:
                print(f"round {rnd} bank {name} call {j}: {dt * 1e3:.1f} ms")
            k += 1
t = time.perf_counter(); b.render_samples(300); print(f"render of b: {(time.perf_counter() - t) * 1e3:.1f} ms")
t = time.perf_counter(); a.process(33); print(f"a.process after it: {(time.perf_counter() - t) * 1e3:.1f} ms")
print(f"total {time.perf_counter() - t_all:.3f} s for {k} calls")
