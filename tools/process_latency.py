"""process()-granularity latency: wall time per fdsp_bank_process(64) call vs the device time of its kernels."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from fundsp_b200 import workloads
from fundsp_b200.bank import GpuBank

name, V = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("saw_svf", 16384)
b = GpuBank(workloads.build(name, V), per_voice=False, mix=True, sample_rate=48000.0)
import ctypes as C
inp = np.zeros((max(1, b.inputs()), 64), np.float32)
out = np.zeros((b.outputs(), 64), np.float32)
ip, op = inp.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float))
for _ in range(50):
    b.L.fdsp_bank_process(b.h, 64, ip, op)
dev = []
N = 500
t0 = time.perf_counter()
for _ in range(N):
    b.L.fdsp_bank_process(b.h, 64, ip, op)
wall = (time.perf_counter() - t0) / N * 1e6
for _ in range(100):
    b.L.fdsp_bank_process(b.h, 64, ip, op)
    dev.append(b.last_kernel_ms())
print(f"{name} V={V}: wall {wall:.1f} us/call, device (first kernel start -> last kernel end) {np.mean(dev) * 1e3:.1f} us")
