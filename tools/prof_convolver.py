"""Kernel time of `noise >> convolve(h)`: V voices, K-tap shared impulse response, 16384 samples (device-resident), tensor-core form
(csrc/dsp/conv_tc_kernel.cuh) and, with FDSP_TC_CONV=0, the direct form on the FP32 pipe. TFLOP/s: `useful` = 2 V n K (the convolution's own
multiply-adds), `issued` = what the Toeplitz tiles execute: 3 (3xTF32) x 2 x V x n x (127 + K rounded up to 32)."""
import os
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
from fundsp_b200.prelude import noise, convolve
from fundsp_b200.bank import GpuBank

cases = [(1024, 64), (1024, 1000), (16384, 64), (16384, 1000), (16384, 4096)]
if len(sys.argv) > 2:
    cases = [(int(sys.argv[1]), int(sys.argv[2]))]
tc = os.environ.get("FDSP_TC_CONV", "1") != "0"
for V, K in cases:
    rng = np.random.default_rng(K)
    h = (rng.uniform(-1, 1, K) * np.exp(-np.arange(K) / (K / 4.0))).astype(np.float32)
    b = GpuBank([noise().seed(i) >> convolve(h) for i in range(V)], per_voice=False, mix=True, sample_rate=48000.0)
    n = 16384
    mix = torch.empty((1, n), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    for _ in range(3):
        b.render_device(n, 0, n, 0, n, mix.data_ptr(), n, sync=True)
    ms = b.last_kernel_ms()
    J = (127 + K + 31) // 32 * 32
    print(f"convolver[{'tensor' if tc and K >= 32 else 'direct'}] V={V} K={K}: {ms:.3f} ms per {n} samples  {V * n / ms / 1e6:.2f} Gsamples/s  useful {2.0 * V * n * K / ms / 1e9:.1f} TFLOP/s"
          + (f"  issued {3 * 2.0 * ((V + 127) // 128 * 128) * n * J / ms / 1e9:.1f} TFLOP/s (tf32)" if tc and K >= 32 else ""), flush=True)
