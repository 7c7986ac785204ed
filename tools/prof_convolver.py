"""Kernel time of the direct-form Convolver: V voices, K-tap shared impulse response, 16384 samples (device-resident)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
from fundsp_b200.prelude import noise, convolve
from fundsp_b200.bank import GpuBank

for V, K in ((1024, 64), (1024, 1000), (16384, 64), (16384, 1000)):
    rng = np.random.default_rng(K)
    h = (rng.uniform(-1, 1, K) * np.exp(-np.arange(K) / (K / 4.0))).astype(np.float32)
    b = GpuBank([noise().seed(i) >> convolve(h) for i in range(V)], per_voice=False, mix=True, sample_rate=48000.0)
    n = 16384
    mix = torch.empty((1, n), device="cuda", dtype=torch.float32)
    for _ in range(3):
        b.render_device(n, 0, n, 0, n, mix.data_ptr(), n, sync=True)
    ms = b.last_kernel_ms()
    print(f"convolver V={V} K={K}: {ms:.3f} ms per {n} samples  {V * n / ms / 1e6:.2f} Gsamples/s  {2.0 * V * n * K / ms / 1e9:.2f} TFLOP/s", flush=True)
