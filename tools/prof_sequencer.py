"""Timing of a sequencer bank (GPU box): V overlapping note events of one instrument class rendered in one long call, against the same
voices as a plain bank (what the per-block scheduling arithmetic of Event<X> costs), and the latency of a note-on into a running bank."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fundsp_b200.bank import GpuBank  # noqa: E402
from fundsp_b200.prelude import lowpass_hz, saw_hz  # noqa: E402
from fundsp_b200.sequencer import Fade, event  # noqa: E402

SR = 48000.0
V = int(sys.argv[sys.argv.index("--voices") + 1]) if "--voices" in sys.argv else 16384
N = int(sys.argv[sys.argv.index("--samples") + 1]) if "--samples" in sys.argv else 48000
rng = np.random.default_rng(1)
voice = lambda i: saw_hz(55.0 * 2.0 ** (5.0 * ((i * 0.6180339887) % 1.0))) >> lowpass_hz(1500.0 + (i % 64) * 20.0, 1.0)
starts = rng.uniform(0.0, 0.6, V)
durs = rng.uniform(0.1, 0.4, V)


def timed(bank, n):
    bank.render_samples(64)          # JIT + warm-up
    bank.reset()
    t = time.perf_counter()
    bank.render_samples(n)
    return time.perf_counter() - t, bank.last_kernel_ms()


plain = GpuBank([voice(i) for i in range(V)], per_voice=False, mix=True, sample_rate=SR)
seq = GpuBank([event(voice(i), starts[i], starts[i] + durs[i], Fade.Smooth, 0.005, 0.02) for i in range(V)], per_voice=False, mix=True, sample_rate=SR)
for name, b in (("plain bank", plain), ("sequencer bank", seq)):
    wall, ms = timed(b, N)
    print(f"{name}: {V} voices x {N} samples: wall {wall * 1e3:.2f} ms, kernels {ms:.2f} ms, {V * N / (max(ms, 1e-9) * 1e-3) / 1e9:.1f} Gsample/s (voice-samples incl. silent ones)")
seq.reset()
seq.render_samples(max(N, 48000))    # everything has ended: every voice is a free slot
t = time.perf_counter()
K = min(200, V)
for k in range(K):
    seq.push_event(event(voice(k), seq.time() + 0.001, seq.time() + 0.2, Fade.Smooth, 0.005, 0.02))
print(f"note-on into a running bank: {(time.perf_counter() - t) / K * 1e6:.1f} us per push_event (lowering + 3 strided uploads + sync)")
