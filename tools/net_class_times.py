"""Config 5 broken down: each of the Net's four voice classes alone (16 384 voices, one 16 384-sample launch, mix-down), then together.
Shows which class is the pole under the concurrent-streams launch. GPU box."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from fundsp_b200 import workloads
from fundsp_b200.bank import GpuBank

N = 16384
names = ["sine >> lowpass >> pan", "saw >> moog >> pan", "white >> bandpass >> pan", "fm >> highpass >> pan"]
def time_bank(voices, label, iters=4):
    b = GpuBank(voices, per_voice=False, mix=True, sample_rate=48000.0)
    c = b.voice_outputs()
    mix = torch.empty((c, N), device="cuda", dtype=torch.float32)
    best = 1e9
    for _ in range(iters):
        b.render_device(N, 0, N, 0, N, mix.data_ptr(), N, sync=True)
        best = min(best, b.last_kernel_ms())
    print(f"{label}: {best:.3f} ms per {len(voices)} x {N}  ({best * 1e-3 * 1.965e9 / N:.0f} cycles per sample at 1.965 GHz)", flush=True)
    return best
tot = 0.0
for k in range(4):
    tot += time_bank([workloads.net_voice(4 * j + k) for j in range(16384)], f"class {k} alone ({names[k]})")
print(f"sum of the four alone: {tot:.3f} ms")
time_bank([workloads.net_voice(i) for i in range(65536)], "all four classes, concurrent streams")
