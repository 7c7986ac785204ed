"""Config 4 both ways: per-voice reverb (the stress form) vs the musically typical shared-bus reverb, where
the V dry voices are mixed to stereo first and ONE reverb_stereo instance processes the mix (examples/keys.rs:164-181)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
from fundsp_b200 import workloads
from fundsp_b200.prelude import multipass, reverb_stereo
from fundsp_b200.bank import GpuBank

V, n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 48000
gate = torch.from_numpy(workloads.gate_signal(n)).cuda()
dry = GpuBank(workloads.build("subtractive_dry", V), per_voice=False, mix=True, sample_rate=48000.0)
bus = GpuBank([multipass(2) & 0.2 * reverb_stereo(10.0, 2.0, 0.5)], per_voice=False, mix=True, sample_rate=48000.0)
per_voice = GpuBank(workloads.build("subtractive", V), per_voice=False, mix=True, sample_rate=48000.0)
mix_a = torch.zeros((2, n), device="cuda"); mix_b = torch.zeros((2, n), device="cuda"); mix_c = torch.zeros((2, n), device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def shared_bus():
    dry.render_device(n, gate.data_ptr(), n, 0, n, mix_a.data_ptr(), n, sync=True)
    bus.render_device(n, mix_a.data_ptr(), n, 0, n, mix_b.data_ptr(), n, sync=True)


def stress():
    per_voice.render_device(n, gate.data_ptr(), n, 0, n, mix_c.data_ptr(), n, sync=True)


for name, fn in (("shared-bus reverb (dry bank -> mix -> 1 reverb)", shared_bus), ("per-voice reverb (stress form)", stress)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.2f} ms per 1 s of audio, {V} voices -> {V * n / ms / 1e6:.2f} Gsamples/s", flush=True)
