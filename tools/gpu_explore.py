"""Scratch GPU exploration: error statistics + quick timings for every workload (run under gpurun)."""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "..")
import numpy as np
from fundsp_b200 import workloads
from fundsp_b200.bank import GpuBank
from oracle import oracle_bank_render, lib as olib
olib().fo_set_denormal_emulation(0)
SR = 48000.0
res = {}
for name, V, n in (("noise_svf", 256, 4800), ("fm", 256, 4800), ("saw_svf", 256, 4800), ("biquad_bank", 32, 4800), ("net", 256, 4800),
                   ("subtractive_dry", 64, 24000), ("subtractive", 16, 9600)):
    try:
        inp = workloads.gate_signal(n) if name.startswith("subtractive") else None
        t = time.time(); b = GpuBank(workloads.build(name, V), per_voice=True, sample_rate=SR); g, _ = b.render_samples(n, inp); tg = time.time() - t
        t = time.time(); o, _ = oracle_bank_render(workloads.build(name, V), SR, n, inp, threads=4); to = time.time() - t
        d = np.abs(g.astype(np.float64) - o)
        peak = np.abs(o).max(axis=-1, keepdims=True)
        rel = (d / np.maximum(np.abs(o), 1e-2 * np.maximum(peak, 1e-30))).max()
        res[name] = dict(max_abs=float(d.max()), rel=float(rel), exact=bool(np.array_equal(g, o)), nonequal=int((g != o).sum()), peak=float(np.abs(o).max()), gpu_s=tg, cpu_s=to)
    except Exception as e:
        res[name] = dict(error=str(e))
    print(name, res[name], flush=True)
# timing of the headline at full size (device-resident)
import torch
for name, V in (("saw_svf", 16384), ("noise_svf", 16384), ("fm", 4096)):
    n = 48000
    b = GpuBank(workloads.build(name, V), per_voice=True, mix=True, sample_rate=SR)
    out = torch.empty((V, n), device="cuda", dtype=torch.float32)
    mix = torch.empty((1, n), device="cuda", dtype=torch.float32)
    for mode in ("voices+mix", "mix"):
        for it in range(3):
            b.render_device(n, 0, 0, out.data_ptr() if mode != "mix" else 0, n, mix.data_ptr(), n, sync=True)
            ms = b.last_kernel_ms()
        print(name, V, mode, f"{ms:.3f} ms  {V * n / ms / 1e6:.1f} Gsamples/s", flush=True)
json.dump(res, open("gpurun_out/explore.json", "w"), indent=1)
