"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU): real GpuBank shards, the mix-down through
the C ABI (`fdsp_group_*` / `fdsp_bank_render_reduced`: NCCL over NVLink + rank-order fold), rank 0 checks the result against the
oracle's index-order mix. The gloo process group only carries the 128-byte NCCL id."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch.distributed as dist  # noqa: E402

from fundsp_b200 import workloads  # noqa: E402
from fundsp_b200.bank import GpuBank  # noqa: E402
from fundsp_b200.parallel import BankGroup, ShardedBank, shard_range  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
group = BankGroup.from_torch_distributed()
ok = True
SR = 48000.0
for name, fn, total, n in (("saw_svf", workloads.saw_svf_voice, 301, 64 * 40 + 5), ("net", workloads.net_voice, 203, 16384 + 64 * 3 + 9),
                           ("subtractive", workloads.subtractive_voice, 2 * world + 3, 4800 + 7)):
    gate = workloads.gate_signal(n) if name == "subtractive" else None
    sb = ShardedBank(fn, total, sample_rate=SR, group=group)
    mix = sb.render_mix(n, gate)
    # the same shard again through the device entry points: partial in HBM, reduce_device, root reads it back
    import torch
    torch.cuda.set_device(sb.device)
    c = sb.bank.voice_outputs()
    dm = torch.zeros((c, n), device="cuda", dtype=torch.float32)
    sb.bank.reset()
    gin = torch.from_numpy(gate).cuda() if gate is not None else None
    sb.bank.render_device(n, gin.data_ptr() if gin is not None else 0, n, 0, n, dm.data_ptr(), n, sync=False)
    group.reduce_device(sb.bank, n, dm.data_ptr(), n, 0)
    sb.bank.sync()
    if rank == 0:
        from oracle import oracle_bank_render
        ref, _ = oracle_bank_render([fn(i) for i in range(total)], SR, n, gate, threads=4)
        want = ref.astype(np.float64).sum(0)
        # sum of V f32 terms in a fixed order vs f64: sqrt(V) * eps * sum|x|; floor at 1e-2 of the peak
        bound = 4.0 * np.sqrt(total) * 2.0 ** -24 * np.abs(ref.astype(np.float64)).sum(0) + 1e-7 * np.abs(want).max()
        good = bool(np.all(np.abs(mix - want) <= bound)) and np.abs(want).max() > 0.05
        same = bool(np.array_equal(dm.cpu().numpy(), mix))
        # rank-order fold: the sum must equal (shard 0 mix) + (shard 1 mix) + ... exactly, each shard's mix being what a 1-GPU bank of that shard gives
        parts = []
        for r in range(world):
            f, cnt = shard_range(total, r, world)
            _, pm = GpuBank([fn(f + i) for i in range(cnt)], device=sb.device, per_voice=False, mix=True, sample_rate=SR).render_samples(n, gate)
            parts.append(pm)
        acc = parts[0].copy()
        for pm in parts[1:]:
            acc = acc + pm
        exact = bool(np.array_equal(acc, mix))
        print(f"{name}: {total} voices on {world} ranks: bound {'ok' if good else 'MISMATCH'} device-path {'ok' if same else 'MISMATCH'} rank-order {'ok' if exact else 'MISMATCH'}", flush=True)
        ok = ok and good and same and exact
flag = [ok]
dist.broadcast_object_list(flag, src=0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag[0] else 1)
