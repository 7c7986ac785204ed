"""Worker of tests/test_mock_bank_cpu.py::test_sharded_bank_on_the_mock_device (launched by torch.distributed.run, gloo, 2 ranks): the
real ShardedBank + GpuBank code path of bench.py --gpus N with the mock device behind the C ABI; rank 0 checks the reduced mix."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch.distributed as dist  # noqa: E402

from fundsp_b200 import workloads  # noqa: E402
from fundsp_b200.parallel import ShardedBank  # noqa: E402

dist.init_process_group("gloo")
total, n = 37, 64 * 9 + 5
ok = True
from fundsp_b200.sequencer import event  # noqa: E402

early_events = lambda i: event(workloads.saw_svf_voice(i), (i % 7) * 0.0011, 1.0e6, 1, 0.003, 0.0)   # noqa: E731  (sequencer events that start within the render)
for name, fn in (("saw_svf", workloads.saw_svf_voice), ("saw_svf_events", early_events), ("fm", workloads.fm_voice)):
    sb = ShardedBank(fn, total, device=0, sample_rate=48000.0)   # the mock has one "device"; on a GPU box the default is LOCAL_RANK
    mix = sb.render_mix(n)
    if dist.get_rank() == 0:
        from oracle import oracle_bank_render
        ref, _ = oracle_bank_render([fn(i) for i in range(total)], 48000.0, n, threads=2)
        want = ref.astype(np.float64).sum(0)
        tol = 1e-5 * np.maximum(np.abs(want), 1e-2 * np.abs(want).max())
        good = bool(np.all(np.abs(mix - want) <= tol)) and np.abs(want).max() > 0.1
        print(f"{name}: shards {sb.first}+{sb.count} of {total}: {'ok' if good else 'MISMATCH'}")
        ok = ok and good
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
