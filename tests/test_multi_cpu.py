"""N > 1 host logic on CPU: world_size-2 gloo process group, voice sharding + the single mix-down reduce.
The per-rank 'bank' is the CPU oracle here (no GPU in this container); on the GPU box the same ShardedBank
wraps GpuBank and NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

from fundsp_b200.parallel import shard_range


def test_shard_ranges_partition_the_voices():
    for total in (1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                f, c = shard_range(total, r, world)
                cover += list(range(f, f + c))
            assert cover == list(range(total))
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, n, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch.distributed as dist
    from fundsp_b200 import workloads
    from fundsp_b200.parallel import ShardedBank
    from oracle import oracle_bank_render

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class OracleBank:  # CPU stand-in with the GpuBank.render_samples signature
        def __init__(self, voices):
            self.voices = voices

        def render_samples(self, n, inp=None):
            return oracle_bank_render(self.voices, 48000.0, n, inp, per_voice=False, mix=True, threads=1)

    sb = ShardedBank(workloads.noise_svf_voice, total, bank_factory=OracleBank)
    mix = sb.render_mix(n)
    q.put((rank, sb.first, sb.count, mix))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_mixdown_matches_single_process():
    import torch.multiprocessing as mp
    from fundsp_b200 import workloads
    from oracle import oracle_bank_render

    total, n, world = 37, 512, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 18), (18, 19)]
    per_voice, _ = oracle_bank_render([workloads.noise_svf_voice(i) for i in range(total)], 48000.0, n, per_voice=True, mix=False)
    ref = per_voice.astype(np.float64).sum(axis=0)
    got = res[0][3]  # rank 0 holds the reduced mix
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(per_voice).sum(axis=0).max()
