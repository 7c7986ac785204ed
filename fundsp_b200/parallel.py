"""Multi-GPU plumbing: one process per GPU, voices sharded contiguously, ONE collective per render — the
sum of the per-GPU partial mixes. The voice kernels never communicate; `torch.distributed`
(NCCL over NVLink on the GPU box, gloo in the CPU tests) only carries the final stereo/mono mix-down.
"""
from __future__ import annotations


def shard_range(total_voices: int, rank: int, world: int):
    """Contiguous voice range [first, first + count) owned by `rank` (reference order: voice index order)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    first = total_voices * rank // world
    last = total_voices * (rank + 1) // world
    return first, last - first


def reduce_mix(mix, dst=0, group=None):
    """Sum the per-rank partial mixes [channels, n] into rank `dst` (in place). Returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(mix, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return mix


class ShardedBank:
    """A bank of `total_voices` voices spread over the ranks of the default process group.

    `builder(i)` returns the `An` expression of global voice i. Every rank renders its shard on its own GPU;
    `render_mix` returns the summed mix on rank `dst` (other ranks get their partial)."""

    def __init__(self, builder, total_voices, device=None, sample_rate=48000.0, bank_factory=None):
        import torch.distributed as dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.first, self.count = shard_range(total_voices, self.rank, self.world)
        voices = [builder(self.first + i) for i in range(self.count)]
        if bank_factory is None:
            from .bank import GpuBank
            bank_factory = lambda v: GpuBank(v, device=device if device is not None else 0, per_voice=False, mix=True, sample_rate=sample_rate)  # noqa: E731
        self.bank = bank_factory(voices)

    def render_mix(self, n, inp=None, dst=0):
        import torch
        _, mix = self.bank.render_samples(n, inp)
        t = torch.from_numpy(mix)
        if torch.cuda.is_available() and self.world > 1 and torch.distributed.get_backend() == "nccl":
            t = t.cuda()
        reduce_mix(t, dst)
        return t.cpu().numpy()
