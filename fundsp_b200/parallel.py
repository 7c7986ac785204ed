"""Multi-GPU plumbing: one process per GPU, voices sharded contiguously, ONE exchange step per render — the
sum of the per-GPU partial mixes. The voice kernels never communicate.

The exchange itself lives BELOW the C ABI (`fdsp_group_*`, `fdsp_bank_render_reduced`: NCCL send/recv gather over NVLink on the
bank's stream + a rank-order fold on the root, csrc/host/group.cpp) — `BankGroup` here is its ctypes wrapper, and the only thing a
host has to carry between ranks is the 128-byte unique id (here: through whatever `torch.distributed` group is up, or a file).
`ShardedBank(..., group=...)` uses it; without a group it falls back to a `torch.distributed` reduce of the host mix (the gloo
CPU tests, where the per-rank bank is the oracle or the mock device).
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


class BankGroup:
    """`fdsp_group`: the ranks whose banks are mixed down together (one NCCL communicator)."""

    def __init__(self, nranks, rank, unique_id, device):
        import ctypes as C
        from . import capi
        self.L = capi.lib()
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        capi.check(self.L.fdsp_group_create(int(nranks), int(rank), buf, int(device), C.byref(h)))
        self.h, self.nranks, self.rank, self.device = h, nranks, rank, device

    @staticmethod
    def unique_id():
        import ctypes as C
        from . import capi
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().fdsp_group_unique_id(buf, 128))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device=None):
        """Rank 0 draws the id, the default process group (any backend) carries it to the others."""
        import os
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 and world > 1 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        dev = int(os.environ.get("LOCAL_RANK", rank)) if device is None else device
        return cls(world, rank, box[0], dev)

    def render_reduced(self, bank, n, inp=None, root=0):
        """`fdsp_bank_render_reduced`: this rank's shard rendered and mixed down across the group; the root gets [channels, n]."""
        import ctypes as C
        import numpy as np
        from . import capi
        fp = C.POINTER(C.c_float)
        x = None if inp is None else np.ascontiguousarray(inp, np.float32)
        out = np.zeros((bank.voice_outputs(), n), np.float32) if self.rank == root else None
        capi.check(self.L.fdsp_bank_render_reduced(bank.h, self.h, int(n), None if x is None else x.ctypes.data_as(fp), None if out is None else out.ctypes.data_as(fp), int(root)))
        return out

    def reduce_device(self, bank, n, mix_ptr, mix_stride, root=0):
        from . import capi
        capi.check(self.L.fdsp_bank_reduce_device(bank.h, self.h, int(n), int(mix_ptr), int(mix_stride), int(root)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.fdsp_group_destroy(self.h)
                self.h = None
        except Exception:
            pass


class ShardedBank:
    """A bank of `total_voices` voices spread over the ranks of the default process group.

    `builder(i)` returns the `An` expression of global voice i. Every rank renders its shard on its own GPU;
    `render_mix` returns the summed mix on rank `dst` (other ranks get their partial)."""

    def __init__(self, builder, total_voices, device=None, sample_rate=48000.0, bank_factory=None, group=None):
        import os
        import torch.distributed as dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if total_voices < self.world:
            raise ValueError(f"{total_voices} voices cannot be sharded over {self.world} ranks: a rank would own no voice")
        self.first, self.count = shard_range(total_voices, self.rank, self.world)
        voices = [builder(self.first + i) for i in range(self.count)]
        # one GPU per rank: LOCAL_RANK (torchrun) unless told otherwise — never every rank on device 0
        self.device = int(os.environ.get("LOCAL_RANK", self.rank)) if device is None else device
        self.group = group
        if bank_factory is None:
            from .bank import GpuBank
            bank_factory = lambda v: GpuBank(v, device=self.device, per_voice=False, mix=True, sample_rate=sample_rate)  # noqa: E731
        self.bank = bank_factory(voices)

    def render_mix(self, n, inp=None, dst=0):
        if self.group is not None:       # the C-ABI path: NCCL gather + rank-order fold on the bank's stream, one D2H on the root
            out = self.group.render_reduced(self.bank, n, inp, dst)
            return out if out is not None else None
        import torch
        _, mix = self.bank.render_samples(n, inp)
        t = torch.from_numpy(mix)
        if torch.cuda.is_available() and self.world > 1 and torch.distributed.get_backend() == "nccl":
            torch.cuda.set_device(self.device)
            t = t.cuda(self.device)
        reduce_mix(t, dst)
        return t.cpu().numpy()
