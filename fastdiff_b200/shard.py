"""Batch-shard mode: utterances are independent (no cross-item op anywhere in FastDiff.forward), so N GPUs
each sample a contiguous slice of the batch.  One process per GPU (torch.distributed, NCCL over NVLink); the ONLY
collective on the path is one broadcast of the packed weight blob at load -- nothing per step.  The reference's
equivalent is DDP's construction-time parameter broadcast + DistributedSampler sharding of the test set
(/root/reference/utils/trainer.py:442-467, tasks/vocoder/vocoder_base.py:37-58)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .engine import Engine
from .sampler import build_steps


def shard_range(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of B items over `world` ranks; the first B % world ranks get one extra item."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedFastDiff:
    """Per-rank sampler over this rank's slice of the batch.

    state_dict is needed on rank 0 only (reference-shaped keys); every other rank receives the packed blob through
    ``dist.broadcast`` into device memory and hands the device pointer to ``fd_load_weights_dev``."""

    def __init__(self, state_dict=None, device=None, group=None, lib_path: Optional[str] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.engine = Engine(device=self.device, lib_path=lib_path)
        if self.rank == 0:
            from .weights import pack_state_dict
            assert state_dict is not None, "rank 0 needs the state dict"
            blob = torch.from_numpy(pack_state_dict(state_dict)).to(self.device)
            n = torch.tensor([blob.numel()], dtype=torch.int64, device=self.device)
        else:
            n = torch.zeros(1, dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.broadcast(n, 0, group=group)
            if self.rank != 0:
                blob = torch.empty(int(n.item()), dtype=torch.uint8, device=self.device)
            dist.broadcast(blob, 0, group=group)
        self.engine.load_blob_device(blob)
        self.blob_bytes = int(n.item())

    def my_slice(self, B: int) -> Tuple[int, int]:
        return shard_range(B, self.world, self.rank)

    def denoise(self, x, mel, t, gather: bool = False):
        """FastDiff.forward on this rank's slice of (x, mel, t) (each given for the WHOLE batch, on any device)."""
        lo, hi = self.my_slice(x.shape[0])
        out = self.engine.denoise(x[lo:hi], mel[lo:hi], t.reshape(-1)[lo:hi]) if hi > lo else x.new_zeros((0,) + tuple(x.shape[1:]), device=self.device)
        return self._gather(out, x.shape[0]) if gather else out

    def sample(self, size, diffusion_hyperparams, schedule, condition, noise=None, seed: int = 0, ddim: bool = False,
               gather: bool = False, presliced: bool = False):
        """Reverse loop on this rank's slice.  size = (B,1,L) of the WHOLE batch; condition = the whole batch's mels, or -- presliced --
        only this rank's.  noise (optional): host tensors [x_T, z_...] for the WHOLE batch in the reference's draw order; each rank
        slices the same indices so sharded == unsharded bitwise.  A rank whose slice is empty (B < world) runs nothing but still
        takes part in the gather."""
        B, _, L = size
        lo, hi = self.my_slice(B)
        _, steps = build_steps(diffusion_hyperparams, schedule, ddim)
        cond = condition if presliced else condition[lo:hi]
        if hi == lo:
            x = torch.empty((0, 1, L), dtype=torch.float32, device=self.device)
        elif noise is not None:
            x = noise[0][lo:hi].to(self.device, torch.float32).contiguous()
            zs = torch.stack([z[lo:hi] for z in noise[1:]]).to(self.device) if len(noise) > 1 else None
            self.engine.sample(x, cond, steps, noise=zs, ddim=ddim)
        else:
            x = torch.empty((hi - lo, 1, L), dtype=torch.float32, device=self.device)
            self.engine.sample(x, cond, steps, noise=None, seed=seed * 65536 + self.rank, fill_xT=True, ddim=ddim)
        return self._gather(x, B) if gather else x

    def _gather(self, part: torch.Tensor, B: int) -> torch.Tensor:
        if self.world == 1:
            return part
        sizes = [shard_range(B, self.world, r) for r in range(self.world)]
        mx = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros((mx,) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
        pad[: part.shape[0]] = part
        bufs = [torch.empty_like(pad) for _ in sizes]
        dist.all_gather(bufs, pad, group=self.group)  # end of the call only -- never inside the reverse loop
        return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
