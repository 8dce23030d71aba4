"""Time-axis shard mode: ONE long utterance (or a small batch of them) split along time over the GPUs.

Each rank owns a contiguous range of mel frames and runs the ordinary single-GPU sampler step on that range extended by a halo of
HALO_FRAMES frames on either side (the receptive field of one denoiser evaluation is +-4,137 samples = 16.2 frames, measured on the
reference; the halo is 17).  Inside the halo the step's output is wrong near the outer edge -- so after EVERY reverse step the
ranks exchange exact boundary samples with their neighbours (one send + one receive per side over NVLink: 17*256 floats per item),
which makes the sharded result equal to the unsharded one for any number of steps.  This is the only mode with a per-step exchange;
batch-shard mode (shard.py) needs none.  The reference has no counterpart (its test_step is batch-1 on one GPU).

Noise, two modes.  "reference": the reference's RNG stream (CPU default generator, x_T then one draw per noisy step, util.py:216-234) is
drawn at full size on every rank from the same seed and sliced, so the result equals the single-GPU parity-mode result (a correctness
mode: every rank pays the full-size host draws).  "device" (the LATENCY mode): every rank draws only its own window on the GPU -- the
Philox element index of a sample is its position in the WHOLE utterance (fd_set_noise_window) and the draw number that of the
single-call loop (option noise_draw_base) -- so the sharded result is bit-identical to the single-GPU device-noise result with the
same seed, and nothing but the halo samples crosses NVLink.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .engine import Engine
from .sampler import build_steps
from .shard import shard_range

HALO_FRAMES = 17
HOP = 256


def frame_ranges(Tm: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(Tm, world, r) for r in range(world)]


class TimeShardedSampler:
    """engine: a loaded fastdiff_b200.engine.Engine on this rank's device (e.g. ShardedFastDiff(...).engine: weights arrive by one
    broadcast).  All ranks call sample() with the same arguments."""

    def __init__(self, engine: Engine, group=None, halo_frames: int = HALO_FRAMES):
        self.engine, self.group, self.halo = engine, group, int(halo_frames)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = engine.device

    def sample(self, size, diffusion_hyperparams, schedule, condition, ddim: bool = False, gather: bool = True, noise: str = "reference",
               seed: int = 0) -> torch.Tensor:
        """size = (B,1,L), condition (B,80,T') on any device, L = 256*T'.  Returns the full (B,1,L) waveform on every rank
        (gather=True) or this rank's interior (B,1,256*(f_hi-f_lo)).  noise: "reference" | "device" (see the module docstring)."""
        B, _, L = size
        Tm = condition.shape[-1]
        if Tm * HOP != L:
            raise AssertionError("length of (x, kernel) is not matched")  # modules.py:236
        H = self.halo
        ranges = frame_ranges(Tm, self.world)
        if self.world > 1 and min(hi - lo for lo, hi in ranges) < H:
            raise ValueError(f"time shard needs at least {H} frames per rank (T'={Tm}, world={self.world})")
        f_lo, f_hi = ranges[self.rank]
        hl, hr = min(H, f_lo), min(H, Tm - f_hi)          # halo actually present (none at the utterance ends)
        e_lo, e_hi = f_lo - hl, f_hi + hr
        _, steps = build_steps(diffusion_hyperparams, schedule, ddim)
        mel = condition[:, :, e_lo:e_hi].to(self.device, torch.float32).contiguous()
        sl = slice(e_lo * HOP, e_hi * HOP)
        if noise == "device":
            x = torch.empty((B, 1, (e_hi - e_lo) * HOP), dtype=torch.float32, device=self.device)
            self.engine.set_noise_window(L, e_lo * HOP)          # my samples are [e_lo*256, e_hi*256) of utterances of L samples
            try:
                with torch.no_grad():
                    draws = 0
                    for i, st in enumerate(steps):
                        self.engine.set_option("noise_draw_base", draws)
                        self.engine.sample(x, mel, [st], noise=None, seed=seed, fill_xT=(i == 0), ddim=ddim)
                        draws += 1 if (st.add_noise and not ddim) else 0
                        self._exchange(x, hl, hr)
            finally:
                self.engine.set_option("noise_draw_base", 0)
                self.engine.set_noise_window(0, 0)
        else:
            x = torch.normal(0, 1, size=size)[:, :, sl].to(self.device).contiguous()        # x_T: same CPU draw on every rank
            with torch.no_grad():
                for st in steps:
                    z = None
                    if st.add_noise and not ddim:
                        z = torch.normal(0, 1, size=size)[:, :, sl].to(self.device).contiguous().unsqueeze(0)
                    self.engine.sample(x, mel, [st], noise=z, ddim=ddim)
                    self._exchange(x, hl, hr)
        interior = x[:, :, hl * HOP: hl * HOP + (f_hi - f_lo) * HOP].contiguous()
        return self._gather(interior, B, ranges) if gather else interior

    # -- one neighbour exchange: my boundary interior samples go out, exact halo samples come in ---------------------------
    def _exchange(self, x: torch.Tensor, hl: int, hr: int):
        if self.world == 1:
            return
        n = self.halo * HOP
        ops, recv = [], []
        left, right = self.rank - 1, self.rank + 1
        lo_i = hl * HOP                                    # start of my interior
        hi_i = x.shape[-1] - hr * HOP                      # end of my interior
        if left >= 0:
            out_l = x[:, :, lo_i: lo_i + n].contiguous()
            in_l = torch.empty_like(out_l)
            ops += [dist.P2POp(dist.isend, out_l, self._peer(left), self.group), dist.P2POp(dist.irecv, in_l, self._peer(left), self.group)]
            recv.append((slice(lo_i - n, lo_i), in_l))
        if right < self.world:
            out_r = x[:, :, hi_i - n: hi_i].contiguous()
            in_r = torch.empty_like(out_r)
            ops += [dist.P2POp(dist.isend, out_r, self._peer(right), self.group), dist.P2POp(dist.irecv, in_r, self._peer(right), self.group)]
            recv.append((slice(hi_i, hi_i + n), in_r))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for s, buf in recv:
            x[:, :, s] = buf

    def _peer(self, group_rank: int) -> int:
        return dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    def _gather(self, part: torch.Tensor, B: int, ranges) -> torch.Tensor:
        if self.world == 1:
            return part
        mx = max(hi - lo for lo, hi in ranges) * HOP
        pad = torch.zeros((B, 1, mx), dtype=part.dtype, device=part.device)
        pad[:, :, : part.shape[-1]] = part
        bufs = [torch.empty_like(pad) for _ in ranges]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([b[:, :, : (hi - lo) * HOP] for b, (lo, hi) in zip(bufs, ranges)], dim=-1)
