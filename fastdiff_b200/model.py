"""``FastDiff`` -- drop-in for /root/reference/modules/FastDiff/module/FastDiff_model.py:10-122.

Same constructor kwargs, same parameter names and shapes (so ``load_state_dict(ckpt["state_dict"]["model"])``
and ``Trainer.restore_weights`` work unchanged), same ``forward((audio, c, diffusion_steps))`` contract.
The module tree below only HOLDS the parameters; ``forward`` runs the hand-written sm_100a kernels through
the C-ABI (include/fastdiff_b200.h).  There is no PyTorch fallback: on a box without the CUDA extension or
without a GPU, ``forward`` raises.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .engine import DEFAULT_ARCH, Engine
from .weights import pack_state_dict


class _DBlockParams(nn.Module):
    """Parameter holder mirroring DiffusionDBlock (modules.py:116-126)."""

    def __init__(self, cin, hidden, factor):
        super().__init__()
        self.factor = factor
        self.residual_dense = nn.Conv1d(cin, hidden, 1)
        self.conv = nn.ModuleList([
            nn.Conv1d(cin, hidden, 3, dilation=1, padding=1),
            nn.Conv1d(hidden, hidden, 3, dilation=2, padding=2),
            nn.Conv1d(hidden, hidden, 3, dilation=4, padding=4),
        ])


class _KernelPredictorParams(nn.Module):
    """Parameter holder mirroring KernelPredictor (modules.py:257-318); Sequential slot numbers are kept so the
    state-dict keys (`residual_conv.{1,3,6,8,11,13}`) match."""

    def __init__(self, cond, cin, cout, layers, ksz, hid, kpk, dropout):
        super().__init__()
        pad = (kpk - 1) // 2
        act = lambda: nn.LeakyReLU(negative_slope=0.1)
        self.input_conv = nn.Sequential(nn.Conv1d(cond, hid, 5, padding=2, bias=True), act())
        self.residual_conv = nn.Sequential(
            nn.Dropout(dropout), nn.Conv1d(hid, hid, kpk, padding=pad), act(), nn.Conv1d(hid, hid, kpk, padding=pad), act(),
            nn.Dropout(dropout), nn.Conv1d(hid, hid, kpk, padding=pad), act(), nn.Conv1d(hid, hid, kpk, padding=pad), act(),
            nn.Dropout(dropout), nn.Conv1d(hid, hid, kpk, padding=pad), act(), nn.Conv1d(hid, hid, kpk, padding=pad), act(),
        )
        self.kernel_conv = nn.Conv1d(hid, cin * cout * ksz * layers, kpk, padding=pad)
        self.bias_conv = nn.Conv1d(hid, cout * layers, kpk, padding=pad)


class _LVCBlockParams(nn.Module):
    """Parameter holder mirroring TimeAware_LVCBlock (modules.py:141-187)."""

    def __init__(self, cin, cond, ratio, layers, ksz, hop, hid, kpk, dropout, emb_out):
        super().__init__()
        self.cond_hop_length = hop
        self.convs = nn.ModuleList()
        self.upsample = nn.ConvTranspose1d(cin, cin, kernel_size=ratio * 2, stride=ratio,
                                           padding=ratio // 2 + ratio % 2, output_padding=ratio % 2)
        self.kernel_predictor = _KernelPredictorParams(cond, cin, 2 * cin, layers, ksz, hid, kpk, dropout)
        self.fc_t = nn.Linear(emb_out, cond)
        for i in range(layers):
            self.convs.append(nn.Conv1d(cin, cin, kernel_size=ksz, padding=(3 ** i) * ((ksz - 1) // 2), dilation=3 ** i))


class FastDiff(nn.Module):
    """FastDiff denoiser eps_theta(x_t, mel, t) on B200."""

    def __init__(self, audio_channels=1, inner_channels=32, cond_channels=80, upsample_ratios=[8, 8, 4],
                 lvc_layers_each_block=4, lvc_kernel_size=3, kpnet_hidden_channels=64, kpnet_conv_size=3, dropout=0.0,
                 diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512, diffusion_step_embed_dim_out=512,
                 use_weight_norm=True):
        super().__init__()
        self.arch = dict(
            audio_channels=audio_channels, inner_channels=inner_channels, cond_channels=cond_channels,
            upsample_ratios=list(upsample_ratios), lvc_layers_each_block=lvc_layers_each_block,
            lvc_kernel_size=lvc_kernel_size, kpnet_hidden_channels=kpnet_hidden_channels, kpnet_conv_size=kpnet_conv_size,
            dropout=dropout, diffusion_step_embed_dim_in=diffusion_step_embed_dim_in,
            diffusion_step_embed_dim_mid=diffusion_step_embed_dim_mid,
            diffusion_step_embed_dim_out=diffusion_step_embed_dim_out, use_weight_norm=use_weight_norm)
        if float(dropout) != 0.0:
            raise NotImplementedError("fastdiff_b200 implements the inference path; dropout must be 0.0 (base.yaml:30)")
        self.diffusion_step_embed_dim_in = diffusion_step_embed_dim_in
        self.audio_channels = audio_channels
        self.cond_channels = cond_channels
        self.lvc_block_nums = len(upsample_ratios)
        self.first_audio_conv = nn.Conv1d(1, inner_channels, kernel_size=7, padding=3, dilation=1, bias=True)
        self.lvc_blocks = nn.ModuleList()
        self.downsample = nn.ModuleList()
        self.fc_t = nn.ModuleList()  # empty in the reference too (FastDiff_model.py:43)
        self.fc_t1 = nn.Linear(diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid)
        self.fc_t2 = nn.Linear(diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out)
        hop = 1
        for n in range(self.lvc_block_nums):
            hop *= upsample_ratios[n]
            self.lvc_blocks.append(_LVCBlockParams(inner_channels, cond_channels, upsample_ratios[n], lvc_layers_each_block,
                                                   lvc_kernel_size, hop, kpnet_hidden_channels, kpnet_conv_size, dropout,
                                                   diffusion_step_embed_dim_out))
            self.downsample.append(_DBlockParams(inner_channels, inner_channels, upsample_ratios[self.lvc_block_nums - n - 1]))
        self.final_conv = nn.Sequential(nn.Conv1d(inner_channels, audio_channels, kernel_size=7, padding=3, dilation=1, bias=True))
        if use_weight_norm:
            self.apply_weight_norm()
        # device-side state
        self._engine: Optional[Engine] = None
        self._packed_version = None
        self.mode = None          # None -> library default (tc_3xf16); or "fp32_simt" | "tc_3xtf32" | "tc_tf32" | "tc_3xf16"
        self.noise_mode = "reference"  # sampler: "reference" (CPU generator, reference order) | "device" (Philox)
        self.seed = 0
        self._lib_path = None     # tests point this at the CPU emulation build
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_weights())

    # -- reference API -------------------------------------------------------------------------
    def apply_weight_norm(self):
        def _apply(m):
            if isinstance(m, (nn.Conv1d, nn.Conv2d)):
                nn.utils.weight_norm(m)
        self.apply(_apply)

    def remove_weight_norm(self):
        def _remove(m):
            try:
                nn.utils.remove_weight_norm(m)
            except ValueError:
                return
        self.apply(_remove)
        self.invalidate_weights()

    def forward(self, data):
        """data = (audio (B,1,L) fp32, c (B,80,T') fp32, diffusion_steps (B,1) float) -> (B,1,L) fp32."""
        audio, c, diffusion_steps = data
        eng = self.engine(audio.device)
        if c.dim() == 2:
            c = c.unsqueeze(0)
        B = audio.shape[0]
        if c.shape[0] != B:
            c = c.expand(B, -1, -1)
        with torch.no_grad():
            return eng.denoise(audio, c, diffusion_steps.reshape(-1).to(torch.float32))

    def cross_check(self, data, modes=("fp32_simt",)):
        """On-device guard for the tensor-core default mode: evaluates ``forward(data)`` in the current mode and in every mode of
        ``modes`` (default: the strict FFMA path) and returns ``{mode: max|eps - eps_current|}``.  The fp16-piece mode ``tc_3xf16``
        saturates operands at |v*S| = 65504 (|activation| >= 4094, |predicted kernel| >= 1023): a checkpoint whose activations
        leave that range shows up here as a large difference (and should run in ``tc_3xtf32``).  Not part of the reference API."""
        keep = self.mode
        eng = self.engine(data[0].device)
        prev = eng.get_mode()
        try:
            ref = self.forward(data)
            out = {}
            for m in modes:
                self.mode = m
                out[m] = float((self.forward(data) - ref).abs().max().item())
            return out
        finally:
            self.mode = keep
            eng.set_mode(prev)

    # -- engine management ---------------------------------------------------------------------
    def invalidate_weights(self):
        """Force a re-pack of the weight blob on the next call.  REQUIRED after edits that autograd cannot see -- ``p.data.copy_()``,
        EMA swaps, manual ``.data`` assignment: they keep the storage pointer and do not bump ``p._version``, so `_weights_version`
        cannot notice them.  ``load_state_dict``, ``remove_weight_norm``, ``.to()`` and in-place ops on the parameters (optimizer steps
        included) are detected automatically."""
        self._packed_version = None

    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device=None) -> Engine:
        """The C-ABI handle for ``device`` with current weights loaded (re-packed when parameters changed)."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda" and self._lib_path is None:
            raise RuntimeError("fastdiff_b200.FastDiff runs on CUDA (sm_100a) only; move the module and inputs to a GPU "
                               "(there is no CPU fallback for the product path)")
        if self._engine is None or self._engine.device != device:
            self._engine = Engine(self.arch, device=device, lib_path=self._lib_path)
            self._packed_version = None
        ver = self._weights_version()
        if self._packed_version != ver:
            self._engine.load_blob(pack_state_dict(self.state_dict()))
            self._packed_version = ver
        if self.mode is not None:
            self._engine.set_mode(self.mode)
        return self._engine
