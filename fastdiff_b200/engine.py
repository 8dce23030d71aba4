"""Device-side engine: one C-ABI handle + workspace cache.  Owns no tensors of the caller; all device
memory is torch-allocated and passed down as raw pointers on torch's current stream."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import fd_config, fd_step

DEFAULT_ARCH = dict(
    audio_channels=1, inner_channels=32, cond_channels=80, upsample_ratios=[8, 8, 4], lvc_layers_each_block=4,
    lvc_kernel_size=3, kpnet_hidden_channels=64, kpnet_conv_size=3, dropout=0.0, diffusion_step_embed_dim_in=128,
    diffusion_step_embed_dim_mid=512, diffusion_step_embed_dim_out=512, use_weight_norm=True,
)


def _cfg_struct(arch: dict) -> fd_config:
    r = list(arch["upsample_ratios"])
    if len(r) > 4:
        raise ValueError("at most 4 upsample ratios")
    cfg = fd_config()
    cfg.audio_channels = arch["audio_channels"]
    cfg.inner_channels = arch["inner_channels"]
    cfg.cond_channels = arch["cond_channels"]
    cfg.n_upsample = len(r)
    for i, v in enumerate(r):
        cfg.upsample_ratios[i] = int(v)
    cfg.lvc_layers_each_block = arch["lvc_layers_each_block"]
    cfg.lvc_kernel_size = arch["lvc_kernel_size"]
    cfg.kpnet_hidden_channels = arch["kpnet_hidden_channels"]
    cfg.kpnet_conv_size = arch["kpnet_conv_size"]
    cfg.diffusion_step_embed_dim_in = arch["diffusion_step_embed_dim_in"]
    cfg.diffusion_step_embed_dim_mid = arch["diffusion_step_embed_dim_mid"]
    cfg.diffusion_step_embed_dim_out = arch["diffusion_step_embed_dim_out"]
    return cfg


class Engine:
    """Thin owner of an ``fd_handle``.  ``device`` is a torch device; in the CPU emulation build used by the
    test-suite (lib_path given, device cpu) "device pointers" are host pointers."""

    def __init__(self, arch: Optional[dict] = None, device="cuda:0", lib_path: Optional[str] = None):
        self.lib = _lib.load(lib_path)
        self.device = torch.device(device)
        self.h = C.c_void_p()
        cfg = _cfg_struct({**DEFAULT_ARCH, **(arch or {})})
        idx = self.device.index if self.device.type == "cuda" else 0
        rc = self.lib.fd_create(C.byref(cfg), int(idx or 0), C.byref(self.h))
        if rc != 0:
            raise _lib.FdError(f"fd_create failed ({rc}): {self.lib.fd_last_error(None).decode()}")
        self._ws: Dict[tuple, torch.Tensor] = {}
        self.loaded = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.fd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ------------------------------------------------------------------------------
    def _stream(self):
        if self.device.type == "cuda":
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return C.c_void_p(0)

    def _check(self, rc, what):
        _lib.check(self.lib, self.h, rc, what)

    def _dev(self, t: torch.Tensor, shape=None) -> torch.Tensor:
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(self.device, torch.float32).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def workspace(self, B: int, Tm: int) -> torch.Tensor:
        key = (B, Tm)
        ws = self._ws.get(key)
        if ws is None:
            n = C.c_size_t()
            self._check(self.lib.fd_workspace_bytes(self.h, B, Tm, C.byref(n)), "fd_workspace_bytes")
            self._ws.clear()  # one live shape at a time keeps HBM use bounded
            ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # -- API -----------------------------------------------------------------------------------
    def load_blob(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob)
        self._check(self.lib.fd_load_weights(self.h, C.c_void_p(blob.ctypes.data), blob.nbytes), "fd_load_weights")
        self.loaded = True

    def load_blob_device(self, blob: torch.Tensor):
        assert blob.dtype == torch.uint8 and blob.is_contiguous()
        self._check(self.lib.fd_load_weights_dev(self.h, C.c_void_p(blob.data_ptr()), blob.numel(), self._stream()),
                    "fd_load_weights_dev")
        self.loaded = True

    def set_mode(self, mode):
        m = _lib.MODE_NAMES[mode] if isinstance(mode, str) else int(mode)
        self._check(self.lib.fd_set_mode(self.h, m), "fd_set_mode")

    def get_mode(self) -> int:
        return self.lib.fd_get_mode(self.h)

    def set_option(self, key: str, value: int):
        self._check(self.lib.fd_set_option(self.h, key.encode(), int(value)), "fd_set_option")

    def set_noise_window(self, total_samples: int = 0, offset: int = 0):
        """Device-noise mode: this engine's tensors are the window [offset, offset + L) of utterances of `total_samples` samples (time shard)."""
        self._check(self.lib.fd_set_noise_window(self.h, int(total_samples), int(offset)), "fd_set_noise_window")

    def launch_count(self) -> int:
        return int(self.lib.fd_launch_count(self.h))

    def check_saturation(self, reset: bool = True) -> bool:
        """True if an fp16 operand piece saturated in a tensor-core kernel since the last reset (mode tc_3xf16 only; synchronises)."""
        flag = C.c_int(0)
        self._check(self.lib.fd_check_saturation(self.h, C.byref(flag), int(reset), self._stream()), "fd_check_saturation")
        return bool(flag.value)

    def timing_enable(self, on: bool = True):
        self._check(self.lib.fd_timing_enable(self.h, int(on)), "fd_timing_enable")

    def timing_report(self) -> dict:
        """Per-kernel-class device time since the last report; synchronises the device first."""
        import json
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        buf = C.create_string_buffer(4096)
        self._check(self.lib.fd_timing_report(self.h, buf, 4096), "fd_timing_report")
        return json.loads(buf.value.decode())

    def denoise(self, x: torch.Tensor, mel: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        B, Tm = mel.shape[0], mel.shape[2]
        L = Tm * 256
        x = self._dev(x, (B, 1, L))
        mel = self._dev(mel, (B, 80, Tm))
        t = self._dev(t.reshape(-1), (B,))
        out = torch.empty_like(x)
        ws = self.workspace(B, Tm)
        rc = self.lib.fd_denoise(self.h, x.data_ptr(), mel.data_ptr(), t.data_ptr(), out.data_ptr(), B, Tm,
                                 ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, "fd_denoise")
        return out

    def sample(self, x: torch.Tensor, mel: torch.Tensor, steps: Sequence[fd_step], noise: Optional[torch.Tensor] = None,
               seed: int = 0, fill_xT: bool = False, ddim: bool = False, seq: Optional[torch.Tensor] = None) -> torch.Tensor:
        """In place on ``x`` ((B,1,L) fp32 on the engine's device)."""
        B, Tm = mel.shape[0], mel.shape[2]
        L = Tm * 256
        assert x.device == self.device and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (B, 1, L)
        mel = self._dev(mel, (B, 80, Tm))
        arr = (fd_step * max(len(steps), 1))(*steps)
        n_noise = 0
        nptr = None
        if noise is not None:
            noise = self._dev(noise)
            n_noise = noise.shape[0]
            assert tuple(noise.shape[1:]) == (B, 1, L)
            nptr = noise.data_ptr()
        sptr = None
        if seq is not None:
            assert seq.device == self.device and seq.is_contiguous() and tuple(seq.shape) == (len(steps) + 1, B, 1, L)
            sptr = seq.data_ptr()
        ws = self.workspace(B, Tm)
        rc = self.lib.fd_sample(self.h, x.data_ptr(), mel.data_ptr(), arr, len(steps), nptr, n_noise, seed, int(fill_xT),
                                int(ddim), sptr, B, Tm, ws.data_ptr(), ws.numel(), self._stream())
        self._check(rc, "fd_sample")
        return x

    def reverse_update(self, x: torch.Tensor, eps: torch.Tensor, step: fd_step, z: Optional[torch.Tensor] = None, ddim: bool = False,
                       seed: int = 0, draw: int = 0, seq: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One reverse-step update in place on ``x`` (util.py:219-229); ``eps`` = the denoiser's output for this step."""
        assert x.device == self.device and x.dtype == torch.float32 and x.is_contiguous()
        eps = self._dev(eps, x.shape)
        zp = None
        if z is not None:
            z = self._dev(z, x.shape)
            zp = z.data_ptr()
        sp = None
        if seq is not None:
            assert seq.device == self.device and seq.dtype == torch.float32 and seq.is_contiguous() and seq.shape == x.shape
            sp = seq.data_ptr()
        rc = self.lib.fd_reverse_update(self.h, x.data_ptr(), eps.data_ptr(), zp, C.byref(step), int(ddim), int(seed), int(draw), sp,
                                        x.numel(), self._stream())
        self._check(rc, "fd_reverse_update")
        return x

    def wav_int16(self, x: torch.Tensor) -> torch.Tensor:
        """(B,1,L) fp32 waveform -> (B,L) int16: wav/|wav|.max() then *32767 and truncate (task/FastDiff.py:110, utils/audio.py:11-16)."""
        B, L = x.shape[0], x.shape[-1]
        x = self._dev(x.reshape(B, 1, L))
        out = torch.empty((B, L), dtype=torch.int16, device=self.device)
        ws = self.workspace(B, max(1, L // 256))
        self._check(self.lib.fd_wav_int16(self.h, x.data_ptr(), out.data_ptr(), B, L, ws.data_ptr(), self._stream()), "fd_wav_int16")
        return out

    def debug_read(self, name: str, B: int, Tm: int) -> torch.Tensor:
        ws = self.workspace(B, Tm)
        n = C.c_size_t()
        self._check(self.lib.fd_debug_read(self.h, name.encode(), None, C.byref(n), B, Tm, ws.data_ptr(), self._stream()),
                    "fd_debug_read")
        out = torch.empty(n.value, dtype=torch.float32, device=self.device)
        self._check(self.lib.fd_debug_read(self.h, name.encode(), out.data_ptr(), C.byref(n), B, Tm, ws.data_ptr(),
                                           self._stream()), "fd_debug_read")
        return out
