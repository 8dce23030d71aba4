"""Mel front-end on the device -- the step BEFORE the sampling path: waveform -> the 80-bin log10-mel the vocoder is
conditioned on, i.e. the reference's ``process_utterance`` (data_gen/tts/data_gen_utils.py:93-147, reached through
``BaseVocoder.wav2spec`` vocoders/base_vocoder.py:32-40 and demo.ipynb cell 2) with the base.yaml parameters
(fft 1024, hop 256, win 1024, hann, 80 mels, fmin 80, fmax 7600, eps 1e-6, 22.05 kHz).

The reference delegates the arithmetic to librosa 0.8.0 (requirements.txt:2; not installable here).  This module builds the
``librosa.filters.mel`` table on the host (float64 -> float32, Slaney scale and area normalisation) and the C-ABI entry
``fd_mel_frontend`` does framing, window, FFT, magnitude, filterbank and log10 on the GPU (csrc/fd_kernels_simt.cuh: k_mel_frontend).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from .engine import Engine

SAMPLE_RATE, FFT_SIZE, HOP_SIZE, NUM_MELS, FMIN, FMAX = 22050, 1024, 256, 80, 80.0, 7600.0

_F_SP = 200.0 / 3.0                 # Hz per mel in the linear region (Slaney's Auditory Toolbox scale, librosa's default)
_BREAK_HZ = 1000.0                  # the scale turns logarithmic here ...
_LOGSTEP = np.log(6.4) / 27.0       # ... with 27 mels per factor of 6.4


def _slaney_hz(mel: np.ndarray) -> np.ndarray:
    brk = _BREAK_HZ / _F_SP
    lin = _F_SP * mel
    return np.where(mel >= brk, _BREAK_HZ * np.exp(_LOGSTEP * (mel - brk)), lin)


def _slaney_mel(hz: float) -> float:
    return _BREAK_HZ / _F_SP + np.log(hz / _BREAK_HZ) / _LOGSTEP if hz >= _BREAK_HZ else hz / _F_SP


def mel_filter_table(sr: int = SAMPLE_RATE, n_fft: int = FFT_SIZE, n_mels: int = NUM_MELS, fmin: float = FMIN,
                     fmax: float = FMAX) -> Tuple[np.ndarray, np.ndarray]:
    """-> (table (n_mels, 1 + n_fft//2) float32 == librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax), ranges (n_mels, 2) int32)."""
    edges = _slaney_hz(np.linspace(_slaney_mel(fmin), _slaney_mel(fmax), n_mels + 2))      # band edges in Hz, float64
    freqs = np.arange(1 + n_fft // 2, dtype=np.float64) * (float(sr) / n_fft)              # centre of every FFT bin
    up = (freqs[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]         # rising ramp of triangle m
    down = (edges[2:, None] - freqs[None, :]) / (edges[2:] - edges[1:-1])[:, None]         # falling ramp
    tri = np.clip(np.minimum(up, down), 0.0, None) * (2.0 / (edges[2:] - edges[:-2]))[:, None]   # unit-area triangles
    table = tri.astype(np.float32)
    ranges = np.zeros((n_mels, 2), dtype=np.int32)
    for m in range(n_mels):
        nz = np.flatnonzero(table[m])
        if nz.size:
            ranges[m] = (nz[0], nz[-1] + 1)
    return table, ranges


_tables: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}


def _device_tables(device: torch.device):
    key = str(device)
    if key not in _tables:
        t, r = mel_filter_table()
        _tables[key] = (torch.from_numpy(t).to(device).contiguous(), torch.from_numpy(r).to(device).contiguous())
    return _tables[key]


def wav2mel(engine: Engine, wav: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """process_utterance on the engine's device.  wav (n,) or (B, n) fp32 in [-1, 1] at 22.05 kHz ->
    (wav zero-padded/cropped to T'*256 samples (B, T'*256), mel (B, 80, T')), T' = 1 + n // 256 -- the pair the reference returns
    (mel already in the (B, 80, T') layout ``FastDiff.forward`` / ``sampling_given_noise_schedule`` take as ``condition``)."""
    if wav.dim() == 1:
        wav = wav.unsqueeze(0)
    B, n = wav.shape
    dev = engine.device
    w = wav.to(dev, torch.float32).contiguous()
    Tm = 1 + n // HOP_SIZE
    mel = torch.empty((B, NUM_MELS, Tm), dtype=torch.float32, device=dev)
    fb, rng = _device_tables(dev)
    rc = engine.lib.fd_mel_frontend(engine.h, w.data_ptr(), B, n, fb.data_ptr(), rng.data_ptr(), mel.data_ptr(), engine._stream())
    _lib.check(engine.lib, engine.h, rc, "fd_mel_frontend")
    out = torch.zeros((B, Tm * HOP_SIZE), dtype=torch.float32, device=dev)
    out[:, : min(n, Tm * HOP_SIZE)] = w[:, : Tm * HOP_SIZE]
    return out, mel
