"""TEST INFRASTRUCTURE -- CPU restatement of the reference's mel front-end (the step BEFORE the sampling path).

Reference call site: /root/reference/data_gen/tts/data_gen_utils.py:93-147 (`process_utterance`), used by
`vocoders/base_vocoder.py` wav2spec and demo.ipynb cell 2:

    x_stft = librosa.stft(wav, n_fft=1024, hop_length=256, win_length=1024, window="hann", pad_mode="constant")
    spc = np.abs(x_stft)
    mel = librosa.filters.mel(22050, 1024, 80, 80, 7600) @ spc
    mel = np.log10(np.maximum(1e-6, mel))
    wav = np.pad(wav, (0, r_pad))[: mel.shape[1] * 256]            # utils/audio.py:67-77 (librosa_pad_lr, pad_sides=1)

The arithmetic lives in **librosa 0.8.0** (pinned by the reference's requirements.txt:2), which is NOT installed in this image
(no network), so this file restates librosa's published algorithm in numpy:
  * `librosa.stft` (core/spectrum.py): centre padding of n_fft//2 samples per side with `pad_mode` (zeros here), periodic Hann
    window (`scipy.signal.get_window("hann", 1024, fftbins=True)`), frames at hop 256, rfft per frame, complex64 result;
  * `librosa.filters.mel` (filters.py): Slaney mel scale (linear below 1 kHz at 200/3 Hz per mel, logarithmic above with
    step ln(6.4)/27), n_mels + 2 band edges between fmin and fmax, triangular weights from the ramps to the neighbouring edges,
    Slaney area normalisation 2/(f[m+2] - f[m]), float32 result.
PARITY UNPINNED against the reference itself (librosa cannot be imported here); it IS cross-checked against an independent
implementation of the same published algorithm -- torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney") and
torch.stft -- in tests/test_mel_frontend.py.  Only tests/ and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import numpy as np

SR, N_FFT, HOP, N_MELS, FMIN, FMAX, EPS = 22050, 1024, 256, 80, 80.0, 7600.0, 1e-6


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (htk=False, norm='slaney') -> (n_mels, 1 + n_fft//2) float32."""
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def filterbank_ranges(fb: np.ndarray) -> np.ndarray:
    """[k_lo, k_hi) of the non-zero FFT bins of every mel filter (int32 (n_mels, 2))."""
    out = np.zeros((fb.shape[0], 2), dtype=np.int32)
    for m in range(fb.shape[0]):
        nz = np.nonzero(fb[m])[0]
        out[m] = (nz[0], nz[-1] + 1) if nz.size else (0, 0)
    return out


def stft_mag(wav: np.ndarray, n_fft=N_FFT, hop=HOP) -> np.ndarray:
    """|librosa.stft(wav, n_fft, hop, n_fft, 'hann', center=True, pad_mode='constant')| -> (1 + n_fft//2, T) float32."""
    wav = np.asarray(wav, dtype=np.float32)
    y = np.pad(wav, n_fft // 2, mode="constant")
    n_frames = 1 + (len(y) - n_fft) // hop
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)).astype(np.float32)     # periodic Hann, float32 like librosa's fft_window
    frames = np.lib.stride_tricks.as_strided(y, shape=(n_frames, n_fft), strides=(hop * y.itemsize, y.itemsize))
    spec = np.fft.rfft(frames * win[None, :], axis=1).astype(np.complex64)                  # librosa's dtype
    return np.abs(spec).T.astype(np.float32)


def wav2mel(wav: np.ndarray):
    """process_utterance(wav) with the base.yaml parameters -> (wav padded/cropped to T*256 samples, mel (80, T) float32)."""
    spc = stft_mag(wav)
    mel = mel_filterbank() @ spc
    mel = np.log10(np.maximum(EPS, mel)).astype(np.float32)
    pad = (len(wav) // HOP + 1) * HOP - len(wav)
    w = np.pad(np.asarray(wav, dtype=np.float32), (0, pad), mode="constant")[: mel.shape[1] * HOP]
    return w, mel
