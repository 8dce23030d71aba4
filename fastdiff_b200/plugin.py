"""Glue for the reference's own callers of the path.

``patch_reference_task()`` is the whole integration: it swaps the three names the reference task module binds at import
time (`FastDiff`, `sampling_given_noise_schedule`, `compute_hyperparams_given_schedule` --
/root/reference/modules/FastDiff/task/FastDiff.py:5,9) for the B200 implementations, so the UNMODIFIED
`FastDiffTask.build_model` / `test_step` (task/FastDiff.py:16-42, 60-119), `tasks/run.py` and the Trainer drive the
CUDA path.  Nothing from the reference is copied; it must be importable (on `sys.path`) for this module's functions to work.
"""
from __future__ import annotations

import importlib

import numpy as np
import torch

from .model import FastDiff
from .sampler import compute_hyperparams_given_schedule, sampling_given_noise_schedule

N4_SCHEDULE = [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]  # task/FastDiff.py:88-89


def patch_reference_task(module_name: str = "modules.FastDiff.task.FastDiff"):
    """Import the reference task module and rebind its model/sampler names; returns the (unmodified) task class."""
    ref_task = importlib.import_module(module_name)
    ref_task.FastDiff = FastDiff
    ref_task.sampling_given_noise_schedule = sampling_given_noise_schedule
    ref_task.compute_hyperparams_given_schedule = compute_hyperparams_given_schedule
    return ref_task.FastDiffTask


class FastDiffVocoder:
    """`BaseVocoder`-shaped wrapper (vocoders/base_vocoder.py:23-40): spec2wav(mel [T,80]) -> wav [T*hop].
    Register it in the reference with `register_vocoder(FastDiffVocoder)` (vocoders/base_vocoder.py:6-9)."""

    def __init__(self, state_dict=None, ckpt_path=None, device="cuda", schedule=None, seed=None):
        self.device = torch.device(device)
        self.model = FastDiff().to(self.device).eval()
        if ckpt_path is not None:
            state_dict = torch.load(ckpt_path, map_location="cpu")["state_dict"]["model"]
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=True)
        self.dh = compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))  # base.yaml:38-40
        self.schedule = torch.FloatTensor(N4_SCHEDULE if schedule is None else schedule)
        if seed is not None:
            self.model.noise_mode, self.model.seed = "device", int(seed)

    def spec2wav(self, mel, int16: bool = False, **kwargs):
        """mel [T, 80] -> wav [T * 256] float32 (vocoders/base_vocoder.py:24-31).  int16=True additionally applies the step AFTER the
        path on the device -- peak-normalise, x 32767, truncate to int16 (task/FastDiff.py:110, utils/audio.py:11-16; fd_wav_int16,
        bit-identical to the reference's float ops) -- so that only 2 bytes per sample leave the GPU."""
        c = torch.as_tensor(np.asarray(mel), dtype=torch.float32).t().unsqueeze(0).to(self.device)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            y = sampling_given_noise_schedule(self.model, (1, 1, c.shape[-1] * 256), self.dh, self.schedule.clone(), condition=c)
        if int16:
            return self.model.engine(self.device).wav_int16(y).view(-1).cpu().numpy()
        return y.view(-1).cpu().numpy()

    @staticmethod
    def wav2spec(wav_fn, engine=None):
        """(wav [T*256], mel [T,80]) of a 22.05 kHz wav file or float array -- `process_utterance`
        (data_gen/tts/data_gen_utils.py:93-147 via vocoders/base_vocoder.py:32-40) on the device (fastdiff_b200/mel.py).
        The reference loads with librosa.core.load(sr=22050), i.e. PCM scaled to [-1, 1) and resampled if needed; only the
        native-rate case is handled here (no resampler on the path)."""
        from .engine import Engine
        from .mel import SAMPLE_RATE, wav2mel
        if isinstance(wav_fn, (str, bytes)) or hasattr(wav_fn, "__fspath__"):
            from scipy.io import wavfile
            sr, data = wavfile.read(wav_fn)
            if sr != SAMPLE_RATE:
                raise ValueError(f"{wav_fn}: {sr} Hz; resample to {SAMPLE_RATE} Hz first")
            if data.ndim > 1:
                data = data.mean(axis=1)
            if np.issubdtype(data.dtype, np.integer):
                data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
            wav = np.ascontiguousarray(data, dtype=np.float32)
        else:
            wav = np.ascontiguousarray(wav_fn, dtype=np.float32)
        eng = engine if engine is not None else Engine(device="cuda")
        w, mel = wav2mel(eng, torch.from_numpy(wav))
        return w[0].cpu().numpy(), mel[0].t().contiguous().cpu().numpy()
