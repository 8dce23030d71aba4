"""Reverse-diffusion sampler -- drop-in for the functions of /root/reference/modules/FastDiff/module/util.py
on the sampling path: ``sampling_given_noise_schedule`` (:158-235), ``compute_hyperparams_given_schedule``
(:365-390), ``map_noise_scale_to_time_step`` (:394-404), ``calc_diffusion_step_embedding`` (:407-432),
``std_normal`` (:63-68).  Same signatures, same asserts/prints, same float results for the host-side tables
(they are evaluated with the same fp32 operation sequence, just vectorised); the N-step loop itself runs
inside the CUDA library (fd_sample): no PyTorch op executes inside the loop.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from ._lib import fd_step

f32 = np.float32


def std_normal(size, device=None):
    """util.py:63-68: CPU default generator, then to the device."""
    z = torch.normal(0, 1, size=size)
    return z.to(device) if device is not None else (z.cuda() if torch.cuda.is_available() else z)


def _sqrt_on(v: np.ndarray, device) -> np.ndarray:
    """sqrt evaluated by torch on ``device`` -- the reference takes these roots with torch.sqrt on whatever device
    the schedule lives on, and torch's CPU (AVX-512) sqrt is not always the correctly-rounded one, so to return the
    floats the reference returns for the same inputs we use the same routine rather than numpy's."""
    with np.errstate(invalid="ignore"):
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=f32)).to(device)
        return torch.sqrt(t).cpu().numpy()


def _f32_tables(beta: np.ndarray, device="cpu"):
    """alpha_t = sqrt(prod(1-beta)), sigma_t per util.py:187-195 / :377-384, sequential fp32 like the reference loop."""
    beta = beta.astype(f32)
    alpha = (f32(1) - beta).astype(f32)
    sigma = beta.copy()
    for t in range(1, len(beta)):
        alpha[t] = alpha[t] * alpha[t - 1]
        sigma[t] = sigma[t] * ((f32(1) - alpha[t - 1]) / (f32(1) - alpha[t]))
    return _sqrt_on(alpha, device), _sqrt_on(sigma, device)


def compute_hyperparams_given_schedule(beta):
    """util.py:365-390.  Returns {"T","beta","alpha","sigma"}; tensors stay on beta's device."""
    T = len(beta)
    dev = beta.device
    alpha, sigma = _f32_tables(beta.detach().cpu().numpy(), dev)
    return {"T": T, "beta": beta, "alpha": torch.from_numpy(alpha).to(dev), "sigma": torch.from_numpy(sigma).to(dev)}


def _map_noise_scale(a: np.float32, alpha: np.ndarray):
    """util.py:394-404 vectorised: first t with alpha[t+1] <= a <= alpha[t]; t + (alpha[t]-a)/(alpha[t]-alpha[t+1])."""
    if a < alpha[-1]:
        return len(alpha) - 1
    if a > alpha[0]:
        return 0
    hit = np.nonzero((alpha[1:] <= a) & (a <= alpha[:-1]))[0]
    if hit.size == 0:
        return -1
    t = int(hit[0])
    d = f32(alpha[t] - a)
    d = f32(d / f32(alpha[t] - alpha[t + 1]))
    return t + float(d)


def map_noise_scale_to_time_step(alpha_infer, alpha):
    a = f32(alpha_infer.item() if torch.is_tensor(alpha_infer) else alpha_infer)
    al = alpha.detach().cpu().numpy().astype(f32) if torch.is_tensor(alpha) else np.asarray(alpha, f32)
    return _map_noise_scale(a, al)


def calc_diffusion_step_embedding(diffusion_steps, diffusion_step_embed_dim_in):
    """util.py:407-432 (kept for API completeness; the CUDA path evaluates it inside k_embed)."""
    assert diffusion_step_embed_dim_in % 2 == 0
    half_dim = diffusion_step_embed_dim_in // 2
    _embed = np.log(10000) / (half_dim - 1)
    _embed = torch.exp(torch.arange(half_dim) * -_embed).to(diffusion_steps.device)
    _embed = diffusion_steps * _embed
    return torch.cat((torch.sin(_embed), torch.cos(_embed)), 1)


calc_noise_scale_embedding = calc_diffusion_step_embedding   # util.py:71-99: the same formula under its DiffWave name (WaveNet.py:6)


def build_steps(diffusion_hyperparams, inference_noise_schedule, ddim=False):
    """Host prologue of util.py:180-204 -> (steps_infer list, [fd_step] in execution order n = N-1..0)."""
    _dh = diffusion_hyperparams
    T, alpha = _dh["T"], _dh["alpha"]
    assert len(alpha) == T
    sched = inference_noise_schedule
    dev = sched.device if torch.is_tensor(sched) else (alpha.device if torch.is_tensor(alpha) else "cpu")
    beta = (sched.detach().cpu().numpy() if torch.is_tensor(sched) else np.asarray(sched)).astype(f32)
    alpha_tr = alpha.detach().cpu().numpy().astype(f32)
    alpha_infer, sigma_infer = _f32_tables(beta, dev)
    steps_infer: List[float] = []
    for n in range(len(beta)):
        s = _map_noise_scale(alpha_infer[n], alpha_tr)
        if s >= 0:
            steps_infer.append(s)
    steps_f32 = np.asarray(steps_infer, dtype=f32)  # torch.FloatTensor(steps_infer)
    N = len(steps_f32)
    one = f32(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        a2 = (alpha_infer * alpha_infer).astype(f32)                    # alpha ** 2.
        s1ma2 = _sqrt_on((one - a2).astype(f32), dev)                   # sqrt(1 - alpha^2)
        coef = (beta / s1ma2).astype(f32)                               # beta / sqrt(1 - alpha^2)     util.py:226
        div = _sqrt_on((one - beta).astype(f32), dev)                   # sqrt(1 - beta)               util.py:227
        alpha_next = (alpha_infer / div).astype(f32)                    # util.py:220
        c1 = (alpha_next / alpha_infer).astype(f32)
        c2 = (-s1ma2 * c1).astype(f32)
        c3 = _sqrt_on((one - (alpha_next * alpha_next).astype(f32)).astype(f32), dev)
    out = []
    for n in range(N - 1, -1, -1):
        st = fd_step()
        st.t = float(steps_f32[n])
        st.coef_eps, st.div, st.sigma = float(coef[n]), float(div[n]), float(sigma_infer[n])
        st.add_noise = 1 if n > 0 else 0
        if ddim:
            st.c1, st.c2, st.c3 = float(c1[n]), float(c2[n]), float(c3[n])
        out.append(st)
    return steps_infer, out


def sampling_given_noise_schedule(net, size, diffusion_hyperparams, inference_noise_schedule, condition=None,
                                  ddim=False, return_sequence=False):
    """Perform the complete sampling p(x_0|x_T) -- signature and semantics of util.py:158-235.

    net: fastdiff_b200.FastDiff on a CUDA device.  ``net.noise_mode``: "reference" draws x_T and every z on the CPU
    default generator in the reference's order (bit-compatible RNG stream; one H2D copy up front), "device" draws
    them on the GPU (Philox4x32-10, seed ``net.seed``) so nothing crosses PCIe inside the call.
    """
    assert len(size) == 3
    steps_infer, steps = build_steps(diffusion_hyperparams, inference_noise_schedule, ddim)
    print(steps_infer, flush=True)
    N = len(steps)
    print('begin sampling, total number of reverse steps = %s' % N)

    dev = next(net.parameters()).device
    if not hasattr(net, "engine"):
        return _sample_with_foreign_net(net, size, steps, steps_infer, condition, ddim, return_sequence, dev)
    eng = net.engine(dev)
    B, _, L = size
    if condition is None:
        raise ValueError("FastDiff is a conditional vocoder: `condition` (mel) is required")
    cond = condition
    if cond.dim() == 2:
        cond = cond.unsqueeze(0)
    if cond.shape[0] != B:
        cond = cond.expand(B, -1, -1)
    if cond.shape[-1] * 256 != L:
        raise AssertionError("length of (x, kernel) is not matched")  # modules.py:236
    cond = cond.to(dev, torch.float32).contiguous()

    n_noise = 0 if ddim else sum(s.add_noise for s in steps)
    seq = torch.empty((N + 1, B, 1, L), dtype=torch.float32, device=dev) if return_sequence else None
    with torch.no_grad():
        if getattr(net, "noise_mode", "reference") == "reference":
            x = torch.normal(0, 1, size=size).to(dev)
            noise = None
            if n_noise:
                noise = torch.stack([torch.normal(0, 1, size=size) for _ in range(n_noise)]).to(dev)
            eng.sample(x, cond, steps, noise=noise, ddim=ddim, seq=seq)
        else:
            x = torch.empty(size, dtype=torch.float32, device=dev)
            eng.sample(x, cond, steps, noise=None, seed=int(getattr(net, "seed", 0)), fill_xT=True, ddim=ddim, seq=seq)
    if return_sequence:
        return [seq[i] for i in range(N + 1)]
    return x


_update_engines = {}


def _update_engine(dev):
    """A weight-less handle per device for fd_reverse_update (the kernel needs no network)."""
    from .engine import Engine
    key = str(dev)
    if key not in _update_engines:
        _update_engines[key] = Engine(device=dev)
    return _update_engines[key]


def _sample_with_foreign_net(net, size, steps, steps_infer, condition, ddim, return_sequence, dev):
    """The loop of util.py:210-234 for any OTHER denoiser with the reference's `net((x, condition, diffusion_steps))` signature --
    the reference's WaveNet_vocoder (modules/FastDiff/module/WaveNet.py:156) and ParallelWaveGANGenerator_Diffusion
    (modules/parallel_wavegan/models/parallel_wavegan.py:23) share this sampler.  The network is the caller's torch module and runs
    as it is; each reverse-step update (the reference's 3-6 eager ops) is one in-place CUDA kernel (fd_reverse_update).  Same RNG
    stream as the reference: x_T, then one CPU draw after every network call with n > 0."""
    if dev.type != "cuda" and str(dev) not in _update_engines:   # (the CPU test-suite registers its emulation build of the library here)
        raise RuntimeError("sampling_given_noise_schedule: the network must live on a CUDA device (there is no CPU path)")
    eng = _update_engine(dev)
    N = len(steps)
    x = torch.normal(0, 1, size=size).to(dev)
    xs = [x.clone()] if return_sequence else None
    with torch.no_grad():
        for i, st in enumerate(steps):                      # execution order n = N-1 .. 0
            n = N - 1 - i
            diffusion_steps = (steps_infer[n] * torch.ones((size[0], 1))).to(dev)
            eps = net((x, condition, diffusion_steps,))
            if tuple(eps.shape) != tuple(x.shape):
                raise ValueError(f"the network returned shape {tuple(eps.shape)}, expected {tuple(x.shape)}")
            z = torch.normal(0, 1, size=size).to(dev) if (st.add_noise and not ddim) else None
            eng.reverse_update(x, eps, st, z=z, ddim=ddim)
            if return_sequence:
                xs.append(x.clone())
    return xs if return_sequence else x
