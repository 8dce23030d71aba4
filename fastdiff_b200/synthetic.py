"""Deterministic random-init weights and inputs for tests and benchmarks.

There is no network access for checkpoints, so every test/bench uses random-init
weights of the LJSpeech architecture (modules/FastDiff/config/base.yaml:21-33).
The generator is keyed by tensor name, not by module construction order, so the
same seed gives the same state dict for the reference model, the oracle and the
CUDA path on any box with the same torch build.

Magnitudes follow torch's default init for the layer types in
FastDiff_model.py:13-72 (uniform +-1/sqrt(fan_in) for weight and bias;
weight_norm initialises g = ||v||), optionally jittering g so that the
weight-norm fold (w = g v/||v||) is actually exercised.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import torch

# (name, shape, kind) for the LJSpeech config; kind: "wn" conv with weight-norm, "plain".
def _layout(C=32, cond=80, ratios=(8, 8, 4), layers=4, ksz=3, hid=64, kpk=3, e_in=128, e_mid=512, e_out=512):
    L = []
    L.append(("first_audio_conv", (C, 1, 7), "wn"))
    for n, r in enumerate(ratios):
        p = f"lvc_blocks.{n}"
        for i in range(layers):
            L.append((f"{p}.convs.{i}", (C, C, ksz), "wn"))
        L.append((f"{p}.upsample", (C, C, 2 * r), "plain_t"))
        kp = f"{p}.kernel_predictor"
        L.append((f"{kp}.input_conv.0", (hid, cond, 5), "wn"))
        for idx in (1, 3, 6, 8, 11, 13):
            L.append((f"{kp}.residual_conv.{idx}", (hid, hid, kpk), "wn"))
        L.append((f"{kp}.kernel_conv", (C * 2 * C * ksz * layers, hid, kpk), "wn"))
        L.append((f"{kp}.bias_conv", (2 * C * layers, hid, kpk), "wn"))
        L.append((f"{p}.fc_t", (cond, e_out), "plain"))
    for n in range(len(ratios)):
        p = f"downsample.{n}"
        L.append((f"{p}.residual_dense", (C, C, 1), "wn"))
        for i in range(3):
            L.append((f"{p}.conv.{i}", (C, C, 3), "wn"))
    L.append(("fc_t1", (e_mid, e_in), "plain"))
    L.append(("fc_t2", (e_out, e_mid), "plain"))
    L.append(("final_conv.0", (1, C, 7), "wn"))
    return L


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def make_state_dict(seed: int = 1234, g_jitter: float = 0.0, use_weight_norm: bool = True, **arch) -> "OrderedDict[str, torch.Tensor]":
    """Reference-shaped state dict (keys as FastDiff().state_dict(): *.weight_g/*.weight_v/*.bias,
    plain *.weight for ConvTranspose1d/Linear)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape, kind in _layout(**arch):
        g = _gen(name, seed)
        if kind == "plain_t":  # ConvTranspose1d: weight (in, out, k); torch fan_in = out*k for this layout
            fan_in = shape[1] * shape[2]
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
        bound = 1.0 / (fan_in ** 0.5)
        w = _uniform(shape, bound, g)
        nb = shape[1] if kind == "plain_t" else shape[0]
        b = _uniform((nb,), bound, g)
        if kind == "wn" and use_weight_norm:
            sd[f"{name}.bias"] = b
            nrm = w.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1)
            if g_jitter:
                nrm = nrm * (1 + g_jitter * _uniform((shape[0], 1, 1), 1.0, g))
            sd[f"{name}.weight_g"] = nrm
            sd[f"{name}.weight_v"] = w
        else:
            sd[f"{name}.weight"] = w
            sd[f"{name}.bias"] = b
    return sd


def make_inputs(B: int, Tm: int, seed: int = 0, hop: int = 256):
    """Synthetic (x_T, mel) of the benchmark shape: mel ~ U(-6, 1.5) (the reference's log10-mel range,
    base.yaml:15-16), x ~ N(0,1)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    mel = torch.rand((B, 80, Tm), generator=g, dtype=torch.float32) * 7.5 - 6.0
    x = torch.randn((B, 1, Tm * hop), generator=g, dtype=torch.float32)
    return x, mel
