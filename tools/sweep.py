#!/usr/bin/env python
"""BASELINE.json configs as a roofline table: the five named configs + the N x mel-length sweep, on one GPU.

    python tools/sweep.py [--quick] > profiles/rNN_sweep.md

Each row: one complete sampling call (device-resident mel, on-device Philox noise), CUDA-event timed, median of `reps`.
Algorithmic FLOPs = 222,601 per audio sample per reverse step.  N=12 / N=100 have no schedule in the
reference (modules/FastDiff/task/FastDiff.py:76-93 raises NotImplementedError): linspace(1e-4,0.5,12) and
linspace(1e-5,0.06,100) are used, as SURVEY.md 8d suggests.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fastdiff_b200 as fb  # noqa: E402
from fastdiff_b200.sampler import build_steps  # noqa: E402
from fastdiff_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

SCHEDULES = {
    3: [9.0000e-05, 9.0000e-03, 6.0000e-01],
    4: [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01],
    6: [1.7838445955931093e-06, 2.7984189728158526e-05, 0.00043231004383414984, 0.006634317338466644, 0.09357017278671265, 0.6000000238418579],
    8: [6.689325005027058e-07, 1.0033881153503899e-05, 0.00015496854030061513, 0.002387222135439515, 0.035597629845142365, 0.3681158423423767,
        0.4735414385795593, 0.5],
}


def schedule(N):
    if N in SCHEDULES:
        return torch.FloatTensor(SCHEDULES[N])
    if N == 12:
        return torch.linspace(1e-4, 0.5, 12)
    if N == 100:
        return torch.linspace(1e-5, 0.06, 100)
    if N == 200:
        return torch.linspace(0.0001, 0.02, 200)
    if N == 1000:
        return torch.linspace(0.000001, 0.01, 1000)
    raise NotImplementedError(N)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--mode", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    net = fb.FastDiff().to(dev).eval()
    net.load_state_dict(make_state_dict(1234))
    eng = net.engine(dev)
    if args.mode:
        eng.set_mode(args.mode)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0

    def run(B, Tm, N, reps):
        _, steps = build_steps(dh, schedule(N))
        _, mel = make_inputs(B, Tm, 0)
        mel = mel.to(dev)
        x = torch.empty((B, 1, Tm * 256), device=dev)
        eng.sample(x, mel, steps, fill_xT=True, seed=1)  # warm-up (also sizes the workspace)
        torch.cuda.synchronize()
        ts = []
        for r in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.sample(x, mel, steps, fill_xT=True, seed=2 + r)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        n_exec = len(steps)
        sps = B * Tm * 256 / (ms * 1e-3)
        tf = 222601.0 * B * Tm * 256 * n_exec / (ms * 1e-3) / 1e12
        assert torch.isfinite(x).all()
        return ms, n_exec, sps, tf

    mode = {0: "fp32_simt", 1: "tc_3xtf32", 2: "tc_tf32", 3: "tc_3xf16"}[eng.get_mode()]
    print(f"# sweep on {torch.cuda.get_device_name(0)}, arithmetic mode {mode}, peak = {peak} TFLOP/s (MEASURED_PEAKS bf16 sustained)\n")
    print("| config | B | T' (s) | N (executed) | ms / call | ms / reverse step | audio samples/s | algorithmic TFLOP/s | frac of bf16 peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    named = [("configs[0] parity gate shape", 1, 86, 4), ("configs[1] headline", 8, 861, 4), ("configs[2] long loop", 8, 861, 1000),
             ("configs[3] per-GPU share (8 of 64)", 8, 861, 4), ("configs[3] whole batch on ONE GPU", 64, 861, 4)]
    for name, B, Tm, N in named:
        if args.quick and N == 1000:
            continue
        ms, ne, sps, tf = run(B, Tm, N, 1 if N >= 100 else 5)
        print(f"| {name} | {B} | {Tm} ({Tm * 256 / 22050:.1f}) | {N} ({ne}) | {ms:.2f} | {ms / ne:.3f} | {sps:.4g} | {tf:.1f} | {tf / peak:.4f} |", flush=True)
    print()
    print("| sweep: N \\\\ T' | " + " | ".join(f"{Tm} ({Tm * 256 / 22050:.0f} s)" for Tm in (86, 430, 861, 2583)) + " |")
    print("|---|---|---|---|---|")
    for N in ((4, 8) if args.quick else (4, 6, 8, 12, 100, 1000)):
        cells = []
        for Tm in (86, 430, 861, 2583):
            ms, ne, sps, tf = run(8, Tm, N, 1 if N >= 100 else 3)
            cells.append(f"{sps / 1e6:.1f} M/s, {ms / ne:.2f} ms/step, {tf:.0f} TF/s")
        print(f"| N={N} | " + " | ".join(cells) + " |", flush=True)
    print("\n(B = 8 in the sweep; cells: audio samples per second for the whole call, time per reverse step, algorithmic TFLOP/s)")


if __name__ == "__main__":
    main()
