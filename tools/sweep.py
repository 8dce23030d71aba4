#!/usr/bin/env python
"""BASELINE.json configs as a roofline table: the five named configs + the N x mel-length sweep (configs[4]), on 1 or G GPUs.

    python tools/sweep.py [--quick] > profiles/rNN_sweep.md
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py > profiles/rNN_sweep_8gpu.md

Under torchrun every rank drives fastdiff_b200.shard.ShardedFastDiff on its own 8 utterances (weak scaling: one NCCL broadcast of the weights
at load, no per-step collective); a cell's time is the MAX over ranks of the CUDA-event time, its throughput the samples of ALL ranks / that.

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
    ap.add_argument("--no-n1000", action="store_true", help="skip the N = 1000 row / configs[2] (minutes of GPU time)")
    args = ap.parse_args()
    import torch.distributed as dist
    from fastdiff_b200.shard import ShardedFastDiff
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sh = ShardedFastDiff(make_state_dict(1234) if rank == 0 else None, device=dev)
    eng = sh.engine
    if args.mode:
        eng.set_mode(args.mode)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        peak = 1400.0

    def out(*a):
        if rank == 0:
            print(*a, flush=True)

    def run(B, Tm, N, reps):
        """B utterances PER GPU.  Returns (ms per call: max over ranks, executed steps, whole-job samples/s, whole-job algorithmic TFLOP/s)."""
        _, steps = build_steps(dh, schedule(N))
        _, mel = make_inputs(B, Tm, rank)
        mel = mel.to(dev)
        x = torch.empty((B, 1, Tm * 256), device=dev)
        eng.sample(x, mel, steps, fill_xT=True, seed=1)  # warm-up (sizes the workspace, captures the graph where N <= 64)
        eng.sample(x, mel, steps, fill_xT=True, seed=1)
        torch.cuda.synchronize()
        ts = []
        for r in range(reps):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.sample(x, mel, steps, fill_xT=True, seed=2 + r)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        ms = statistics.median(ts)
        n_exec = len(steps)
        sps = world * B * Tm * 256 / (ms * 1e-3)
        tf = 222601.0 * world * B * Tm * 256 * n_exec / (ms * 1e-3) / 1e12
        assert torch.isfinite(x).all()
        return ms, n_exec, sps, tf

    mode = {0: "fp32_simt", 1: "tc_3xtf32", 2: "tc_tf32", 3: "tc_3xf16"}[eng.get_mode()]
    out(f"# sweep on {world} x {torch.cuda.get_device_name(local)}, arithmetic mode {mode}, peak = {peak} TFLOP/s per GPU (MEASURED_PEAKS bf16 sustained); "
        f"B = utterances per GPU, throughput = whole job, time = max over ranks\n")
    out("| config | B per GPU | T' (s) | N (executed) | ms / call | ms / reverse step | audio samples/s (all GPUs) | algorithmic TFLOP/s (all GPUs) | frac of bf16 peak (per GPU) |")
    out("|---|---|---|---|---|---|---|---|---|")
    named = [("configs[0] parity gate shape", 1, 86, 4), ("configs[1] headline", 8, 861, 4), ("configs[2] long loop", 8, 861, 1000),
             ("configs[3] 64 utterances over 8 GPUs = 8 per GPU", 8, 861, 4)]
    if world == 1:
        named.append(("configs[3] whole batch of 64 on ONE GPU", 64, 861, 4))
    for name, B, Tm, N in named:
        if (args.quick or args.no_n1000) and N == 1000:
            continue
        ms, ne, sps, tf = run(B, Tm, N, 1 if N >= 100 else 5)
        out(f"| {name} | {B} | {Tm} ({Tm * 256 / 22050:.1f}) | {N} ({ne}) | {ms:.2f} | {ms / ne:.3f} | {sps:.4g} | {tf:.1f} | {tf / peak / world:.4f} |")
    out()
    out("| sweep: N \\\\ T' | " + " | ".join(f"{Tm} ({Tm * 256 / 22050:.0f} s)" for Tm in (86, 430, 861, 2583)) + " |")
    out("|---|---|---|---|---|")
    Ns = (4, 8) if args.quick else ((4, 6, 8, 12, 100) if args.no_n1000 else (4, 6, 8, 12, 100, 1000))
    for N in Ns:
        cells = []
        for Tm in (86, 430, 861, 2583):
            ms, ne, sps, tf = run(8, Tm, N, 1 if N >= 100 else 3)
            cells.append(f"{sps / 1e6:.1f} M/s, {ms / ne:.2f} ms/step, {tf:.0f} TF/s ({tf / peak / world:.3f})")
        out(f"| N={N} | " + " | ".join(cells) + " |")
    out("\n(B = 8 per GPU in the sweep; cells: audio samples per second of the whole job, time per reverse step, algorithmic TFLOP/s of the whole job and, "
        "in brackets, the per-GPU fraction of the measured bf16 peak.  N = 12 / 100 use linspace(1e-4, 0.5, 12) / linspace(1e-5, 0.06, 100): the reference "
        "defines no schedule for them, task/FastDiff.py:76-93)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
