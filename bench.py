#!/usr/bin/env python
"""Benchmark of the FastDiff reverse-diffusion sampling hot path (BASELINE.json metric:
audio samples/sec (22.05 kHz) at N=4 reverse steps).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU yardstick (oracle port of the reference)

A "step" = one complete sampling call (all N=4 reverse steps incl. noise generation) over one batch of
synthetic input.  Workload at every N: BASELINE.json configs[1] per GPU -- batch 8 x 10 s synthetic mel
(T'=861, L=220,416), LJSpeech network, random-init weights -- i.e. weak scaling, utterances sharded over
ranks, one NCCL broadcast of the packed weights at load, no per-step collective.
Prints ONE JSON line on rank 0 (stdout); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

N4_SCHEDULE = [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]  # modules/FastDiff/task/FastDiff.py:88-89
FLOP_PER_SAMPLE_STEP = 222601.0      # SURVEY.md section 8(d), torch FlopCounterMode on the reference
KC_FLOP_PER_FRAME = 2.0 * 24832 * 192  # kernel_conv+bias_conv GEMM: 2*N*K per mel frame per LVC block
METRIC = "audio samples/sec (22.05 kHz) at N=4 reverse steps"
UNIT = "samples/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout must carry exactly ONE JSON line.  Libraries (NCCL's version banner, the sampler's reference-compatible prints) write
# to fd 1 behind Python's back, so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"tensor": float(p["bf16_tflops_sustained"]), "tensor_burst": float(p["bf16_tflops"]), "hbm": float(p["hbm_gbs"]),
                "src": "MEASURED_PEAKS.json"}
    except Exception:
        return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed regions: NVML polled every ~5 ms from a thread that is
    started before the warm-up (so it is already running when the short timed region begins); only samples whose host
    timestamp falls inside a [mark_begin, mark_end] window are reported.  Falls back to `nvidia-smi -lms` if NVML is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.windows, self._t0 = index, [], [], None
        self._stop = threading.Event()
        self.thr = None
        self.backend = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    idx = self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = [getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop.is_set():
                    try:
                        t = time.perf_counter()
                        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                        rs = int(get_reasons(h))
                        self.rows.append((t, sm, mx, pw, [n for n, b in zip(self.NAMES, bits) if rs & b]))
                    except Exception:
                        pass
                    time.sleep(0.004)
            self.backend = "nvml"
        except Exception as e:
            log("NVML unavailable (%r); falling back to nvidia-smi" % (e,))
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception as e2:  # pragma: no cover
                log("clock sampler unavailable:", e2)
                return
            self.proc = proc

            def poll():
                for line in proc.stdout:
                    c = [x.strip() for x in line.split(",")]
                    try:
                        self.rows.append((time.perf_counter(), float(c[0]), float(c[1]), float(c[2]),
                                          [n for n, v in zip(self.NAMES, c[3:7]) if v.lower().startswith("active")]))
                    except Exception:
                        pass
                    if self._stop.is_set():
                        break
            self.backend = "nvidia-smi"
        self.thr = threading.Thread(target=poll, daemon=True)
        self.thr.start()

    def mark_begin(self):
        self._t0 = time.perf_counter()

    def mark_end(self):
        if self._t0 is not None:
            self.windows.append((self._t0, time.perf_counter()))
            self._t0 = None

    def stop(self):
        if self.thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        self._stop.set()
        if getattr(self, "proc", None):
            self.proc.terminate()
        self.thr.join(timeout=2)
        rows = [r for r in self.rows if any(a <= r[0] <= b for a, b in self.windows)]
        sm = [r[1] for r in rows]
        reasons = sorted({n for r in rows for n in r[4]}, key=self.NAMES.index)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                "reasons": reasons, "samples": len(rows), "power_w_max": max(r[3] for r in rows) if rows else None,
                "source": f"{self.backend}, polled during the timed regions (device-resident + e2e)"}


WORKLOAD = "batch=8 x 10 s synthetic mel (T'=861) per GPU, N=4, LJSpeech config, random-init weights"


def bench_config(world):
    """The `config` object -- identical for both arms (`--impl ours` / `--impl reference`) at the same N."""
    return {"workload": WORKLOAD, "global_batch": world * 8, "frames_per_utterance": 861, "reverse_steps": 4,
            "parallelism": f"batch-shard x{world} (weights broadcast once, no per-step collective)",
            "l2": "inputs+activations per call (>=450 MB) exceed the 126 MB L2; no explicit flush"}


class CpuYardstick:
    """The reference's CPU implementation of the path on the host cores (fp32, all intra-op threads).

    kind "reference": the UNMODIFIED reference (`sampling_given_noise_schedule` over `FastDiff.forward`,
    modules/FastDiff/module/util.py:158-235) imported from /root/reference or from the staged copy baseline/_ref
    (oracle/stage_reference.py), `.cuda()` shimmed to identity.  kind "port": oracle/fastdiff_oracle.py -- only when no copy of
    the reference is reachable (it is ~3.8x slower than the reference: its LVC is three einsum taps where the reference unfolds
    once and calls bmm; say so wherever the number is shown)."""

    def __init__(self, seed=1234):
        from fastdiff_b200.synthetic import make_state_dict
        from oracle import refimport
        sd = make_state_dict(seed)
        self.kind = "reference" if refimport.reference_root() else "port"
        if self.kind == "reference":
            self.R = refimport.load("cpu")
            self.model = self.R.FastDiff().eval()
            self.model.load_state_dict(sd)
            self.dh = self.R.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
            self.where = self.R.root
        else:
            from oracle import fastdiff_oracle as O
            self.O, self.W = O, O.fold_weight_norm(sd)
            self.dh = O.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
            self.where = "oracle/fastdiff_oracle.py"

    def sample_once(self, B, Tm):
        """One complete N=4 sampling call; returns seconds."""
        from fastdiff_b200.synthetic import make_inputs
        _, mel = make_inputs(B, Tm, 0)
        sched = torch.FloatTensor(N4_SCHEDULE)
        t0 = time.perf_counter()
        with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):
            if self.kind == "reference":
                self.R.sampling_given_noise_schedule(self.model, (B, 1, Tm * 256), self.dh, sched, condition=mel, ddim=False,
                                                     return_sequence=False)
            else:
                self.O.sample(self.W, (B, 1, Tm * 256), self.dh, sched, mel)
        return time.perf_counter() - t0

    def describe(self, sample):
        return {"unit": UNIT, "cores": torch.get_num_threads(), "kind": self.kind, "sample": sample, "host_cpus": os.cpu_count(),
                "source": self.where, "torch_threads": torch.get_num_threads()}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (see CpuYardstick), all host threads.  A step = one
    complete N=4 sampling call on a BOUNDED sample of the workload: 1 of the 8 utterances of a batch (B=1 x 10 s).  (The reference
    is FASTER per sample at B=1 than at B=8 -- SURVEY.md section 6: 74 k vs 47 k samples/s on 8 cores -- so the bounded sample
    favours the reference.)  Under torchrun only rank 0 runs; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cy = CpuYardstick()
    B, Tm = 1, 861
    for _ in range(args.warmup):
        cy.sample_once(B, Tm)
    ts = [cy.sample_once(B, Tm) for _ in range(args.steps)]
    per = sum(ts) / len(ts)
    val = B * Tm * 256 / per
    sample = (f"{'UNMODIFIED reference' if cy.kind == 'reference' else 'oracle port of the reference'} (torch CPU fp32, "
              f"{torch.get_num_threads()} threads): 1 utterance x 10 s of the 8 x 10 s batch, all N=4 reverse steps, per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(world),
        "cpu_baseline": {"value": val, **cy.describe(sample)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import torch.distributed as dist
    import fastdiff_b200 as fb
    from fastdiff_b200.sampler import build_steps
    from fastdiff_b200.synthetic import make_inputs, make_state_dict
    from fastdiff_b200.weights import pack_state_dict
    from fastdiff_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, Tm = args.batch, args.frames
    L = Tm * 256

    # ---- weights: rank 0 packs, one broadcast of the blob, every rank loads from device memory -------------
    if world > 1:
        if rank == 0:
            blob = torch.from_numpy(pack_state_dict(make_state_dict(1234))).to(dev)
            n = torch.tensor([blob.numel()], device=dev, dtype=torch.int64)
        else:
            n = torch.zeros(1, device=dev, dtype=torch.int64)
        dist.broadcast(n, 0)
        if rank != 0:
            blob = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
        dist.broadcast(blob, 0)   # the only collective on the path
        eng = Engine(device=dev)
        eng.load_blob_device(blob)
        net = None
    else:
        net = fb.FastDiff().to(dev).eval()
        net.load_state_dict(make_state_dict(1234))
        net.noise_mode = "device"
        eng = net.engine(dev)
    if args.mode:
        eng.set_mode(args.mode)
        if net is not None:
            net.mode = args.mode
    for kv in args.opt:
        key, val = kv.split("=")
        eng.set_option(key, int(val))
    mode_name = {0: "fp32_simt", 1: "tc_3xtf32", 2: "tc_tf32", 3: "tc_3xf16"}[eng.get_mode()]

    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    sched = torch.FloatTensor(N4_SCHEDULE)
    _, steps = build_steps(dh, sched)
    _, mel_host = make_inputs(B, Tm, seed=rank)  # each rank owns its own utterances
    mel = mel_host.to(dev)
    x = torch.empty((B, 1, L), dtype=torch.float32, device=dev)

    def one_call(i):
        eng.sample(x, mel, steps, noise=None, seed=1000 + i, fill_xT=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    for i in range(args.warmup):
        one_call(i)
    barrier()

    # ---- timed region: device-resident inputs ------------------------------------------------------------
    eng.timing_enable(True)
    eng.timing_report()  # drop warm-up records
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if clocks:
        clocks.mark_begin()
    ev0.record()
    for i in range(args.steps):
        one_call(args.warmup + i)
    ev1.record()
    barrier()
    if clocks:
        clocks.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    per_kernel = eng.timing_report()
    eng.timing_enable(False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * L / (ms_per_step * 1e-3)

    # ---- e2e: public API, host buffers, H2D/D2H inside the timed region ------------------------------------
    e2e = None
    try:
        mel_pinned = mel_host.pin_memory()
        out_host = torch.empty((B, 1, L), dtype=torch.float32).pin_memory()
        if net is None:
            # multi-rank: same call sequence as sampling_given_noise_schedule, on this rank's shard
            def e2e_call(i):
                m = mel_pinned.to(dev, non_blocking=True)
                eng.sample(x, m, steps, noise=None, seed=5000 + i, fill_xT=True)
                out_host.copy_(x, non_blocking=True)
                torch.cuda.synchronize()
        else:
            def e2e_call(i):
                net.seed = 5000 + i
                with contextlib.redirect_stdout(sys.stderr):
                    y = fb.sampling_given_noise_schedule(net, (B, 1, L), dh, sched, condition=mel_pinned.to(dev, non_blocking=True))
                out_host.copy_(y, non_blocking=True)
                torch.cuda.synchronize()
        e2e_call(0)
        barrier()
        if clocks:
            clocks.mark_begin()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_call(1 + i)
        barrier()
        dt = time.perf_counter() - t0
        if clocks:
            clocks.mark_end()
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * B * L * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": world * mel_host.numel() * 4,
               "d2h_bytes_per_step": world * B * L * 4, "ms_per_step": dt / args.steps * 1e3,
               "api": "fastdiff_b200.sampling_given_noise_schedule" if net is not None else "Engine.sample (batch shard)"}
    except Exception as e:  # pragma: no cover
        log("e2e leg failed:", repr(e))

    clk = clocks.stop() if clocks else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event time inside the timed region) -------------------
    peaks = measured_peaks()
    total_ms = sum(v["ms"] for v in per_kernel.values()) or 1.0
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"]) if per_kernel else None
    roofline = None
    if dom:
        n_dom = per_kernel[dom]["n"]
        avg_ms = per_kernel[dom]["ms"] / n_dom
        frames = B * Tm
        samples = B * L
        flop_per_launch = {
            "kc_gemm": KC_FLOP_PER_FRAME * frames * 3,                    # one launch covers the 3 LVC blocks
            "lvc_layer_b2": (2.0 * 96 * 32 + 2.0 * 96 * 64) * samples,     # dilated conv + LVC at full rate
            "lvc_layer_b1": (2.0 * 96 * 32 + 2.0 * 96 * 64) * samples / 4,
            "lvc_layer_b0": (2.0 * 96 * 32 + 2.0 * 96 * 64) * samples / 32,
        }.get(dom)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                tr = json.load(f).get(dom)
            if tr and B == 8 and Tm == 861:
                traffic = {"bytes_per_launch": tr["bytes_per_launch"], "algorithmic_bytes_per_launch": tr["algorithmic_bytes_per_launch"],
                           "source": "ncu --set full, " + tr["source"]}
        except Exception:
            pass
        if flop_per_launch:
            ach = flop_per_launch / (avg_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["tensor"], "unit": "TFLOP/s",
                        "frac": ach / peaks["tensor"], "traffic": traffic, "peak_source": peaks["src"] + " bf16_tflops_sustained (of measured)",
                        "avg_launch_ms": avg_ms, "launches": n_dom, "share_of_step": per_kernel[dom]["ms"] / (ms_per_step * args.steps),
                        "algorithmic_flop_per_launch": flop_per_launch, "mode": mode_name}
    whole = {"achieved_tflops": FLOP_PER_SAMPLE_STEP * 4 * world * B * L / (ms_per_step * 1e-3) / 1e12,
             "frac_of_bf16_sustained": FLOP_PER_SAMPLE_STEP * 4 * B * L / (ms_per_step * 1e-3) / 1e12 / peaks["tensor"]}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ---------------------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        try:
            cy = CpuYardstick()
            cy.sample_once(1, 86)  # warm-up (1 s)
            reps = [cy.sample_once(1, 861) for _ in range(3 if cy.kind == "reference" else 1)]
            tcpu = min(reps)
            cpu_baseline = {"value": 861 * 256 / tcpu, **cy.describe(
                f"{'UNMODIFIED reference' if cy.kind == 'reference' else 'oracle port of the reference'} (torch CPU fp32): 1 utterance x 10 s "
                f"of the batch, all N=4 steps, best of {len(reps)} runs"), "seconds": tcpu}
        except Exception as e:  # pragma: no cover
            log("cpu baseline failed:", repr(e))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if mode_name != "tc_tf32" else "tf32", "data": "synthetic",
        "config": (bench_config(world) if (B, Tm) == (8, 861) else
                   {**bench_config(world), "workload": f"batch={B} x {Tm * 256 / 22050:.1f} s synthetic mel (T'={Tm}) per GPU, N=4, LJSpeech config, "
                    "random-init weights", "global_batch": world * B, "frames_per_utterance": Tm}),
        "arith_mode": mode_name, "noise": "on-device Philox4x32-10 (inside the timed region)",
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
        **({"experiment": {"options": args.opt, "nvcc_extra": os.environ.get("FD_NVCC_EXTRA", "")}} if (args.opt or os.environ.get("FD_NVCC_EXTRA")) else {}),
        "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in per_kernel.items()},
        "kernel_ms_note": "CUDA events around every launch on its own stream; the DBlock chain runs on a side stream concurrently with embed / "
                          "kernel predictor / GEMM, so those classes include time spent sharing the SMs and the classes sum to more than ms_per_step",
        "whole_step": whole,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=None, choices=[None, "fp32_simt", "tc_3xtf32", "tc_tf32", "tc_3xf16"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=861)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option key=value (fd_set_option), repeatable")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
