# Multi-GPU check of the time-axis shard mode (run under torchrun, NCCL): long utterances, N = 4, default arithmetic mode.
#   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_scripts/timeshard_check.py
# Per shape: (a) "reference" noise (full-size host draws on every rank) == the single-GPU parity-mode result, bitwise;
#            (b) "device" noise -- the LATENCY mode: per-rank Philox windows -- == the single-GPU device-noise result, bitwise, and timed
#                against the single-GPU call with CUDA events (max over ranks).
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch.distributed as dist
import fastdiff_b200 as fb
from fastdiff_b200.shard import ShardedFastDiff
from fastdiff_b200.timeshard import TimeShardedSampler
from fastdiff_b200.synthetic import make_inputs, make_state_dict
from fastdiff_b200.sampler import build_steps
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
sh = ShardedFastDiff(make_state_dict(1234) if rank == 0 else None)
ts = TimeShardedSampler(sh.engine)
N4 = torch.FloatTensor([3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01])
dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
_, steps = build_steps(dh, N4)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


for B, Tm in ((1, 2583), (1, 5166), (2, 861)):
    _, mel = make_inputs(B, Tm, 3)
    size = (B, 1, Tm * 256)
    torch.manual_seed(5)
    out_ref = ts.sample(size, dh, N4, mel)                                   # (a) reference RNG stream
    out_dev, ms_sh = timed(lambda: ts.sample(size, dh, N4, mel.to(dev), noise="device", seed=9))   # (b) latency mode
    if rank == 0:
        torch.manual_seed(5)
        x = torch.normal(0, 1, size=size).cuda()
        zs = torch.stack([torch.normal(0, 1, size=size) for _ in range(3)]).cuda()
        ref = sh.engine.sample(x, mel.cuda(), steps, noise=zs)
        err_a = (out_ref - ref).abs().max().item()
    md = mel.to(dev)

    def single():
        xx = torch.empty(size, dtype=torch.float32, device=dev)
        return sh.engine.sample(xx, md, steps, noise=None, seed=9, fill_xT=True)
    ref_dev, ms_un = timed(single)                                           # every rank runs the unsharded call (same work, same time)
    if rank == 0:
        err_b = (out_dev - ref_dev).abs().max().item()
        print(f"time-shard x{world}: B={B} T'={Tm} ({Tm*256/22050:.1f} s)  reference-noise max|sharded - unsharded| = {err_a:.3e}  "
              f"device-noise max|sharded - unsharded| = {err_b:.3e}  |  latency: sharded {ms_sh:.2f} ms vs one GPU {ms_un:.2f} ms "
              f"(x{ms_un / ms_sh:.2f})", flush=True)
    dist.barrier()
dist.destroy_process_group()
