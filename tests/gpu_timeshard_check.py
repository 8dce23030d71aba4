# 2-GPU check of the time-axis shard mode (run under torchrun, NCCL): one 30 s utterance, N = 4, default arithmetic mode.
#   torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_timeshard_check.py
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import fastdiff_b200 as fb
from fastdiff_b200.shard import ShardedFastDiff
from fastdiff_b200.timeshard import TimeShardedSampler
from fastdiff_b200.synthetic import make_inputs, make_state_dict
from fastdiff_b200.sampler import build_steps
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
sh = ShardedFastDiff(make_state_dict(1234) if rank == 0 else None)
ts = TimeShardedSampler(sh.engine)
N4 = torch.FloatTensor([3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01])
dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
for B, Tm in ((1, 2583), (2, 861)):
    _, mel = make_inputs(B, Tm, 3)
    size = (B, 1, Tm * 256)
    torch.manual_seed(5)
    out = ts.sample(size, dh, N4, mel)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    torch.manual_seed(5)
    out = ts.sample(size, dh, N4, mel)
    torch.cuda.synchronize(); dist.barrier()
    t_sh = time.perf_counter() - t0
    if rank == 0:
        _, steps = build_steps(dh, N4)
        torch.manual_seed(5)
        x = torch.normal(0, 1, size=size).cuda()
        zs = torch.stack([torch.normal(0, 1, size=size) for _ in range(3)]).cuda()
        sh.engine.sample(x.clone(), mel.cuda(), steps, noise=zs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref = sh.engine.sample(x, mel.cuda(), steps, noise=zs)
        torch.cuda.synchronize()
        t_un = time.perf_counter() - t0
        err = (out - ref).abs().max().item()
        print(f"time-shard x{world}: B={B} T'={Tm} ({Tm*256/22050:.1f} s)  max|sharded - unsharded| = {err:.3e} (rms {ref.pow(2).mean().sqrt():.3f})  "
              f"sharded call {t_sh*1e3:.2f} ms incl. host RNG + H2D  |  unsharded device-only {t_un*1e3:.2f} ms", flush=True)
    dist.barrier()
dist.destroy_process_group()
