"""Time-axis shard mode on 2 CPU ranks (gloo, emulated CUDA source): one utterance split along time with a
17-frame halo exchanged after every reverse step == the unsharded sampler under the same RNG stream, for the N = 4 schedule and a
batch of 2 (23 + 22 frames)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N4 = [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]
TM = 45


def _worker(rank, world, port, emu_lib, q, mode, B, noise):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fastdiff_b200 as fb
    from fastdiff_b200.shard import ShardedFastDiff
    from fastdiff_b200.synthetic import make_inputs, make_state_dict
    from fastdiff_b200.timeshard import TimeShardedSampler
    sd = make_state_dict(1234, g_jitter=0.1) if rank == 0 else None
    sh = ShardedFastDiff(sd, device="cpu", lib_path=emu_lib)           # weights: one broadcast
    sh.engine.set_mode(mode)
    ts = TimeShardedSampler(sh.engine)
    _, mel = make_inputs(B, TM, 8)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    torch.manual_seed(21)                                              # same stream on every rank
    out = ts.sample((B, 1, TM * 256), dh, torch.FloatTensor(N4), mel, gather=True, noise=noise, seed=77)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,B,noise", [("fp32_simt", 1, "reference"), ("tc_3xf16", 2, "reference"), ("tc_3xf16", 2, "device")])
def test_two_rank_time_shard_matches_unsharded(emu_lib, synth, mode, B, noise):
    """noise = "device" is the latency mode: every rank draws only its own window on the device (Philox indexed by the sample's position in
    the whole utterance, draw numbers of the single-call loop) and must reproduce the single-engine device-noise result."""
    import fastdiff_b200 as fb
    from fastdiff_b200.engine import Engine
    from fastdiff_b200.sampler import build_steps
    from fastdiff_b200.synthetic import make_inputs
    from fastdiff_b200.weights import pack_state_dict
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + (7 if mode == "tc_3xf16" else 0) + (13 if noise == "device" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib, q, mode, B, noise)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    import time
    got, t0 = None, time.time()
    while got is None and time.time() - t0 < 900:
        try:
            got = q.get(timeout=2)
        except queue.Empty:
            assert all(p.exitcode in (None, 0) for p in procs), "a rank died"
    assert got is not None
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sd, _ = synth
    eng = Engine(device="cpu", lib_path=emu_lib)
    eng.load_blob(pack_state_dict(sd))
    eng.set_mode(mode)
    _, mel = make_inputs(B, TM, 8)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    _, steps = build_steps(dh, torch.FloatTensor(N4))
    torch.manual_seed(21)
    size = (B, 1, TM * 256)
    if noise == "device":
        x = torch.empty(size)
        eng.sample(x, mel, steps, noise=None, seed=77, fill_xT=True)
    else:
        x = torch.normal(0, 1, size=size)
        zs = torch.stack([torch.normal(0, 1, size=size) for _ in range(3)])
        eng.sample(x, mel, steps, noise=zs)
    assert got.shape == x.shape
    err = (got - x).abs().max().item()
    assert err <= 1e-6, err          # same arithmetic per output sample; only the tiling differs (tc_3xf16: the tensor-core model)
