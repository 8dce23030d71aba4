"""Batch-shard mode on 2 CPU ranks (gloo): one broadcast of the packed blob, each rank runs its slice through the
(emulated) CUDA source, gathered result == unsharded result bitwise, and both == the oracle within tolerance."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_lib, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fastdiff_b200.shard import ShardedFastDiff
    from fastdiff_b200.synthetic import make_inputs, make_state_dict
    sd = make_state_dict(1234, g_jitter=0.1) if rank == 0 else None
    sh = ShardedFastDiff(sd, device="cpu", lib_path=emu_lib)
    B, Tm = 3, 1
    x, mel = make_inputs(B, Tm, 2)
    t = torch.tensor([7.413235, 498.0537, 74.99228])
    full = sh.denoise(x, mel, t, gather=True)
    lo, hi = sh.my_slice(B)
    # B < world: rank 1's slice is empty -- it must skip the engine call and still take part in the gather (no hang, no FdError)
    import fastdiff_b200 as fb
    x1, mel1 = make_inputs(1, 1, 5)
    one = sh.denoise(x1, mel1, torch.tensor([74.99228]), gather=True)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    torch.manual_seed(5)
    noise = [torch.normal(0, 1, size=(1, 1, 256)) for _ in range(4)]
    smp = sh.sample((1, 1, 256), dh, torch.FloatTensor([3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]), mel1, noise=noise, gather=True)
    assert tuple(one.shape) == (1, 1, 256) and tuple(smp.shape) == (1, 1, 256) and sh.my_slice(1) == ((0, 1) if rank == 0 else (1, 1))
    if rank == 0:
        q.put((full, sh.blob_bytes, one, smp))
    else:
        q.put((lo, hi, sh.blob_bytes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batch_shard_matches_unsharded(emu_lib, synth):
    from fastdiff_b200.engine import Engine
    from fastdiff_b200.synthetic import make_inputs
    from fastdiff_b200.weights import pack_state_dict
    from oracle import fastdiff_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emu_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    import time
    got, t0 = [], time.time()
    while len(got) < 2 and time.time() - t0 < 600:
        try:
            got.append(q.get(timeout=2))
        except queue.Empty:
            assert all(p.exitcode in (None, 0) for p in procs), "a rank died"
    assert len(got) == 2
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = next(g for g in got if torch.is_tensor(g[0]))
    other = next(g for g in got if not torch.is_tensor(g[0]))
    sd, W = synth
    assert other[:2] == (2, 3) and other[2] == full[1] == pack_state_dict(sd).nbytes
    x, mel = make_inputs(3, 1, 2)
    t = torch.tensor([7.413235, 498.0537, 74.99228])
    eng = Engine(device="cpu", lib_path=emu_lib)
    eng.load_blob(pack_state_dict(sd))
    unsharded = eng.denoise(x, mel, t)
    assert torch.equal(full[0], unsharded)
    assert (unsharded - O.denoise(W, x, mel, t.reshape(3, 1))).abs().max() < 5e-5
    x1, mel1 = make_inputs(1, 1, 5)
    assert torch.equal(full[2], eng.denoise(x1, mel1, torch.tensor([74.99228])))       # B = 1 on 2 ranks == unsharded, bitwise
    assert torch.isfinite(full[3]).all()
