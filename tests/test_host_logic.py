"""Host-side logic and the C-ABI surface, no GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fastdiff_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", txt)))


def test_cabi_library_exports_every_declared_symbol(cuda_lib):
    """The in-tree .so loads without a GPU and exports exactly what include/fastdiff_b200.h declares."""
    from fastdiff_b200 import _lib
    lib = ctypes.CDLL(cuda_lib)
    decl = _declared_symbols()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == decl  # the ctypes binding covers the header, nothing more
    lib.fd_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.fd_version()


def test_emu_library_exports_every_declared_symbol(emu_lib):
    lib = ctypes.CDLL(emu_lib)
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_missing_extension_fails_loudly(tmp_path):
    from fastdiff_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "nope.so"))


def test_cpu_forward_refuses_without_gpu():
    import fastdiff_b200 as fb
    net = fb.FastDiff()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            net((torch.zeros(1, 1, 256), torch.zeros(1, 80, 1), torch.zeros(1, 1)))


def test_unsupported_architecture_is_rejected(emu_lib):
    from fastdiff_b200._lib import FdError
    from fastdiff_b200.engine import Engine
    with pytest.raises(FdError, match="base.yaml"):
        Engine(arch={"upsample_ratios": [4, 4, 4]}, device="cpu", lib_path=emu_lib)
    eng = Engine(device="cpu", lib_path=emu_lib)
    with pytest.raises(FdError, match="not loaded"):
        eng.denoise(torch.zeros(1, 1, 256), torch.zeros(1, 80, 1), torch.zeros(1))
    with pytest.raises(FdError, match="magic"):
        eng.load_blob(np.zeros(4096, dtype=np.uint8))


def test_packer_layouts(synth):
    """Spot-check the permutations the kernels rely on (fd_blob.h) against the reference channel maps (modules.py:333-342)."""
    from fastdiff_b200 import weights as Wt
    from oracle import fastdiff_oracle as O
    sd, W = synth
    S = Wt.build_sections(sd)
    kc = S["LB1_KC_W"].reshape(192, 24832)
    kw = W["lvc_blocks.1.kernel_predictor.kernel_conv.weight"]  # (24576, 64, 3)
    bw = W["lvc_blocks.1.kernel_predictor.bias_conv.weight"]    # (256, 64, 3)
    rng = np.random.default_rng(0)
    for _ in range(200):
        l, i, o, k, c, j = (int(rng.integers(n)) for n in (4, 32, 64, 3, 64, 3))
        assert kc[j * 64 + c, l * 6208 + ((k * 64 + o) * 8 + ((i // 4) ^ (o & 7))) * 4 + i % 4] == kw[((l * 32 + i) * 64 + o) * 3 + k, c, j].item()
        assert kc[j * 64 + c, l * 6208 + 6144 + o] == bw[l * 64 + o, c, j].item()
    kc0 = S["LB0_KC_W"].reshape(192, 24832)   # block 0: panel order
    kw0 = W["lvc_blocks.0.kernel_predictor.kernel_conv.weight"]
    for _ in range(100):
        l, i, o, k, c, j = (int(rng.integers(n)) for n in (4, 32, 64, 3, 64, 3))
        assert kc0[j * 64 + c, l * 6208 + ((k * 8 + i // 4) * 64 + o) * 4 + i % 4] == kw0[((l * 32 + i) * 64 + o) * 3 + k, c, j].item()
    up = S["LB2_UP_W"].reshape(8, 32, 32)
    assert up[5, 3, 7] == W["lvc_blocks.2.upsample.weight"][3, 7, 5].item()
    cw = S["LB0_CONV_W"].reshape(4, 3, 32, 32)
    assert cw[2, 1, 5, 9] == W["lvc_blocks.0.convs.2.weight"][9, 5, 1].item()
    # header: magic/version/section table consistent, 256-byte aligned sections
    blob = Wt.pack_state_dict(sd)
    hdr = blob.view(np.uint64)
    assert hdr[0] == Wt.BLOB_MAGIC and hdr[2] == len(Wt.SECTION_NAMES)
    assert all(int(hdr[3 + 2 * s]) % 64 == 0 for s in range(len(Wt.SECTION_NAMES)))
    # weight-norm fold equals torch's own parametrisation
    v, g = sd["final_conv.0.weight_v"], sd["final_conv.0.weight_g"]
    assert torch.allclose(W["final_conv.0.weight"], torch._weight_norm(v, g, 0), atol=1e-7)
    with pytest.raises(KeyError):
        Wt.pack_state_dict({k: v for k, v in sd.items() if "fc_t1" not in k})


def test_state_dict_roundtrip_and_repack(synth, emu_lib):
    """load_state_dict triggers a re-pack; plain (no weight-norm) checkpoints are accepted too."""
    import fastdiff_b200 as fb
    from fastdiff_b200.synthetic import make_state_dict
    from fastdiff_b200.weights import pack_state_dict
    sd, _ = synth
    net = fb.FastDiff()
    net._lib_path = emu_lib
    net.load_state_dict(sd)
    v0 = net._weights_version()
    net.engine()
    assert net._packed_version == v0
    net.load_state_dict(make_state_dict(99))
    assert net._packed_version is None  # invalidated by the post-hook
    plain = fb.FastDiff(use_weight_norm=False)
    assert all(not k.endswith("weight_g") for k in plain.state_dict())
    assert pack_state_dict(plain.state_dict()).nbytes == pack_state_dict(sd).nbytes


def test_shard_ranges():
    from fastdiff_b200.shard import shard_range
    for B in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            r = [shard_range(B, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    assert shard_range(64, 8, 3) == (24, 32)


def test_sampler_api_errors(emu_lib, synth):
    import fastdiff_b200 as fb
    sd, _ = synth
    net = fb.FastDiff()
    net._lib_path = emu_lib
    net.load_state_dict(sd)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    with pytest.raises(AssertionError):
        fb.sampling_given_noise_schedule(net, (1, 256), dh, torch.FloatTensor([0.5]), condition=torch.zeros(1, 80, 1))
    with pytest.raises(AssertionError):
        fb.sampling_given_noise_schedule(net, (1, 1, 256), {"T": 3, "alpha": dh["alpha"]}, torch.FloatTensor([0.5]),
                                         condition=torch.zeros(1, 80, 1))
    with pytest.raises(AssertionError, match="not matched"):
        fb.sampling_given_noise_schedule(net, (1, 1, 300), dh, torch.FloatTensor([0.5]), condition=torch.zeros(1, 80, 1))


def _ref_int16(x):
    """wav / |wav|.max() (modules/FastDiff/task/FastDiff.py:110) then utils/audio.py:11-16 (wav *= 32767; astype(int16))."""
    out = []
    for w in x:
        w = w / w.abs().max()
        a = w.view(-1).cpu().float().numpy().copy()
        a *= 32767
        out.append(a.astype(np.int16))
    return np.stack(out)


def test_wav_int16_encode_matches_reference_ops_bitwise(emu_lib, synth):
    from fastdiff_b200.engine import Engine
    from fastdiff_b200.weights import pack_state_dict
    sd, _ = synth
    eng = Engine(device="cpu", lib_path=emu_lib)
    eng.load_blob(pack_state_dict(sd))
    torch.manual_seed(0)
    x = torch.randn(3, 1, 1300) * 2.5
    assert np.array_equal(eng.wav_int16(x).numpy(), _ref_int16(x))


def test_embedding_chunks_equal_single_chunk(emu_lib, synth):
    """fd_sample computes the step embeddings of up to 64 reverse steps per launch (csrc/fd_api.cu: launch_embed).  With the chunk
    size forced to 1, 2, 3 (option "emb_slots") the N = 4 sampler must give the same bits as with one chunk, through the emulated
    CUDA source: slot indexing across chunk boundaries."""
    import fastdiff_b200 as fb
    from fastdiff_b200.engine import Engine
    from fastdiff_b200.sampler import build_steps
    from fastdiff_b200.synthetic import make_inputs
    from fastdiff_b200.weights import pack_state_dict
    sd, _ = synth
    eng = Engine(device="cpu", lib_path=emu_lib)
    eng.load_blob(pack_state_dict(sd))
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    _, steps = build_steps(dh, torch.FloatTensor([3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]))
    B, Tm = 2, 3
    _, mel = make_inputs(B, Tm, 4)
    torch.manual_seed(0)
    x0 = torch.randn(B, 1, Tm * 256)
    noise = torch.randn(3, B, 1, Tm * 256)
    ref = eng.sample(x0.clone(), mel, steps, noise=noise)
    assert torch.isfinite(ref).all()
    for slots in (1, 2, 3):
        eng.set_option("emb_slots", slots)
        assert torch.equal(eng.sample(x0.clone(), mel, steps, noise=noise), ref)
    eng.set_option("emb_slots", 64)
