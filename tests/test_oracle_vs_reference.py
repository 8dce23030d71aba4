"""Pins the oracle to the UNMODIFIED reference imported from /root/reference (present in the build container only;
skipped on the GPU box, where tests/golden/*.npz -- outputs of this same reference -- take over)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules", "FastDiff")), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference hard-codes .cuda() (util.py:68,217,427)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from modules.FastDiff.module import util as rutil
        from modules.FastDiff.module.FastDiff_model import FastDiff
    return FastDiff, rutil


def test_denoiser_matches_reference(ref, synth):
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    FastDiff, _ = ref
    sd, W = synth
    m = FastDiff().eval()
    m.load_state_dict(sd)
    for B, Tm, seed in ((2, 12, 0), (1, 1, 1), (3, 5, 2)):
        x, mel = make_inputs(B, Tm, seed)
        t = torch.tensor([7.413235, 498.0537, 23.46759][:B]).reshape(B, 1)
        with torch.no_grad():
            r = m((x, mel, t))
        o = O.denoise(W, x, mel, t)
        assert (r - o).abs().max() < 2e-5


def test_lvc_closed_form_matches_reference(ref):
    from oracle import fastdiff_oracle as O
    FastDiff, _ = ref
    blk = FastDiff().lvc_blocks[2]
    torch.manual_seed(0)
    for hop, Tm in ((256, 3), (64, 5), (8, 7)):
        x = torch.randn(2, 32, hop * Tm)
        k = torch.randn(2, 32, 64, 3, Tm)
        b = torch.randn(2, 64, Tm)
        r = blk.location_variable_convolution(x, k, b, 1, hop)
        assert (r - O.lvc(x, k, b, hop)).abs().max() < 5e-5


def test_sampler_matches_reference(ref, synth):
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    FastDiff, rutil = ref
    sd, W = synth
    m = FastDiff().eval()
    m.load_state_dict(sd)
    _, mel = make_inputs(1, 4, 5)
    dh = rutil.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    sched = [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]
    for ddim in (False, True):
        torch.manual_seed(5)
        r = rutil.sampling_given_noise_schedule(m, (1, 1, 1024), dh, torch.FloatTensor(sched), condition=mel, ddim=ddim, return_sequence=True)
        torch.manual_seed(5)
        o = O.sample(W, (1, 1, 1024), dh, torch.FloatTensor(sched), mel, ddim=ddim, return_sequence=True)
        assert len(r) == len(o)
        for a, b in zip(r, o):
            assert (a - b).abs().max() < 1e-4


def test_host_tables_bit_identical_to_reference(ref):
    """The product's vectorised host prologue returns the very floats the reference's python loops return."""
    import fastdiff_b200 as fb
    _, rutil = ref
    beta = torch.linspace(1e-6, 0.01, 1000)
    a, b = rutil.compute_hyperparams_given_schedule(beta.clone()), fb.compute_hyperparams_given_schedule(beta.clone())
    assert torch.equal(a["alpha"], b["alpha"]) and torch.equal(a["sigma"], b["sigma"]) and a["T"] == b["T"]
    for al in (0.5, 0.9999, 0.05, 1.0, 0.0813796):
        assert rutil.map_noise_scale_to_time_step(torch.tensor(al), a["alpha"]) == fb.map_noise_scale_to_time_step(torch.tensor(al), b["alpha"])
    st = torch.tensor([[0.0], [3.5], [999.0]])
    assert torch.equal(rutil.calc_diffusion_step_embedding(st, 128), fb.calc_diffusion_step_embedding(st, 128))


def test_state_dict_keys_match_reference(ref):
    import fastdiff_b200 as fb
    FastDiff, _ = ref
    r, m = FastDiff().state_dict(), fb.FastDiff().state_dict()
    assert list(r.keys()) == list(m.keys())
    assert all(r[k].shape == m[k].shape for k in r)
    # a reference state dict loads strictly, and the packer accepts it
    fb.FastDiff().load_state_dict(r, strict=True)
    from fastdiff_b200.weights import pack_state_dict
    assert pack_state_dict(r).nbytes > 60e6
