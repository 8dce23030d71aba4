"""GPU tests of fd_reverse_update (collected last): the stand-alone reverse-step update against the reference's eager
operation sequence evaluated on the CPU (IEEE add / mul / div: the same bits on both), and against the fused loop of fd_sample."""
import pytest
import torch

gpu = pytest.mark.gpu


def _step():
    from fastdiff_b200._lib import fd_step
    st = fd_step()
    st.t, st.coef_eps, st.div, st.sigma, st.add_noise = 74.99, 0.151118, 0.98723, 0.051044, 1
    st.c1, st.c2, st.c3 = 1.01, -0.16, 0.12
    return st


@gpu
@pytest.mark.parametrize("n", [3, 4099, 8 * 220416])
def test_reverse_update_bitwise_vs_eager_ops(cuda_lib, n):
    from fastdiff_b200.engine import Engine
    eng = Engine(device="cuda:0")
    g = torch.Generator().manual_seed(n)
    x0, eps, z = (torch.randn(n, generator=g) * s for s in (3.0, 1.3, 1.0))
    st = _step()
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    want = x0.clone()
    want -= f(st.coef_eps) * eps          # util.py:226
    want /= f(st.div)                     # util.py:227
    want = want + f(st.sigma) * z         # util.py:229
    x = x0.cuda()
    eng.reverse_update(x, eps.cuda(), st, z=z.cuda())
    assert torch.equal(x.cpu(), want)
    want = f(st.c1) * x0 + f(st.c2) * eps + f(st.c3) * eps      # util.py:224
    x = x0.cuda()
    eng.reverse_update(x, eps.cuda(), st, ddim=True)
    assert torch.equal(x.cpu(), want)


@gpu
@pytest.mark.parametrize("ddim", [False, True])
def test_fused_loop_equals_denoise_plus_reverse_update_gpu(synth, cuda_lib, ddim):
    import fastdiff_b200 as fb
    from fastdiff_b200.sampler import build_steps
    from fastdiff_b200.synthetic import make_inputs
    sd, _ = synth
    net = fb.FastDiff().to("cuda:0").eval()
    net.load_state_dict(sd)
    eng = net.engine()
    B, Tm = 2, 9
    x0, mel = make_inputs(B, Tm, 8)
    x0, mel = x0.cuda(), mel.cuda()
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    _, steps = build_steps(dh, torch.FloatTensor([3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]), ddim)
    for noise, seed in ((torch.randn(3, B, 1, Tm * 256).cuda(), 0), (None, 99)):
        a = eng.sample(x0.clone(), mel, steps, noise=noise, seed=seed, ddim=ddim)
        b = x0.clone()
        draw = 0
        for st in steps:
            eps = eng.denoise(b, mel, torch.full((B,), st.t))
            z = None
            if st.add_noise and not ddim:
                z = noise[draw] if noise is not None else None
                draw += 1
            eng.reverse_update(b, eps, st, z=z, ddim=ddim, seed=seed, draw=draw)
        assert torch.equal(a, b)
