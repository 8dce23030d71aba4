"""The shipped CUDA source (SIMT kernels + host orchestration, compiled for the CPU fibre emulator in tests/cudaemu) against the
oracle at BASELINE.json configs[0] -- single 1 s mel (T' = 86), N = 4 -- and at ragged shapes; no GPU needed.  The tensor-core
kernels cannot be emulated: they are covered by tests/test_gpu_parity.py against the same oracle and against this FFMA path."""
import pytest
import torch

N4 = [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01]


def _net(sd, emu_lib):
    import fastdiff_b200 as fb
    net = fb.FastDiff().eval()
    net._lib_path = emu_lib
    net.load_state_dict(sd)
    return net


@pytest.mark.parametrize("B,Tm", [(1, 86), (2, 33), (3, 1), (1, 7)])
def test_emulated_denoiser_vs_oracle(synth, emu_lib, B, Tm):
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    sd, W = synth
    net = _net(sd, emu_lib)
    x, mel = make_inputs(B, Tm, 3)
    t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
    ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    eps = net((x, mel, t))
    assert (eps - ref).abs().max() < 5e-5          # the stated fp32 tolerance (eps rms ~1.3)
    eng = net.engine()
    L = Tm * 256
    for n, T in enumerate((L // 4, L // 32, L // 256)):
        assert (eng.debug_read(f"down{n}", B, Tm).reshape(B, 32, T) - inter[f"down{n}"]).abs().max() < 1e-5
    assert (eng.debug_read("lvc2", B, Tm).reshape(B, 32, L) - inter["lvc2"]).abs().max() < 1e-4


@pytest.mark.parametrize("ddim", [False, True])
def test_emulated_sampler_config0_vs_oracle(synth, emu_lib, ddim):
    """configs[0]: 1 s, N = 4, the reference's RNG stream; every intermediate x_t of the reverse loop (return_sequence)."""
    import fastdiff_b200 as fb
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    sd, W = synth
    net = _net(sd, emu_lib)
    B, Tm = 1, 86
    _, mel = make_inputs(B, Tm, 5)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    torch.manual_seed(11)
    ref = O.sample(W, (B, 1, Tm * 256), dh, torch.FloatTensor(N4), mel, ddim=ddim, return_sequence=True)
    torch.manual_seed(11)
    got = fb.sampling_given_noise_schedule(net, (B, 1, Tm * 256), dh, torch.FloatTensor(N4), condition=mel, ddim=ddim,
                                           return_sequence=True)
    assert len(got) == len(ref) == 5
    assert torch.equal(got[0], ref[0])
    for i in range(1, 5):
        assert (got[i] - ref[i]).abs().max() < 5e-4, i      # end-to-end tolerance at x rms ~3.4


@pytest.mark.parametrize("B,Tm", [(2, 33), (3, 1), (2, 129)])
def test_emulated_tensor_core_mode_vs_oracle(synth, emu_lib, B, Tm):
    """Mode tc_3xf16 on the CPU, every kernel of the default path: k_kp_hidden_tc, k_lvc_layer_h, the CTA-pair k_kc_gemm_tc2
    (kind::f16; 2-SM TMA boxes, cta_group::2 MMA, multicast commit, remote arrives) and k_dblock0_tc, k_upsample_tc (kind::tf32) run
    on a functional model of tcgen05 / TMEM / mbarrier / TMA / cp.async.bulk (tests/cudaemu/tcemu.h: real descriptors, SWIZZLE_128B
    with absolute-address XOR, shifted start addresses, fp32 accumulation).  Checks the DBlocks, the predicted kernels, every LVC
    block and eps against the oracle -- the same assertions the GPU test makes -- and the GEMM's fp16-piece output against an
    independent statement of that layout (FFMA GEMM + k_emu_kern_to_pieces, option emu_gemm_tc = 0)."""
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    import torch.nn.functional as F
    sd, W = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    x, mel = make_inputs(B, Tm, 4)
    t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
    ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    eps = net((x, mel, t))
    eng = net.engine()
    assert eng.get_mode() == 3
    assert (eps - ref).abs().max() < 5e-5
    L = Tm * 256
    for n, T in enumerate((L // 4, L // 32, L // 256)):          # down0 comes from the tensor-core DBlock 0
        assert (eng.debug_read(f"down{n}", B, Tm).reshape(B, 32, T) - inter[f"down{n}"]).abs().max() < 2e-5
    e = inter["embed"]
    for n in range(3):
        p = f"lvc_blocks.{n}"
        noise = F.linear(e, W[f"{p}.fc_t.weight"], W[f"{p}.fc_t.bias"]).unsqueeze(-1)
        k, bb = O.kernel_predictor(W, f"{p}.kernel_predictor", mel + noise)
        gk = eng.debug_read(f"kernels{n}", B, Tm).reshape(k.shape)      # blocks 1, 2: decoded from the fp16 pieces
        gb = eng.debug_read(f"kbias{n}", B, Tm).reshape(bb.shape)
        assert (gk - k).abs().max() < 4e-5 and (gb - bb).abs().max() < 4e-5
    assert (eng.debug_read("lvc2", B, Tm).reshape(B, 32, Tm * 256) - inter["lvc2"]).abs().max() < 2e-4
    eng.set_option("stop_after", 4)
    net((x, mel, t))
    assert (eng.debug_read("lvc1", B, Tm).reshape(B, 32, Tm * 64) - inter["lvc1"]).abs().max() < 2e-4
    eng.set_option("stop_after", 99)
    k_tc = [eng.debug_read(f"kernels{n}", B, Tm).clone() for n in range(3)]
    eng.set_option("emu_gemm_tc", 0)
    eps_conv = net((x, mel, t))
    eng.set_option("emu_gemm_tc", 1)
    for n in range(3):
        assert (eng.debug_read(f"kernels{n}", B, Tm) - k_tc[n]).abs().max() < 2e-5
    assert (eps_conv - eps).abs().max() < 5e-5
    net.mode = "fp32_simt"
    assert (net((x, mel, t)) - eps).abs().max() < 5e-5


def test_emulated_tensor_core_mode_is_batch_and_tiling_independent(synth, emu_lib):
    """The carried halo rows (descending tile walk) and the second-MMA-pass fallback give the same bits: an item evaluated alone
    (different chunking of the tile walk) equals the same item inside a batch, bitwise -- on the tensor-core model."""
    from fastdiff_b200.synthetic import make_inputs
    sd, _ = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    B, Tm = 3, 40
    x, mel = make_inputs(B, Tm, 9)
    t = torch.full((B, 1), 74.99228)
    full = net((x, mel, t))
    for lo, hi in ((0, 1), (1, 3)):
        assert torch.equal(net((x[lo:hi], mel[lo:hi], t[lo:hi])), full[lo:hi])


def test_emulated_tensor_core_sampler_vs_oracle(synth, emu_lib):
    """N = 4 reverse loop in mode tc_3xf16 on the model, the reference's RNG stream, every intermediate x_t."""
    import fastdiff_b200 as fb
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    sd, W = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    B, Tm = 2, 20
    _, mel = make_inputs(B, Tm, 5)
    dh = fb.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    torch.manual_seed(11)
    ref = O.sample(W, (B, 1, Tm * 256), dh, torch.FloatTensor(N4), mel, return_sequence=True)
    torch.manual_seed(11)
    got = fb.sampling_given_noise_schedule(net, (B, 1, Tm * 256), dh, torch.FloatTensor(N4), condition=mel, return_sequence=True)
    assert net.engine().get_mode() == 3
    for i in range(1, 5):
        assert (got[i] - ref[i]).abs().max() < 5e-4, i


def test_emulated_kernel_conv_gemm_forms_agree(synth, emu_lib):
    """k_kc_gemm_tc2, resident-frame-tile form (option `kc_res` = 1: contiguous item range per CTA pair, one 130-row tile per piece serving the
    three im2col taps through shifted descriptor starts, two tile buffers, 4-stage weight ring, early accumulator release) against the
    default whole-stage ring (`kc_res` = 0): the same MMAs in the same order, so the same bits.  `kc_clusters` = 1 makes ONE pair walk all three frame
    tiles (a tile buffer is reused: the producer waits for the MMA issuer's release), 3 splits the range inside tiles."""
    from fastdiff_b200.synthetic import make_inputs
    sd, _ = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    B, Tm = 1, 129                      # one frame tile per block: three tiles in all (the GPU test walks twelve)
    x, mel = make_inputs(B, Tm, 9)
    eng = net.engine()
    eng.set_option("stop_after", 1)     # embedding + kernel predictor + GEMM: the GEMM's output is what is compared

    def run(res, clusters, t):
        eng.set_option("kc_res", res)
        eng.set_option("kc_clusters", clusters)
        net((x, mel, t))
        return [eng.debug_read(f"kernels{n}", B, Tm).clone() for n in range(3)] + [eng.debug_read(f"kbias{n}", B, Tm).clone() for n in range(3)]

    # every variant runs at its OWN diffusion step, followed by the whole-stage ring at the same step: an item a variant skipped would
    # keep the previous step's value and differ
    for i, (res, clusters) in enumerate(((1, 1), (1, 3))):
        t = torch.tensor([[7.413235 + 40.0 * i]])
        got = run(res, clusters, t)
        want = run(0, 0, t)
        for a, b in zip(got, want):
            assert torch.equal(a, b), (res, clusters)
        if i:
            assert not torch.equal(got[1], prev[1])
        prev = got
    eng.set_option("kc_res", 0)
    eng.set_option("kc_clusters", 0)
    eng.set_option("stop_after", 99)


@pytest.mark.parametrize("B,Tm", [(1, 2), (1, 9), (3, 17), (1, 65), (5, 7), (1, 113), (1, 225)])
def test_emulated_tensor_core_mode_ragged_shapes(synth, emu_lib, B, Tm):
    """Tile-boundary shapes of the tensor-core kernels (kernel-predictor tiles of 112 frames, LVC tiles of 128 steps with 64- and
    256-step frames, GEMM tiles of 256 frames): eps against the oracle on the model.  The emulation launches few CTAs, so every
    group walks several tiles (carried halo rows, kernel reuse, ring wrap-around) even at these sizes."""
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    sd, W = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    x, mel = make_inputs(B, Tm, 10 + Tm)
    t = torch.linspace(3.0, 900.0, B).reshape(B, 1)
    assert (net((x, mel, t)) - O.denoise(W, x, mel, t)).abs().max() < 5e-5


@pytest.mark.parametrize("B,Tm", [(2, 33), (1, 129), (3, 17)])
def test_emulated_block0_tensor_core(synth, emu_lib, B, Tm):
    """LVC block 0 (hop 8) on the tensor-core model in swapped-operand form (default of mode tc_3xf16 since round 2) -- k_lvc_layer_b0h
    (kernels of a frame pair as the M = 128 operand, 16 step columns, 3-slot kernel ring across tiles, second MMA pass for the halo rows,
    tanh/sigmoid exchange through shared memory), fed by the GEMM writing block 0 as fp16 pieces itself (k_kc_gemm_tc2<true, 16, true> on
    the image-ordered weight rows LB0_KCT_F16P).  Block-0 output, the decoded pieces and eps against the oracle and the SIMT block-0 path."""
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    import torch.nn.functional as F
    sd, W = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    x, mel = make_inputs(B, Tm, 21)
    t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
    ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    eng = net.engine()
    eps = net((x, mel, t))
    assert (eps - ref).abs().max() < 5e-5
    noise = F.linear(inter["embed"], W["lvc_blocks.0.fc_t.weight"], W["lvc_blocks.0.fc_t.bias"]).unsqueeze(-1)
    k, bb = O.kernel_predictor(W, "lvc_blocks.0.kernel_predictor", mel + noise)
    assert (eng.debug_read("kernels0", B, Tm).reshape(k.shape) - k).abs().max() < 4e-5      # decoded from the pieces
    assert (eng.debug_read("kbias0", B, Tm).reshape(bb.shape) - bb).abs().max() < 4e-5
    eng.set_option("stop_after", 3)
    net((x, mel, t))
    assert (eng.debug_read("lvc0", B, Tm).reshape(B, 32, Tm * 8) - inter["lvc0"]).abs().max() < 1e-4
    eng.set_option("stop_after", 99)
    if B > 1:      # an item evaluated alone = the same item inside the batch, bitwise (per-tile work is independent of the tile walk)
        assert torch.equal(net((x[1:2], mel[1:2], t[1:2])), eps[1:2])
    eng.set_option("tc_b0", 0)                      # SIMT block 0 (fp32 panel image of the kernels)
    eps_simt_b0 = net((x, mel, t))
    eng.set_option("tc_b0", 1)
    assert (eps_simt_b0 - ref).abs().max() < 5e-5 and (eps_simt_b0 - eps).abs().max() < 5e-5
    net.mode = "fp32_simt"                          # the option only applies to mode tc_3xf16
    assert (net((x, mel, t)) - ref).abs().max() < 5e-5


@pytest.mark.parametrize("B,Tm", [(1, 3), (2, 9), (3, 17), (1, 40), (2, 35)])
def test_emulated_piece_row_path_vs_row_path(synth, emu_lib, B, Tm):
    """Blocks 1, 2 on the piece-row protocol (k_lvc_p + k_upsample_tc<R, true>, the default; fd_kernels_lvcp.cuh) against the oracle and
    against the fp32-row kernels they replace (k_lvc_layer_h, option lvc_p = 0): lvc1 / lvc2 stage outputs and eps.  The two paths
    differ by the 22-bit state carry only (~2e-6 on eps).  (1,3): single-tile CTAs; (2,9), (3,17): odd T' -> a half tile ends every
    item of block 1; (1,40): chunks of several tiles (carried rows, kernel reuse across the two tiles of a hop-256 frame, ring wrap); (2,35): an
    utterance boundary inside a 32-frame sub-chunk of the kernel_conv GEMM epilogue (its gap path: T' + 2 > 34) and rows past the end."""
    from fastdiff_b200.synthetic import make_inputs
    from oracle import fastdiff_oracle as O
    sd, W = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    x, mel = make_inputs(B, Tm, 61)
    t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
    ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    eng = net.engine()
    out = {}
    for lp in (1, 0):
        eng.set_option("lvc_p", lp)
        out[lp] = net((x, mel, t))
        assert (out[lp] - ref).abs().max() < 5e-5
        assert (eng.debug_read("lvc2", B, Tm).reshape(B, 32, Tm * 256) - inter["lvc2"]).abs().max() < 2e-4
        eng.set_option("stop_after", 4)
        net((x, mel, t))
        assert (eng.debug_read("lvc1", B, Tm).reshape(B, 32, Tm * 64) - inter["lvc1"]).abs().max() < 2e-4
        eng.set_option("stop_after", 99)
    eng.set_option("lvc_p", 1)
    assert (out[1] - out[0]).abs().max() < 1e-5
    assert not eng.check_saturation()
    if B > 1:      # an item alone = the same item inside the batch, bitwise
        assert torch.equal(net((x[1:2], mel[1:2], t[1:2])), out[1][1:2])


def test_emulated_saturation_flag(synth, emu_lib):
    """The range guard of mode tc_3xf16: operands are fp16 pieces of 16 x activation and saturate at 65504 instead of overflowing; every
    kernel that forms pieces raises a sticky device flag (fd_check_saturation) -- ADVICE round 1: the mode must not fail silently."""
    from fastdiff_b200.synthetic import make_inputs
    sd, _ = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    x, mel = make_inputs(1, 3, 62)
    t = torch.tensor([[74.99228]])
    eng = net.engine()
    net((x, mel, t))
    assert not eng.check_saturation()
    out = net((x * 3e4, mel, t))                    # |activation| * 16 >> 65504
    assert torch.isfinite(out).all()
    assert eng.check_saturation()                   # raised ... and cleared by the call
    assert not eng.check_saturation()


def test_emulated_kernels_under_all_async_schedules(synth, emu_lib):
    """cp.async.bulk global->shared copies and tcgen05.mma are asynchronous.  The model can land a copy's bytes at issue (adversarial for
    a target something still reads) or when its mbarrier is first polled (adversarial for a target read or written before the wait),
    and can run the MMAs at issue or when their commit barrier is first polled (adversarial for an operand tile refilled, or an
    accumulator read, too early).  The default path (warp-specialised k_lvc_p with its 3 + 2 + 2 stage rings, ring-fed block 0) and the
    fp32-row kernels behind the options must give the same bits under all of them."""
    import ctypes
    from fastdiff_b200.synthetic import make_inputs
    sd, _ = synth
    net = _net(sd, emu_lib)
    net.mode = "tc_3xf16"
    eng = net.engine()
    model = ctypes.CDLL(emu_lib)
    x, mel = make_inputs(2, 12, 4)
    t = torch.tensor([[7.413235], [498.0537]])
    variants = {"default": {}, "rows": {"lvc_p": 0}, "simt_b0": {"tc_b0": 0}}
    out = {}
    try:
        for late in (0, 1, 2, 3):       # early / late loads / late MMAs / everything late
            model.cudaemu_set_bulk_late(late)
            for name, opts in variants.items():
                eng.set_option("lvc_p", opts.get("lvc_p", 1))
                eng.set_option("tc_b0", opts.get("tc_b0", 1))
                out[(late, name)] = net((x, mel, t))
    finally:
        model.cudaemu_set_bulk_late(0)
        eng.set_option("lvc_p", 1)
        eng.set_option("tc_b0", 1)
    for name in variants:
        for late in (1, 2, 3):
            assert torch.equal(out[(0, name)], out[(late, name)]), (name, late)
