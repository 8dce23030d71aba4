"""fp16-piece operands of mode tc_3xf16: the packer's split, the device-side conversion (software definition compiled from the
shipped header) and the blob layouts documented in csrc/fd_blob.h -- no GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fastdiff_b200", "csrc")


@pytest.fixture(scope="module")
def conv_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("f16")
    src = d / "conv.cpp"
    src.write_text('''#define FD_EMU
#include "fd_common.cuh"
extern "C" void to_f16(const float* in, uint16_t* out, int n) { for (int i = 0; i < n; ++i) out[i] = fd::f16_bits_rn_soft(in[i]); }
extern "C" void from_f16(const uint16_t* in, float* out, int n) { for (int i = 0; i < n; ++i) out[i] = fd::f16_bits_to_float_soft(in[i]); }
extern "C" void split(const float* in, float scale, uint16_t* hi, uint16_t* lo, int n) { for (int i = 0; i < n; ++i) fd::f16_split(in[i], scale, hi[i], lo[i]); }
''')
    so = d / "conv.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + CSRC, "-I" + os.path.join(ROOT, "tests", "cudaemu"),
                    "-o", str(so), str(src), "-lpthread"], check=True)
    return ctypes.CDLL(str(so))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_soft_f16_conversion_is_ieee_round_to_nearest_even(conv_lib):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * s for s in (1, 1e-3, 1e-5, 1e-7, 1e4, 1e5)])
    allh = np.arange(65536, dtype=np.uint16)
    fh = allh.view(np.float16).astype(np.float32)
    fin = np.sort(fh[np.isfinite(fh)])
    mid = ((fin[:-1].astype(np.float64) + fin[1:].astype(np.float64)) / 2).astype(np.float32)   # exact ties
    x = np.concatenate([x, fin, mid, np.array([65504, 65519.99, 65520, 65536, 2.9802322e-8, 3e-8, -0.0, 0.0], dtype=np.float32)])
    out = np.zeros(x.size, dtype=np.uint16)
    conv_lib.to_f16(_p(x), _p(out), x.size)
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    assert np.array_equal(out, ref)
    back = np.zeros(65536, dtype=np.float32)
    conv_lib.from_f16(_p(allh), _p(back), 65536)
    m = np.isfinite(fh)
    assert np.array_equal(back[m], fh[m])


def test_device_split_equals_packer_split_and_is_22_bit(conv_lib):
    from fastdiff_b200.weights import f16_split
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-9, 5, 100000))).astype(np.float32)
    for scale in (1.0, 16.0, 64.0, 131072.0):
        hi = np.zeros(x.size, dtype=np.uint16); lo = np.zeros(x.size, dtype=np.uint16)
        conv_lib.split(_p(x), ctypes.c_float(scale), _p(hi), _p(lo), x.size)
        h2, l2 = f16_split(x, scale)
        assert np.array_equal(hi, h2) and np.array_equal(lo, l2)
        s = np.clip(x.astype(np.float64) * scale, -65504, 65504)
        rec = hi.view(np.float16).astype(np.float64) + lo.view(np.float16).astype(np.float64)
        assert np.isfinite(rec).all()                                   # saturating: never inf
        inr = np.abs(s) < 65504
        err = np.abs(rec - s)[inr]
        # 22 significant bits where lo is a normal fp16 (|s| >= 2^-3), absolute 2^-25 below
        assert (err <= np.maximum(np.abs(s[inr]) * 2.0 ** -21.9, 2.0 ** -25)).all()


def test_f16_scale_is_power_of_two_with_headroom():
    from fastdiff_b200.weights import f16_scale
    for m in (1e-4, 0.0721, 1.0, 3.9, 1000.0):
        sc = f16_scale(np.array([m, -m / 3], dtype=np.float32))
        assert np.log2(sc) == np.round(np.log2(sc))
        assert 8192 < m * sc <= 16384
    assert f16_scale(np.zeros(4, dtype=np.float32)) == 1.0


def test_blob_f16_sections_decode_to_the_folded_weights(synth):
    """LBn_KCT_F16 / LBn_CONV_F16 / LBn_KPW_F16 layouts as documented in fd_blob.h, decoded with independent index arithmetic."""
    from fastdiff_b200.weights import build_sections
    sd, W = synth
    S = build_sections(sd)
    sc = S["SCALES16"]
    f16 = lambda u: np.asarray(u, dtype=np.uint16).view(np.float16).astype(np.float64)
    for n in range(3):
        kct = S[f"LB{n}_KC_W"].reshape(192, 24832).T.astype(np.float64)             # [n][kk]
        u = S[f"LB{n}_KCT_F16"].view(np.uint16).reshape(2, 24832, 192)
        rec = (f16(u[0]) + f16(u[1])) / sc[n]
        assert np.abs(rec - kct).max() <= np.abs(kct).max() * 2.0 ** -21
    rng = np.random.default_rng(2)
    for n in (1, 2):
        u = S[f"LB{n}_CONV_F16"].view(np.uint16).reshape(4, 3, 32, 64)             # [l][k][co][64 halves of the 128-byte row]
        for _ in range(200):
            l, k, co, ci = rng.integers(4), rng.integers(3), rng.integers(32), rng.integers(32)
            w = float(W[f"lvc_blocks.{n}.convs.{l}.weight"][co, ci, k])
            hi = f16(u[l, k, co, (((ci >> 3) ^ (co & 7)) << 3) + (ci & 7)])
            lo = f16(u[l, k, co, (((4 + (ci >> 3)) ^ (co & 7)) << 3) + (ci & 7)])
            assert abs((hi + lo) / sc[4 + 4 * n + l] - w) <= abs(w) * 2.0 ** -21 + 1e-12
    for n in range(3):
        u = S[f"LB{n}_KPW_F16"].view(np.uint16).reshape(28, 2, 64, 64)              # [slot][tile][co][64 halves]
        kp = f"lvc_blocks.{n}.kernel_predictor"
        for _ in range(200):
            j, co, ci = rng.integers(5), rng.integers(64), rng.integers(80)
            w = float(W[f"{kp}.input_conv.0.weight"][co, ci, j])
            if ci < 64:
                pos = (((ci >> 3) ^ (co & 7)) << 3) + (ci & 7)
                rec = f16(u[2 * j, 0, co, pos]) + f16(u[2 * j, 1, co, pos])
            else:
                c2 = ci - 64
                rec = f16(u[2 * j + 1, 0, co, (((c2 >> 3) ^ (co & 7)) << 3) + (c2 & 7)]) + \
                      f16(u[2 * j + 1, 0, co, (((2 + (c2 >> 3)) ^ (co & 7)) << 3) + (c2 & 7)])
            assert abs(rec / sc[16 + 8 * n] - w) <= abs(w) * 2.0 ** -21 + 1e-12
            l, j3, ci3 = rng.integers(6), rng.integers(3), rng.integers(64)
            w = float(W[f"{kp}.residual_conv.{(1, 3, 6, 8, 11, 13)[l]}.weight"][co, ci3, j3])
            pos = (((ci3 >> 3) ^ (co & 7)) << 3) + (ci3 & 7)
            rec = f16(u[10 + 3 * l + j3, 0, co, pos]) + f16(u[10 + 3 * l + j3, 1, co, pos])
            assert abs(rec / sc[17 + 8 * n + l] - w) <= abs(w) * 2.0 ** -21 + 1e-12


def test_merged_n_sections_decode_to_the_folded_weights(synth):
    """The round-2 merged-N operand images (fd_blob.h: LBn_CONV_F16M, LB2_UP_F16M, FIRST_F16U), decoded with independent index arithmetic:
    hi and lo pieces are separate ROWS of SWIZZLE_128B tiles whose 128-byte rows hold two taps side by side."""
    from fastdiff_b200.weights import build_sections
    sd, W = synth
    S = build_sections(sd)
    sc = S["SCALES16"]
    f16 = lambda u: np.asarray(u, dtype=np.uint16).view(np.float16).astype(np.float64)
    rng = np.random.default_rng(3)

    def row_val(tile, r, half):   # fp16 value `half` (0..63) of logical row r of a [rows][64 halves] SWIZZLE_128B tile
        return f16(tile[r, (((half >> 3) ^ (r & 7)) << 3) + (half & 7)])

    for n in (1, 2):   # LBn_CONV_F16M: per layer T01 [64 rows = piece * 32 + co][tap 0 | tap 1], T2 [32 rows co][hi | lo] of tap 2
        u = S[f"LB{n}_CONV_F16M"].view(np.uint16).reshape(4, 96, 64)
        for _ in range(300):
            l, k, co, ci = rng.integers(4), rng.integers(3), rng.integers(32), rng.integers(32)
            w = float(W[f"lvc_blocks.{n}.convs.{l}.weight"][co, ci, k])
            if k < 2:
                rec = row_val(u[l, :64], co, k * 32 + ci) + row_val(u[l, :64], 32 + co, k * 32 + ci)
            else:
                rec = row_val(u[l, 64:], co, ci) + row_val(u[l, 64:], co, 32 + ci)
            assert abs(rec / sc[4 + 4 * n + l] - w) <= abs(w) * 2.0 ** -21 + 1e-12
    # LB2_UP_F16M: per output phase ph, rows piece * 32 + co = [tap kk1 | tap kk1 + 4], kk1 = (ph + 2) % 4, ConvTranspose1d weight (ci, co, k)
    u = S["LB2_UP_F16M"].view(np.uint16).reshape(4, 64, 64)
    up = W["lvc_blocks.2.upsample.weight"]
    for _ in range(300):
        ph, tap, co, ci = rng.integers(4), rng.integers(2), rng.integers(32), rng.integers(32)
        w = float(up[ci, co, (ph + 2) % 4 + 4 * tap])
        rec = row_val(u[ph], co, tap * 32 + ci) + row_val(u[ph], 32 + co, tap * 32 + ci)
        assert abs(rec / sc[41] - w) <= abs(w) * 2.0 ** -21 + 1e-12
    # FIRST_F16U: rows ph * 32 + co of 32 halves: W_ph[i] = first_w[co][i - ph] (0 <= i - ph <= 6), W_ph[10] = first_b[co], else 0; hi 16 | lo 16
    u = S["FIRST_F16U"].view(np.uint16).reshape(128, 32)
    fw, fb = W["first_audio_conv.weight"][:, 0, :], W["first_audio_conv.bias"]
    for ph in range(4):
        for co in range(32):
            rec = (f16(u[ph * 32 + co, :16]) + f16(u[ph * 32 + co, 16:])) / sc[41]
            want = np.zeros(16)
            want[ph:ph + 7] = fw[co].numpy()
            want[10] = float(fb[co])
            assert np.abs(rec - want).max() <= np.abs(want).max() * 2.0 ** -20 + 1e-12
    assert sc[41] == min(sc[40], sc[41]) and np.log2(sc[41]) == np.round(np.log2(sc[41]))
