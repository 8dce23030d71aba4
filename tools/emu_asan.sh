#!/bin/sh
# AddressSanitizer pass over the CUDA source under the CPU emulator (TEST INFRASTRUCTURE): every kernel of the default mode
# (tensor-core model) and of the FFMA mode at two shapes, plus the optional kernels; catches out-of-bounds global / shared-memory accesses.
#   sh tools/emu_asan.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/libfastdiff_emu_asan.so
g++ -O1 -g -std=c++17 -fPIC -shared -DFD_EMU -ffp-contract=off -Wno-psabi -fsanitize=address -fno-omit-frame-pointer -x c++ \
    -I"$ROOT/tests/cudaemu" -o "$OUT" "$ROOT/fastdiff_b200/csrc/fd_api.cu" "$ROOT/tests/cudaemu/cudaemu.cpp" -lpthread
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 \
python - "$ROOT" "$OUT" <<'PY'
import sys, torch
root, lib = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
import fastdiff_b200 as fb
from fastdiff_b200.synthetic import make_state_dict, make_inputs
from oracle import fastdiff_oracle as O
sd = make_state_dict(1234, g_jitter=0.1); W = O.fold_weight_norm(sd)
net = fb.FastDiff().eval(); net._lib_path = lib; net.load_state_dict(sd)
for mode in ("tc_3xf16", "fp32_simt"):
    net.mode = mode
    for B, Tm in ((1, 5), (2, 33)):
        x, mel = make_inputs(B, Tm, 3); t = torch.tensor([7.413235, 498.0537][:B]).reshape(B, 1)
        err = (net((x, mel, t)) - O.denoise(W, x, mel, t)).abs().max().item()
        print(mode, B, Tm, "max|eps - oracle| =", err, flush=True)
        assert err < 5e-5
# the kernels that are OFF by default, one at a time and all together, and the stand-alone reverse-step update
net.mode = "tc_3xf16"
eng = net.engine()
x, mel = make_inputs(2, 33, 3); t = torch.tensor([[7.413235], [498.0537]])
ref = O.denoise(W, x, mel, t)
opts = ("tc_b0", "b2_skipbuf", "kc_stage", "lvc_pipe")
for on in opts + ("all",):          # "all" includes lvc_pipe + b2_skipbuf = k_lvc_layer_p<true> for layers 1..3
    for k in opts:
        eng.set_option(k, 1 if on == "all" or k == on else 0)
    err = (net((x, mel, t)) - ref).abs().max().item()
    print("option", on, "max|eps - oracle| =", err, flush=True)
    assert err < 5e-5
for k in opts:
    eng.set_option(k, 0)
from fastdiff_b200._lib import fd_step
st = fd_step(); st.coef_eps, st.div, st.sigma, st.add_noise = 0.15, 0.98, 0.05, 1
xx = torch.randn(4099); eng.reverse_update(xx, torch.randn(4099), st, z=torch.randn(4099)); eng.reverse_update(xx, torch.randn(4099), st, seed=3, draw=1)
print("ASan pass complete: no errors")
PY
