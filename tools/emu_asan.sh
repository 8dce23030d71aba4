#!/bin/sh
# AddressSanitizer pass over the CUDA source under the CPU emulator (TEST INFRASTRUCTURE): every kernel of the default mode
# (tensor-core model) and of the FFMA mode at two shapes; catches out-of-bounds global / shared-memory accesses.
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
print("ASan pass complete: no errors")
PY
