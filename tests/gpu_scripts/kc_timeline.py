"""kernel_conv GEMM role timeline on a B200 (NOT collected by pytest).  Needs a build with -DKC_TIMELINE=1:
    FD_NVCC_EXTRA="-DKC_TIMELINE=1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)"; python tests/gpu_scripts/kc_timeline.py
CTA 0 (leader of cluster 0), config 2: clock64 stamps of the TMA producer, the MMA issuer and one epilogue warp for the first 32 items."""
import sys

import torch

sys.path.insert(0, ".")
import fastdiff_b200 as fb  # noqa: E402
from fastdiff_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

net = fb.FastDiff().cuda().eval()
net.load_state_dict(make_state_dict(1234))
B, Tm = 8, 861
x, mel = make_inputs(B, Tm, 3)
t = torch.full((B, 1), 74.99)
eng = net.engine()
eng.set_option("overlap", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(3):
    net((x.cuda(), mel.cuda(), t.cuda()))
torch.cuda.synchronize()
NI = 32
tl = eng.debug_read("kc_timeline", B, Tm).cpu().reshape(3, NI, 8).double()
names = {0: ["top", "empty ok a0", "a1", "a2"], 1: ["top", "tempty ok", "full a0", "full a1", "full a2", "committed"],
         2: ["top", "tfull ok", "ld0", "st0 issued", "ld1", "st1 issued", "arrived"]}
role = ["producer", "mma", "epilogue w2"]
for r in range(3):
    print(f"== {role[r]}: " + " | ".join(names[r]))
    for n in range(NI):
        print(f"  item {n:2d}: " + " ".join(f"{int(tl[r, n, k]):8d}" for k in range(len(names[r]))))
m = tl[1, :, 5]
per = [(m[n + 1] - m[n]).item() for n in range(8, NI - 1)]
print("steady-state cycles per item (MMA commits):", sum(per) / len(per))
try:
    ut = eng.debug_read("ut_timeline", B, Tm).cpu().reshape(24, 8).double()
    print("== k_upsample_tc<4, POUT> CTA 0: top | loads ok | split+sync | mma issued | mma done | epilogue done | end sync")
    for n in range(12):
        print(f"  tile {n:2d}: " + " ".join(f"{int(ut[n, k]):8d}" for k in range(7)))
except Exception as e:  # built without -DUT_TIMELINE
    print("no ut_timeline:", e)
