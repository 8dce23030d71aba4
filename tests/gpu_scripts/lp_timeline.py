"""k_lvc_p role timeline on a B200 (NOT collected by pytest).  Needs a build with -DLP_TIMELINE=1:
    FD_NVCC_EXTRA="-DLP_TIMELINE=1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)"; python tests/gpu_scripts/lp_timeline.py
Prints, for CTA 0 of the block-2 layer with dilation 9, the clock64 stamps of every role per tile (cycles since the first stamp)."""
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
for _ in range(3):
    net((x.cuda(), mel.cuda(), t.cuda()))
torch.cuda.synchronize()
NT = 24
tl = net.engine().debug_read("lp_timeline", B, Tm).cpu().reshape(4, NT, 8).double()
names = {
    0: ["top", "a_free ok", "loads issued"],
    1: ["top", "a_full+cacc_free ok", "conv issued", "lvc top", "y_full ok", "w_full+lacc_free ok", "lvc issued"],
    2: ["top", "cacc_full ok", "ld + lacc_full(n-2) ok", "rows emitted", "a_full ok", "y_full arrive"],
    3: ["top", "a_full ok", "z rebuilt", "lacc_full ok", "tmem read", "gate math done", "stored + a_free"],
}
role = ["loader", "mma", "conv-epi", "gate-epi"]
for r in range(4):
    print(f"== {role[r]}: " + " | ".join(names[r]))
    for n in range(NT):
        row = [int(tl[r, n, k]) for k in range(len(names[r]))]
        print(f"  tile {n:2d}: " + " ".join(f"{v:7d}" for v in row))
# steady-state period from the gate epilogue's last stamp
g = tl[3, :, 6]
per = [(g[n + 1] - g[n]).item() for n in range(8, NT - 1) if g[n] > 0 and g[n + 1] > 0]
print("steady-state cycles per tile (gate epilogue, tiles 8..):", sum(per) / max(len(per), 1))
