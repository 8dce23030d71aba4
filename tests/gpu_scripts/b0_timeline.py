"""k_lvc_layer_b0h phase timeline on a B200 (NOT collected by pytest).  Needs a build with -DB0_TIMELINE=1."""
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
eng.set_option("overlap", 0)
for _ in range(3):
    net((x.cuda(), mel.cuda(), t.cuda()))
torch.cuda.synchronize()
tl = eng.debug_read("b0_timeline", B, Tm).cpu().reshape(4, 8).double()
print("== k_lvc_layer_b0h CTA 0 (last launch = layer 3, dil 27): top | rows ok | pieces + sync | conv MMAs done | Y epilogue + sync | LVC MMAs done | gate + end sync")
for n in range(4):
    print(f"  tile {n}: " + " ".join(f"{int(tl[n, k]):8d}" for k in range(7)))
