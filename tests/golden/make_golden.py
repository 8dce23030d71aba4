"""Generates tests/golden/*.npz from the UNMODIFIED reference imported from /root/reference (CPU, fp32).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The reference hard-codes `.cuda()` (util.py:68,217,427); `torch.Tensor.cuda` is shimmed to identity so the whole
path runs on the CPU unmodified.  Weights are the seeded synthetic state dict (fastdiff_b200/synthetic.py, seed 1234,
g_jitter 0.1) loaded into the reference model with load_state_dict, so any box can rebuild the same weights.
"""
import os
import sys

import numpy as np
import torch

torch.Tensor.cuda = lambda self, *a, **k: self
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from modules.FastDiff.module.FastDiff_model import FastDiff  # noqa: E402  (the reference)
from modules.FastDiff.module import util as rutil  # noqa: E402

from fastdiff_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N_SCHEDULES = {
    3: [9.0000e-05, 9.0000e-03, 6.0000e-01],
    4: [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01],
    6: [1.7838445955931093e-06, 2.7984189728158526e-05, 0.00043231004383414984, 0.006634317338466644, 0.09357017278671265,
        0.6000000238418579],
    8: [6.689325005027058e-07, 1.0033881153503899e-05, 0.00015496854030061513, 0.002387222135439515, 0.035597629845142365,
        0.3681158423423767, 0.4735414385795593, 0.5],
}


def main():
    torch.manual_seed(0)
    model = FastDiff().eval()
    model.load_state_dict(make_state_dict(1234, g_jitter=0.1))

    # ---- denoiser: eps + stage outputs captured with forward hooks --------------------------------------
    B, Tm = 2, 12
    x, mel = make_inputs(B, Tm, 3)
    t = torch.tensor([[7.413235], [498.0537]])
    cap = {}
    hooks = []
    for n in range(3):
        def down_hook(m, i, o, n=n):
            cap[f"down{n}"] = o.detach().clone()
        def lvc_hook(m, i, o, n=n):
            cap[f"lvc{n}"] = o.detach().clone()
        hooks.append(model.downsample[n].register_forward_hook(down_hook))
        hooks.append(model.lvc_blocks[n].register_forward_hook(lvc_hook))
        def kp_hook(m, i, o, n=n):
            cap[f"kernels{n}"] = o[0].detach().clone()
            cap[f"kbias{n}"] = o[1].detach().clone()
        hooks.append(model.lvc_blocks[n].kernel_predictor.register_forward_hook(kp_hook))
    with torch.no_grad():
        eps = model((x, mel, t))
    for h in hooks:
        h.remove()
    arrays = {"x": x, "mel": mel, "t": t, "eps": eps}
    for n in range(3):
        arrays[f"down{n}"] = cap[f"down{n}"]
        arrays[f"lvc{n}"] = cap[f"lvc{n}"]
        arrays[f"kbias{n}"] = cap[f"kbias{n}"]
        arrays[f"kernels{n}_sub"] = cap[f"kernels{n}"][:, :, ::8, ::8]  # (B,4,4,8,3,T') slice keeps the file small
    np.savez_compressed(os.path.join(OUT, "denoise_b2_t12.npz"), **{k: v.numpy() for k, v in arrays.items()})

    # ---- sampler: N=4 DDPM + DDIM, fixed CPU RNG seed, reference draw order ------------------------------
    B, Tm = 1, 6
    _, mel = make_inputs(B, Tm, 5)
    dh = rutil.compute_hyperparams_given_schedule(torch.linspace(1e-6, 0.01, 1000))
    out = {"mel": mel.numpy(), "seed": np.array(11)}
    for ddim in (False, True):
        torch.manual_seed(11)
        seq = rutil.sampling_given_noise_schedule(model, (B, 1, Tm * 256), dh, torch.FloatTensor(N_SCHEDULES[4]), condition=mel,
                                                  ddim=ddim, return_sequence=True)
        out["seq_ddim" if ddim else "seq_ddpm"] = torch.stack(seq).numpy()
    np.savez_compressed(os.path.join(OUT, "sample_n4_b1_t6.npz"), **out)

    # ---- weight-free known answers: schedule tables and noise-scale -> time-step maps ---------------------
    ka = {"alpha": dh["alpha"].numpy(), "sigma": dh["sigma"].numpy()}
    for N, s in N_SCHEDULES.items():
        b = torch.FloatTensor(s)
        a = 1 - b
        sg = b + 0
        for n in range(1, len(b)):
            a[n] *= a[n - 1]
            sg[n] *= (1 - a[n - 1]) / (1 - a[n])
        a, sg = torch.sqrt(a), torch.sqrt(sg)
        ka[f"steps_{N}"] = np.array([rutil.map_noise_scale_to_time_step(a[n], dh["alpha"]) for n in range(len(b))], dtype=np.float64)
        ka[f"alpha_infer_{N}"] = a.numpy()
        ka[f"sigma_infer_{N}"] = sg.numpy()
    emb = rutil.calc_diffusion_step_embedding(torch.tensor([[0.0], [7.413235], [498.0537], [999.0]]), 128)
    ka["embed_in"] = emb.numpy()
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **ka)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
