# GPU scratch check (not a pytest): fp16-piece tensor-core mode vs oracle / other modes, stage by stage
import sys, torch
sys.path.insert(0, '/root/repo')
import fastdiff_b200 as fb
from fastdiff_b200.synthetic import make_state_dict, make_inputs
from oracle import fastdiff_oracle as O
import torch.nn.functional as F
sd = make_state_dict(1234, g_jitter=0.1); W = O.fold_weight_norm(sd)
net = fb.FastDiff().cuda().eval(); net.load_state_dict(sd)
MODES = ('tc_3xf16', 'tc_3xf16:tc_kp=0')
for (B, Tm) in ((1, 5), (2, 33), (3, 300)):
    x, mel = make_inputs(B, Tm, 3); t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
    eps_o, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    e = O.embed_mlp(W, t)
    for mode in MODES:
        net.mode = mode.split(':')[0]
        eng = net.engine(); eng.set_option('stop_after', 1)
        eng.set_option('tc_kp', 0 if 'tc_kp=0' in mode else 1)
        try:
            net((x.cuda(), mel.cuda(), t.cuda())); torch.cuda.synchronize()
        except Exception as ex:
            print('FAILED', mode, repr(ex), flush=True); eng.set_option('stop_after', 99); continue
        for n in range(3):
            p = f'lvc_blocks.{n}'
            noise = F.linear(e, W[f'{p}.fc_t.weight'], W[f'{p}.fc_t.bias']).unsqueeze(-1)
            k, bb = O.kernel_predictor(W, f'{p}.kernel_predictor', mel + noise)
            gk = eng.debug_read(f'kernels{n}', B, Tm).cpu().reshape(k.shape); gb = eng.debug_read(f'kbias{n}', B, Tm).cpu().reshape(bb.shape)
            print(f'B={B} Tm={Tm} blk{n} {mode:10s} kernels max|d| {(gk-k).abs().max().item():.3e} kbias {(gb-bb).abs().max().item():.3e} (rms {k.pow(2).mean().sqrt():.3f})', flush=True)
        for stop, n, hop in ((4, 1, 64), (5, 2, 256)):
            eng.set_option('stop_after', stop); net((x.cuda(), mel.cuda(), t.cuda()))
            got = eng.debug_read(f'lvc{n}', B, Tm).cpu().reshape(B, 32, Tm * hop)
            d = (got - inter[f'lvc{n}']).abs()
            print(f'B={B} Tm={Tm} {mode:10s} lvc{n} max|d| {d.max().item():.3e} mean|d| {d.mean().item():.3e} (rms {inter[f"lvc{n}"].pow(2).mean().sqrt():.3f})', flush=True)
        eng.set_option('stop_after', 99)
        eps = net((x.cuda(), mel.cuda(), t.cuda())).cpu()
        print(f'B={B} Tm={Tm} eps {mode} max|d| vs oracle {(eps-eps_o).abs().max().item():.3e}', flush=True)
