# GPU: tensor-core KC GEMM vs SIMT vs oracle
import sys, torch, time
sys.path.insert(0,'/root/repo')
import fastdiff_b200 as fb
from fastdiff_b200.synthetic import make_state_dict, make_inputs
from oracle import fastdiff_oracle as O
import torch.nn.functional as F
sd = make_state_dict(1234, g_jitter=0.1); W = O.fold_weight_norm(sd)
net = fb.FastDiff().cuda().eval(); net.load_state_dict(sd)
for (B,Tm) in ((1,5),(2,33),(3,300)):
    x, mel = make_inputs(B,Tm,3); t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B,1)
    res = {}
    for mode in ('fp32_simt','tc_3xtf32','tc_tf32'):
        net.mode = mode
        eng = net.engine(); eng.set_option('stop_after', 1)
        net((x.cuda(), mel.cuda(), t.cuda())); torch.cuda.synchronize()
        res[mode] = [ (eng.debug_read(f'kernels{n}',B,Tm).cpu(), eng.debug_read(f'kbias{n}',B,Tm).cpu()) for n in range(3)]
        eng.set_option('stop_after', 99)
    e = O.embed_mlp(W, t)
    for n in range(3):
        p=f'lvc_blocks.{n}'
        noise = F.linear(e, W[f'{p}.fc_t.weight'], W[f'{p}.fc_t.bias']).unsqueeze(-1)
        k, bb = O.kernel_predictor(W, f'{p}.kernel_predictor', mel+noise)
        for mode in res:
            dk = (res[mode][n][0].reshape(k.shape)-k).abs().max().item(); db = (res[mode][n][1].reshape(bb.shape)-bb).abs().max().item()
            print(f'B={B} Tm={Tm} blk{n} {mode:10s} kernels max|d| {dk:.3e} kbias {db:.3e}  (rms {k.pow(2).mean().sqrt():.3f})', flush=True)
    for mode in ('tc_3xtf32','tc_tf32'):
        net.mode = mode
        eps = net((x.cuda(), mel.cuda(), t.cuda())).cpu()
        print(f'B={B} Tm={Tm} eps {mode} max|d| vs oracle {(eps-O.denoise(W,x,mel,t)).abs().max().item():.3e}', flush=True)

print('--- LVC stages (tensor-core layers) ---', flush=True)
for (B,Tm) in ((1,5),(2,33)):
    x, mel = make_inputs(B,Tm,3); t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B,1)
    eps_o, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    for mode in ('fp32_simt','tc_3xtf32','tc_tf32'):
        net.mode = mode; eng = net.engine()
        for stop,n,hop in ((4,1,64),(5,2,256)):
            eng.set_option('stop_after', stop); net((x.cuda(), mel.cuda(), t.cuda()))
            got = eng.debug_read(f'lvc{n}',B,Tm).cpu().reshape(B,32,Tm*hop)
            d = (got-inter[f'lvc{n}']).abs()
            print(f'B={B} Tm={Tm} {mode:10s} lvc{n} max|d| {d.max().item():.3e} mean|d| {d.mean().item():.3e} (rms {inter[f"lvc{n}"].pow(2).mean().sqrt():.3f})', flush=True)
        eng.set_option('stop_after', 99)

print('--- DBlock 0 (tensor cores) ---', flush=True)
for (B,Tm) in ((1,5),(2,33),(3,300)):
    x, mel = make_inputs(B,Tm,3); t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B,1)
    eps_o, inter = O.denoise(W, x, mel, t, return_intermediates=True)
    for mode in ('fp32_simt','tc_3xtf32','tc_tf32'):
        net.mode = mode; eng = net.engine(); eng.set_option('stop_after', 2)
        net((x.cuda(), mel.cuda(), t.cuda()))
        for n,T in enumerate((Tm*64, Tm*8, Tm)):
            d = (eng.debug_read(f'down{n}',B,Tm).cpu().reshape(B,32,T) - inter[f'down{n}']).abs()
            print(f'B={B} Tm={Tm} {mode:10s} down{n} max|d| {d.max().item():.3e} (rms {inter[f"down{n}"].pow(2).mean().sqrt():.3f})', flush=True)
        eng.set_option('stop_after', 99)
