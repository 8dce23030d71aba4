import sys, torch
sys.path.insert(0,'/root/repo')
import fastdiff_b200 as fb
from fastdiff_b200.synthetic import make_state_dict, make_inputs
from oracle import fastdiff_oracle as O
import torch.nn.functional as F
sd = make_state_dict(1234, g_jitter=0.1); W = O.fold_weight_norm(sd)
net = fb.FastDiff().cuda().eval(); net.load_state_dict(sd); net.mode='tc_3xtf32'
eng = net.engine()
for (B,Tm) in ((1,5),(2,33),(3,300)):
    x, mel = make_inputs(B,Tm,3); t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B,1)
    e = O.embed_mlp(W, t)
    for two in (0,1):
        eng.set_option('kc_2cta', two); eng.set_option('stop_after', 1)
        net((x.cuda(), mel.cuda(), t.cuda())); torch.cuda.synchronize()
        for n in range(3):
            p=f'lvc_blocks.{n}'
            noise = F.linear(e, W[f'{p}.fc_t.weight'], W[f'{p}.fc_t.bias']).unsqueeze(-1)
            k, bb = O.kernel_predictor(W, f'{p}.kernel_predictor', mel+noise)
            dk = (eng.debug_read(f'kernels{n}',B,Tm).cpu().reshape(k.shape)-k).abs().max().item()
            db = (eng.debug_read(f'kbias{n}',B,Tm).cpu().reshape(bb.shape)-bb).abs().max().item()
            print(f'B={B} Tm={Tm} kc_2cta={two} blk{n} kernels max|d| {dk:.3e} kbias {db:.3e}', flush=True)
        eng.set_option('stop_after', 99)
B,Tm=8,861
x, mel = make_inputs(B,Tm,3); t = torch.full((B,1), 74.99); xc,mc,tc = x.cuda(), mel.cuda(), t.cuda()
for two in (0,1):
    eng.set_option('kc_2cta', two)
    for _ in range(2): net((xc,mc,tc))
    eng.timing_enable(True); eng.timing_report()
    for _ in range(5): net((xc,mc,tc))
    rep = eng.timing_report(); eng.timing_enable(False)
    print('kc_2cta', two, 'kc_gemm ms/launch', round(rep['kc_gemm']['ms']/5,4), flush=True)
