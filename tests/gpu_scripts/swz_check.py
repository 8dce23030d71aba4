import sys, torch
sys.path.insert(0,'/root/repo')
import fastdiff_b200 as fb
from fastdiff_b200.synthetic import make_state_dict, make_inputs
from oracle import fastdiff_oracle as O
sd = make_state_dict(1234, g_jitter=0.1); W = O.fold_weight_norm(sd)
net = fb.FastDiff().cuda().eval(); net.load_state_dict(sd); net.mode='tc_3xtf32'
B,Tm = 2,33
x, mel = make_inputs(B,Tm,3); t = torch.tensor([7.413235, 498.0537]).reshape(B,1)
eps_o, inter = O.denoise(W, x, mel, t, return_intermediates=True)
eng = net.engine()
for swz in (0,1,2):
    eng.set_option('lvc_swizzle', swz)
    try:
        eps = net((x.cuda(), mel.cuda(), t.cuda())).cpu(); torch.cuda.synchronize()
        l2 = eng.debug_read('lvc2',B,Tm).cpu().reshape(B,32,Tm*256)
        print(f'swizzle mode {swz}: eps max|d| {(eps-eps_o).abs().max().item():.3e}  lvc2 max|d| {(l2-inter["lvc2"]).abs().max().item():.3e}', flush=True)
    except Exception as e:
        print('swizzle mode', swz, 'FAILED', repr(e)[:300], flush=True); break
# timing of block-2 layers per mode at full size
import time
B,Tm=8,861
x, mel = make_inputs(B,Tm,3); t = torch.full((B,1), 74.99); xc,mc,tc = x.cuda(), mel.cuda(), t.cuda()
for swz in (0,1):
    eng = net.engine(); eng.set_option('lvc_swizzle', swz)
    for _ in range(2): net((xc,mc,tc))
    eng.timing_enable(True); eng.timing_report()
    for _ in range(5): net((xc,mc,tc))
    rep = eng.timing_report(); eng.timing_enable(False)
    print('swizzle', swz, {k: round(v['ms']/5,3) for k,v in rep.items() if 'lvc' in k}, flush=True)
    tl = eng.debug_read('lvc_timeline', B, Tm).cpu().reshape(-1)[:120].reshape(12,10)
    names = ['ld+lwsplit','ausync','Abuilt','sync','convIssued','convDone','Ywritten','lvcIssued','lvcDone']
    print('   tile 6 phases:', dict(zip(names, [int(tl[6,j]-tl[6,j-1]) for j in range(1,10)])), flush=True)
