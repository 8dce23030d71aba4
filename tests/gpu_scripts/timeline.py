import sys, torch
sys.path.insert(0,'/root/repo')
import fastdiff_b200 as fb
from fastdiff_b200.synthetic import make_state_dict, make_inputs
net = fb.FastDiff().cuda().eval(); net.load_state_dict(make_state_dict(1234))
B,Tm=8,861
x, mel = make_inputs(B,Tm,3); t = torch.full((B,1), 74.99)
for _ in range(2): net((x.cuda(), mel.cuda(), t.cuda()))
torch.cuda.synchronize()
tl = net.engine().debug_read('lvc_timeline', B, Tm).cpu().reshape(-1)[:120].reshape(12,10)
names = ['start','loadwait','ausync','Abuilt','sync','convIssued+halo','convDone','Ywritten','lvcIssued+xs','lvcDone']
print('per-tile phase durations (cycles), group 0 of CTA 0, last LVC layer (dil 27) of block 2:')
for i in range(12):
    d = [int(tl[i,j]-tl[i,j-1]) for j in range(1,10)]
    nxt = int(tl[i+1,0]-tl[i,9]) if i<11 else -1
    print(i, dict(zip(names[1:], d)), 'epilogue+sync->next', nxt, 'total', int(tl[i,9]-tl[i,0])+max(nxt,0))
