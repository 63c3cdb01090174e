import os, sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from quantization_b200 import synth, _lib
from helpers import make_quantizer
dev=torch.device('cuda:0'); D,N,B=512,8,1<<20
p=synth.synth_params(D,N,256,0); q=make_quantizer(D,N,256,p,dev); x=synth.synth_x(B,D,1235).to(dev)
q._prepared()
for _ in range(2): c=q.encode(x)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): c=q.encode(x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/3
print(os.environ.get('MCQ_CHUNK_WAVES','4'), f'{ms:.2f} ms -> {B/ms/1e3:.3f} Mvec/s', int(c.sum()))
