import os, sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from quantization_b200 import synth
from helpers import make_quantizer
dev=torch.device('cuda:0')
for (D,N,B) in ((512,8,200000),(256,4,50000),(1024,16,20000),(768,8,30000)):
    p=synth.synth_params(D,N,256,3); q=make_quantizer(D,N,256,p,dev,logits_scale=0.013); x=synth.synth_x(B,D,77).to(dev)
    os.environ['MCQ_ARGMAX']='unfused'; a=q.encode(x, refine_indexes_iters=0, as_bytes=False); a5=q.encode(x)
    del os.environ['MCQ_ARGMAX']; b=q.encode(x, refine_indexes_iters=0, as_bytes=False); b5=q.encode(x)
    print(D,N,B,'init equal:',bool(torch.equal(a,b)),'codes equal:',bool(torch.equal(a5,b5)))
