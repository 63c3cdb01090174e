import sys, os
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
from quantization_b200 import JointCodebookLoss
B,P,H,N,K=32768,512,512,8,256
dev=torch.device("cuda:0"); torch.manual_seed(0)
x=torch.randn(B,P,device=dev,requires_grad=True); codes=torch.randint(0,K,(B,N),device=dev,dtype=torch.uint8)
mod=JointCodebookLoss(P,N,hidden_channels=H,codebook_size=K,checkpoint=False).to(dev)
def f():
    mod.zero_grad(set_to_none=True); x.grad=None; mod(x,codes).backward()
for _ in range(3): f()
torch.cuda.synchronize()
import time
t=time.time()
for _ in range(10): f()
torch.cuda.synchronize(); print("wall ms/iter", (time.time()-t)*100)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): f()
    torch.cuda.synchronize()
ka=prof.key_averages()
rows=[(k.key, k.count, getattr(k,'device_time_total',0) or getattr(k,'cuda_time_total',0)) for k in ka if (getattr(k,'device_time_total',0) or getattr(k,'cuda_time_total',0))>0 and k.device_type.name=='CUDA']
rows.sort(key=lambda r:-r[2])
tot=sum(r[2] for r in rows)
print("total device us/iter", tot/5, "kernels/iter", sum(r[1] for r in rows)/5)
for r in rows[:25]: print("%9.1f us/iter %5.1f%% n=%5.1f %s"%(r[2]/5,100*r[2]/tot,r[1]/5,r[0][:90]))
