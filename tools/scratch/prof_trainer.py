import os, sys, random
sys.path.insert(0, '/root/repo')
import torch
from torch.profiler import profile, ProfilerActivity
from quantization_b200 import QuantizerTrainer, synth
dev = torch.device('cuda:0')
torch.manual_seed(1); random.seed(1)
B, D = 65536, 256
x = synth.synth_x(B, D, 1236, torch.bfloat16).to(dev)
tr = QuantizerTrainer(dim=D, bytes_per_frame=4, device=dev)
tr.cur_iter = tr.phase_one_iters
tr.step(x)
tr.cur_iter = tr.phase_one_iters + 2
for _ in range(3): tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): tr.step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
