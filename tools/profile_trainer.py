#!/usr/bin/env python3
"""Where a QuantizerTrainer.step goes, kernel by kernel (torch profiler, CUDA time):
    python tools/profile_trainer.py [batch] [phase]          phase 1: K=16 N=8, phase 2: K=256 N=4 (BASELINE configs[2])"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile

from quantization_b200 import QuantizerTrainer, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
phase = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
torch.manual_seed(1)
random.seed(1)
x = synth.synth_x(B, 256, 1236, torch.bfloat16).to(dev)
tr = QuantizerTrainer(dim=256, bytes_per_frame=4, device=dev, phase_one_iters=10000, phase_two_iters=10000)
tr.cur_iter = 1
if phase == 2:
    tr.cur_iter = tr.phase_one_iters
    tr.step(x)
    tr.cur_iter = tr.phase_one_iters + 2
for _ in range(5):
    tr.step(x)
torch.cuda.synchronize()
steps = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        tr.step(x)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / steps, e.count / steps) for e in prof.key_averages() if e.device_time_total > 0
        and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"phase {phase}: {tot / 1e3:.3f} ms of kernels per step, {sum(r[2] for r in rows):.0f} launches per step")
for k, t, c in rows[:40]:
    print(f"{t:9.1f} us  x{c:4.1f}  {k[:110]}")
