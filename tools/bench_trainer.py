#!/usr/bin/env python3
"""QuantizerTrainer.step timing at BASELINE configs[2]: dim=256, bytes_per_frame=4, batch=65536 bf16.
    python tools/bench_trainer.py [steps_per_phase] [batch]
Times `steps` steps in phase 1 (K=16, N=8) and in phase 2 (K=256, N=4) with CUDA events, and prints where a step's
GPU time goes (library kernels by kind via mcq_profile -- only meaningful with MCQ_TRAINER_GRAPH=0: launches replayed
from a CUDA graph carry no timing events -- the rest = PyTorch loss arithmetic / Adam)."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from quantization_b200 import QuantizerTrainer, _lib, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
D = 256
dev = torch.device("cuda:0")
torch.manual_seed(1)
random.seed(1)
x = synth.synth_x(B, D, 1234 + 2, torch.bfloat16).to(dev)


def run(tr, tag):
    for _ in range(16):  # long enough for both refinement-pass counts to be captured (3 eager steps each, then the capture)
        tr.step(x)
    torch.cuda.synchronize()
    _lib.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        tr.step(x)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    ms = e0.elapsed_time(e1) / steps
    prof = _lib.profile_read()
    _lib.profile(False)
    parts = {k: round(v[0] / steps, 3) for k, v in prof.items()}
    q = tr.quantizer
    print(f"{tag}: K={q.codebook_size} N={q.num_codebooks} B={B}: {ms:.2f} ms/step (wall {wall:.2f}), "
          f"library kernels ms/step {parts} -> {B / ms / 1e3:.2f} Mframes/s; 20,001 steps = {ms * 20001 / 1e3:.0f} s",
          flush=True)


tr = QuantizerTrainer(dim=D, bytes_per_frame=4, device=dev, phase_one_iters=10000, phase_two_iters=10000)
tr.cur_iter = 1  # stay clear of the every-200-iterations diagnostics
run(tr, "phase 1")
tr.cur_iter = tr.phase_one_iters
tr.step(x)  # switches to phase 2 (get_product_quantizer)
tr.cur_iter = tr.phase_one_iters + 2
run(tr, "phase 2")


# ---- the unmodified reference trainer (baseline/_ref) on the same GPU, same batch (upcast: the reference
# raises on bf16 frames with fp32 parameters, quantization.py:277-279; its own usage upcasts, test_train_hdf5.py:30)
ref_dir = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(ref_dir, "quantization")):
    import types
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, ref_dir)
    import quantization as refq
    xf = x.float()

    def run_ref(tr, tag, n):
        for _ in range(2):
            tr.step(xf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            tr.step(xf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        q = tr.quantizer
        print(f"reference {tag}: K={q.codebook_size} N={q.num_codebooks} B={B}: {ms:.2f} ms/step; "
              f"20,001 steps = {ms * 20001 / 1e3:.0f} s", flush=True)

    torch.manual_seed(1)
    random.seed(1)
    rt = refq.QuantizerTrainer(dim=D, bytes_per_frame=4, device=dev, phase_one_iters=10000, phase_two_iters=10000)
    rt.cur_iter = 1
    run_ref(rt, "phase 1", 5)
    rt.cur_iter = rt.phase_one_iters
    rt.step(xf)
    rt.cur_iter = rt.phase_one_iters + 2
    run_ref(rt, "phase 2", 5)
