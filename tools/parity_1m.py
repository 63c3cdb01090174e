#!/usr/bin/env python3
"""One-off parity run at full size: our encode vs the unmodified reference (baseline/_ref) on the same GPU, 2^20 frames
of BASELINE configs[1].  Prints the differing-frame rate and, for the differing frames, the fp64 reconstruction errors."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from quantization_b200 import synth
from helpers import make_quantizer

dev = torch.device("cuda:0")
D, N, K, B = 512, 8, 256, 1 << 20
p = synth.synth_params(D, N, K, 0)
q = make_quantizer(D, N, K, p, dev)
x = synth.synth_x(B, D, 1235).to(dev)
ours = q.encode(x)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import quantization as refq

r = refq.Quantizer(dim=D, codebook_size=K, num_codebooks=N)
with torch.no_grad():
    r.centers.copy_(p["centers"])
    r.to_logits.weight.copy_(p["weight"])
    r.to_logits.bias.copy_(p["bias"])
r = r.to(dev)
with torch.no_grad():
    ref = torch.cat([r.encode(x[i:i + 16384]) for i in range(0, B, 16384)])
bad = (ours != ref).any(1)
nb = int(bad.sum())
print(f"frames: {B}, differing: {nb} ({nb / B:.2e})")
if nb:
    c64 = p["centers"].double().to(dev)
    xb = x[bad].double()

    def err(codes):
        rec = sum(c64[n][codes[:, n].long()] for n in range(N))
        return ((rec - xb) ** 2).sum(1)
    eo, er = err(ours[bad]), err(ref[bad])
    ratio = (eo / er)
    print(f"fp64 reconstruction error ours/reference on the differing frames: median {ratio.median():.6f}, "
          f"min {ratio.min():.6f}, max {ratio.max():.6f}; ours better on {int((eo < er).sum())}, worse on "
          f"{int((eo > er).sum())}")
