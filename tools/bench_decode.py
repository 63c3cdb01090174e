#!/usr/bin/env python3
"""Times Quantizer.decode (mcq_decode) alone: Gvectors/s and bytes written per second against the HBM copy peak.
    [MCQ_DECODE_SLAB=0|1] python tools/bench_decode.py [N] [D] [out_dtype f32|f16]
Run once with MCQ_DECODE_SLAB=0 (row-gather kernel) and once with 1 (slab kernel) for an A/B."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from quantization_b200 import _lib, synth
from helpers import make_quantizer

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
OD = sys.argv[3] if len(sys.argv) > 3 else "f32"
K = 256
dev = torch.device("cuda:0")
q = make_quantizer(D, N, K, synth.synth_params(D, N, K, 0), dev)
L = _lib.lib()
blob = q._prepared()
dt, code, esz = {"f32": (torch.float32, _lib.F32, 4), "f16": (torch.float16, _lib.F16, 2)}[OD]
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
for B in (4096, 65536, 262144, 1 << 20):
    codes = torch.randint(0, K, (B, N), dtype=torch.uint8, device=dev)
    out = torch.empty(B, D, dtype=dt, device=dev)

    def run():
        _lib.check(L.mcq_decode(codes.data_ptr(), _lib.U8, B, N, N, K, D, blob.data_ptr(), out.data_ptr(), code,
                                _lib.stream_ptr(dev)), "mcq_decode")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = B * (N + D * esz) / (ms * 1e-3) / 1e9
    print(f"slab={os.environ.get('MCQ_DECODE_SLAB', 'auto')} N={N} D={D} {OD} B={B}: {ms * 1e3:.1f} us  "
          f"{B / ms / 1e6:.3f} Gvec/s  {gbs:.0f} GB/s = {gbs / peak:.3f} of the HBM copy peak", flush=True)
