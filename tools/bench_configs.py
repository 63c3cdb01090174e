#!/usr/bin/env python3
"""Encode / decode throughput on one GPU for the shapes of every BASELINE.json config (bounded batches).
    python tools/bench_configs.py
Prints one JSON line per config: encode Mvec/s (device-resident frames, 5 refine passes), per-kernel split, decode
Mvec/s and its HBM fraction, round-trip relative MSE."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from quantization_b200 import _lib, synth
from helpers import make_quantizer

dev = torch.device("cuda:0")
CONFIGS = [
    ("C1 d256 n4 b4096 f32", 256, 4, 4096, torch.float32),
    ("C2 d512 n8 b1M f32", 512, 8, 1 << 20, torch.float32),
    ("C3-phase2 d256 n4 b65536 bf16", 256, 4, 65536, torch.bfloat16),
    ("C4 d1024 n16 b131072/GPU-slice f32", 1024, 16, 131072, torch.float32),
    ("C5 d768 n8 b262144 f16", 768, 8, 262144, torch.float16),
]
only = sys.argv[1:] or None
for name, D, N, B, dt in CONFIGS:
    if only and not any(o in name for o in only):
        continue
    p = synth.synth_params(D, N, 256, 0)
    q = make_quantizer(D, N, 256, p, dev)
    x = synth.synth_x(B, D, 1234, dt).to(dev)
    q._prepared()
    for _ in range(2):
        codes = q.encode(x)
    torch.cuda.synchronize()
    _lib.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        codes = q.encode(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    prof = {k: round(v[0] / reps, 3) for k, v in _lib.profile_read().items()}
    _lib.profile(False)
    with torch.no_grad():
        for _ in range(2):
            dec = q.decode(codes)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            dec = q.decode(codes)
        e1.record()
        torch.cuda.synchronize()
        dms = e0.elapsed_time(e1) / 10
        xf = x.float()
        rel = float(((dec - xf) ** 2).sum() / (xf ** 2).sum())
    dbytes = B * (N + D * 4)
    print(json.dumps({"config": name, "encode_Mvec_s": round(B / ms / 1e3, 3), "encode_ms": round(ms, 3),
                      "kernel_ms": prof, "decode_Mvec_s": round(B / dms / 1e3, 1),
                      "decode_GBps": round(dbytes / dms / 1e6, 1), "round_trip_rel_mse": rel}), flush=True)
    del x, codes, dec, q
    torch.cuda.empty_cache()
