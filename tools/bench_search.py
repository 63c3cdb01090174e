#!/usr/bin/env python3
"""Times the search kernel alone (mcq_search) on the C2 shape, for both kernel versions, and checks that they agree.
    python tools/bench_search.py [frames] [N] [D]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from quantization_b200 import _lib, synth
from helpers import make_quantizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
K = 256
dev = torch.device("cuda:0")
p = synth.synth_params(D, N, K, 0)
q = make_quantizer(D, N, K, p, dev)
x = synth.synth_x(B, D, 1235).to(dev)
L = _lib.lib()
blob = q._prepared()
base = blob.data_ptr()
g_ptr = L.mcq_prepared_gram(base, N, K, D)
P = torch.empty(B, N * K, dtype=torch.float32, device=dev)
ws = q._workspace(B)
_lib.check(L.mcq_xct(x.data_ptr(), 0, B, D, N, K, base, P.data_ptr(), ws.data_ptr(), ws.numel(),
                     _lib.stream_ptr(dev)), "xct")
idx0 = q.encode(x, refine_indexes_iters=0, as_bytes=False).to(torch.int32).contiguous()
# refinement passes these frames actually execute (converged frames stop early): counted by the kernels themselves
_lib.search_stats(ws, reset=True, read=False)
q.encode(x)
passes, nfr = _lib.search_stats(ws)
print(f"frame_passes {passes} frames {nfr} passes_per_frame {passes / max(nfr, 1):.4f}", flush=True)
VERS = tuple(os.environ.get("MCQ_ONLY", "v1,v2").split(","))
res = {}
for ver in VERS:
    os.environ["MCQ_SEARCH"] = ver
    out = torch.empty_like(idx0)
    for it in range(2):
        _lib.check(L.mcq_search(P.data_ptr(), g_ptr, B, N, K, 5, idx0.data_ptr(), out.data_ptr(),
                                _lib.stream_ptr(dev)), "search")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for it in range(reps):
        _lib.check(L.mcq_search(P.data_ptr(), g_ptr, B, N, K, 5, idx0.data_ptr(), out.data_ptr(),
                                _lib.stream_ptr(dev)), "search")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[ver] = out.clone()
    print(f"{ver}: {ms:.3f} ms per launch of {B} frames (N={N}) -> {B / ms / 1e3:.2f} Mvec/s search-only", flush=True)
bad = int((res[VERS[0]] != res[VERS[-1]]).any(1).sum())
print(f"v1 vs v2: {bad}/{B} frames differ")
# one-pass variant (pass count 1) to get time per pass
for ver in VERS:
    os.environ["MCQ_SEARCH"] = ver
    out = torch.empty_like(idx0)
    L.mcq_search(P.data_ptr(), g_ptr, B, N, K, 1, idx0.data_ptr(), out.data_ptr(), _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(5):
        L.mcq_search(P.data_ptr(), g_ptr, B, N, K, 1, idx0.data_ptr(), out.data_ptr(), _lib.stream_ptr(dev))
    e1.record()
    torch.cuda.synchronize()
    print(f"{ver}: 1 pass {e0.elapsed_time(e1) / 5:.3f} ms")
sys.exit(1 if bad else 0)
