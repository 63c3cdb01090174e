#!/usr/bin/env python3
"""JointCodebookLoss forward + backward timing beside the unmodified reference module
(baseline/_ref) on the same GPU and the same inputs.
    python tools/bench_jcl.py [frames] [predictor_channels] [hidden_channels] [num_codebooks]
Prints ms per forward+backward for both, per-stage CUDA-event times of the three library kernels, and the HBM bytes
they move against the measured copy peak."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from quantization_b200 import JointCodebookLoss, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
N = int(sys.argv[4]) if len(sys.argv) > 4 else 8
K = 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(B, P, device=dev, requires_grad=True)
codes = torch.randint(0, K, (B, N), device=dev, dtype=torch.uint8)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd_bwd(mod, c):
    def f():
        mod.zero_grad(set_to_none=True)
        x.grad = None
        mod(x, c).backward()
    return f


res = {"frames": B, "predictor_channels": P, "hidden_channels": H, "num_codebooks": N, "codebook_size": K}
for ck in (False, True):
    mod = JointCodebookLoss(P, N, hidden_channels=H, codebook_size=K, checkpoint=ck).to(dev)
    res[f"ours_ms_checkpoint_{ck}"] = round(timed(fwd_bwd(mod, codes)), 3)
    torch.cuda.reset_peak_memory_stats()
    fwd_bwd(mod, codes)()
    res[f"ours_peak_gb_checkpoint_{ck}"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)

# the three library kernels alone
L = _lib.lib()
peak = 6650.0  # fallback of /opt/skills/guides/B200_PROFILING.md when MEASURED_PEAKS.json is absent
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res["hbm_peak_GBps"] = peak
hidden = torch.randn(B, H, device=dev)
emb = torch.randn((N - 1) * K, H, device=dev)
act = torch.empty(N, B, H, device=dev)
st = _lib.stream_ptr(dev)
t = timed(lambda: L.mcq_jcl_hidden_forward(hidden.data_ptr(), codes.data_ptr(), 0, B, N, K, H, emb.data_ptr(), 1.0,
                                           act.data_ptr(), st))
by = (B * H * 4 + N * B * H * 4 + B * N)
res["hidden_forward"] = {"ms": round(t, 4), "GBps": round(by / t / 1e6, 1), "frac_of_copy_peak": round(by / t / 1e6 / peak, 3)}
gact = torch.randn(N, B, H, device=dev)
gh = torch.empty(B, H, device=dev)
gemb = torch.zeros_like(emb)
t = timed(lambda: L.mcq_jcl_hidden_backward(gact.data_ptr(), act.data_ptr(), codes.data_ptr(), 0, B, N, K, H, 1.0,
                                            gh.data_ptr(), gemb.data_ptr(), st))
by = (2 * N * B * H * 4 + B * H * 4 + B * N)
res["hidden_backward"] = {"ms": round(t, 4), "GBps": round(by / t / 1e6, 1), "frac_of_copy_peak": round(by / t / 1e6 / peak, 3)}
logits = torch.randn(B, N * K, device=dev)
bias = torch.zeros(N, K, device=dev)
row = torch.empty(B, N, device=dev)
sums = torch.empty(2, device=dev)
part = torch.empty(L.mcq_jcl_partials(), device=dev)
t = timed(lambda: L.mcq_jcl_cross_entropy(logits.data_ptr(), bias.data_ptr(), codes.data_ptr(), 0, B, N, K, -100, 1,
                                          row.data_ptr(), sums.data_ptr(), part.data_ptr(), st))
by = (2 * B * N * K * 4 + B * N * 5)
res["cross_entropy"] = {"ms": round(t, 4), "GBps": round(by / t / 1e6, 1), "frac_of_copy_peak": round(by / t / 1e6 / peak, 3)}

ref_dir = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(os.path.join(ref_dir, "quantization")):
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, ref_dir)
    import quantization as refq
    c64 = codes.to(torch.int64)
    for ck in (False, True):
        rm = refq.JointCodebookLoss(P, N, hidden_channels=H, codebook_size=K, checkpoint=ck).to(dev)
        res[f"reference_ms_checkpoint_{ck}"] = round(timed(fwd_bwd(rm, c64), n=5, warm=2), 3)
        torch.cuda.reset_peak_memory_stats()
        fwd_bwd(rm, c64)()
        res[f"reference_peak_gb_checkpoint_{ck}"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
print(json.dumps(res))
