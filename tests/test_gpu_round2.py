"""Round-2 GPU tests (-m gpu): parity against the reference WITH a control (the reference against its own
feature-permuted self), full-size parity against the reference run live on the same GPU, the trainer's end quality
against a reference-trained run, and the regressions the round-1 review asked for (stale gradients around CUDA-graph
replays, the tail chunk of the classifier-loss GEMM, get_product_quantizer / compute_codebook_correlations values)."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from quantization_b200 import Quantizer, QuantizerTrainer, _lib, synth
from helpers import (disagreement, load_npz, make_quantizer, reference_package, trainer_quality_data)

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _record(kind, name, payload):
    import json
    try:
        os.makedirs(os.path.join(_ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(_ROOT, "gpurun_out", "measurements.jsonl"), "a") as f:
            f.write(json.dumps({"kind": kind, "case": name, **payload}) + "\n")
    except OSError:
        pass


def _width(ratios):
    """Spread of fp64 error ratios around 1: the largest |log ratio|."""
    return float(np.abs(np.log(ratios)).max()) if len(ratios) else 0.0


# ---------------------------------------------------------------------------------------------------------------------
# Parity criterion with a control.  SURVEY 8(c) asked: differing-frame rate <= 1e-4 and, on every differing frame, our
# fp64 error <= the reference's x (1 + 1e-5).  The second half cannot hold for ANY fp32 implementation that sums in a
# different order -- the reference fails it against itself: permuting the feature dimension of x / centers / weights
# consistently (the same function, another fp32 summation order) changes 4..7 of 65,536 frames, with error ratios on
# both sides of 1.  That control is the yardstick: our differences from the reference must be no more frequent than
# twice the control's and no wider than the control's.

def test_parity_against_reference_with_control():
    g, m = load_npz("golden_control.npz")
    D, N, K, B = m["D"], m["N"], m["K"], m["B"]
    p = synth.synth_params(D, N, K, m["seed_p"])
    x = synth.synth_x(B, D, m["seed_x"])
    assert synth.sha256_of(x) == m["sha_x"] and synth.sha256_of(p["centers"], p["weight"], p["bias"]) == m["sha_params"]
    ref = g["codes_ref"].astype(np.int64)
    cs = p["centers"].numpy()
    # the control: the reference against its permuted selves
    ctl_counts, ctl_ratios = [], []
    for ps in m["perm_seeds"]:
        rows, ratios = disagreement(g[f"codes_perm{ps}"].astype(np.int64), ref, x.numpy(), cs)
        ctl_counts.append(len(rows))
        ctl_ratios.extend(ratios.tolist())
    ctl_ratios = np.array(ctl_ratios)
    assert min(ctl_counts) >= 1, "the control shows no re-association noise: fixture broken?"
    # the strict 1 + 1e-5 criterion is violated by the reference itself:
    assert (ctl_ratios > 1.0 + 1e-5).any() and (ctl_ratios < 1.0 - 1e-5).any()
    q = make_quantizer(D, N, K, p, DEV)
    ours = q.encode(x.to(DEV)).cpu().numpy().astype(np.int64)
    rows, ratios = disagreement(ours, ref, x.numpy(), cs)
    _record("parity_control", "c2_65536", {"ours_differing": int(len(rows)), "control_differing": ctl_counts,
                                           "ours_ratios": ratios.tolist(), "control_ratios": ctl_ratios.tolist()})
    assert len(rows) / B <= 1e-4, f"{len(rows)}/{B} frames differ"
    # (a) no more frequent than 2x the control (mean over the permutations; +2 frames of Poisson slack at these counts)
    assert len(rows) <= 2 * np.mean(ctl_counts) + 2, (len(rows), ctl_counts)
    # (b) no wider than the control, and not one-sided: ours better on some frames, worse on others (or too few to tell)
    assert _width(ratios) <= max(_width(ctl_ratios) * 1.25, 1e-3), (_width(ratios), _width(ctl_ratios))
    if len(ratios) >= 6:
        assert (ratios < 1).any() and (ratios > 1).any(), ratios


def test_full_size_parity_against_live_reference():
    """BASELINE config 2 in full (2^20 frames): ours against the unmodified reference (baseline/_ref) run on the same
    GPU, with the reference's permuted self as the control, 65,536-frame reference sub-batches."""
    refq = reference_package()
    if refq is None:
        pytest.skip("baseline/_ref is not installed (see DESIGN.md 6); the 65,536-frame fixture test covers this")
    D, N, K, B = 512, 8, 256, 1 << 20
    p = synth.synth_params(D, N, K, 0)
    x = synth.synth_x(B, D, 1234 + 1).to(DEV)
    q = make_quantizer(D, N, K, p, DEV)
    ours = q.encode(x)

    def ref_codes(params, xin):
        r = refq.Quantizer(dim=D, codebook_size=K, num_codebooks=N)
        with torch.no_grad():
            r.centers.copy_(params["centers"])
            r.to_logits.weight.copy_(params["weight"])
            r.to_logits.bias.copy_(params["bias"])
            r = r.to(DEV)
            return torch.cat([r.encode(xin[i:i + 32768], refine_indexes_iters=5) for i in range(0, B, 32768)])
    ref = ref_codes(p, x)
    perm = torch.randperm(D, generator=torch.Generator().manual_seed(11))
    pp = dict(centers=p["centers"][:, :, perm].contiguous(), weight=p["weight"][:, perm].contiguous(), bias=p["bias"])
    ctl = ref_codes(pp, x[:, perm.to(DEV)].contiguous())
    xs, cs = x.cpu().numpy(), p["centers"].numpy()
    r_o, ratios_o = disagreement(ours.cpu().numpy().astype(np.int64), ref.cpu().numpy().astype(np.int64), xs, cs)
    r_c, ratios_c = disagreement(ctl.cpu().numpy().astype(np.int64), ref.cpu().numpy().astype(np.int64), xs, cs)
    _record("parity_full_size", "c2_1M", {"ours_differing": int(len(r_o)), "control_differing": int(len(r_c)),
                                          "ours_ratio_min_med_max": [float(np.min(ratios_o)), float(np.median(ratios_o)),
                                                                     float(np.max(ratios_o))] if len(r_o) else None,
                                          "control_ratio_min_med_max": [float(np.min(ratios_c)), float(np.median(ratios_c)),
                                                                        float(np.max(ratios_c))] if len(r_c) else None,
                                          "ours_better": int((ratios_o < 1).sum()), "ours_worse": int((ratios_o > 1).sum())})
    assert len(r_o) / B <= 1e-4, f"{len(r_o)}/{B} frames differ from the reference"
    assert len(r_o) <= 2 * len(r_c) + 10, (len(r_o), len(r_c))
    assert _width(ratios_o) <= max(_width(ratios_c) * 1.5, 1e-3), (_width(ratios_o), _width(ratios_c))
    assert abs(float(np.median(np.log(ratios_o)))) <= 5e-3  # symmetric around 1: near-tie noise, not a bias
    assert (ratios_o < 1).sum() >= len(ratios_o) * 0.25 and (ratios_o > 1).sum() >= len(ratios_o) * 0.25


# ---------------------------------------------------------------------------------------------------------------------
# Trainer end quality against a reference-trained run (reference's own yardstick, test_quantization.py:11-48, 51-84)

@pytest.mark.parametrize("kind", ["mlp", "gauss"])
def test_trainer_quality_parity(kind):
    g, m = load_npz("golden_trainer_quality.npz")
    dim, B = m["dim"], m["B"]
    torch.manual_seed(1)
    random.seed(1)
    gen_x = trainer_quality_data(kind, dim)
    tr = QuantizerTrainer(dim=dim, bytes_per_frame=m["bytes_per_frame"], device=DEV,
                          phase_one_iters=m["phase_one_iters"], phase_two_iters=m["phase_two_iters"])
    while not tr.done():
        tr.step(gen_x(B).to(DEV))
    q = tr.get_quantizer()
    assert (q.codebook_size, q.num_codebooks) == (256, m["bytes_per_frame"])
    x_mean = q.get_data_mean()
    errs = []
    with torch.no_grad():
        for _ in range(m["eval_batches"]):
            x = gen_x(B).to(DEV)
            xa = q.decode(q.encode(x))
            errs.append(float(((x - xa) ** 2).sum() / ((x - x_mean) ** 2).sum()))
    ours, ref = float(np.mean(errs)), m[f"{kind}_avg_rel_err"]
    _record("trainer_quality", kind, {"ours": ours, "reference": ref, "shannon": m["shannon_distortion"]})
    # SURVEY section 7 hard part 5: final relative error of the two implementations within ~1-2 % of each other
    assert abs(ours - ref) <= 0.02 * ref, (ours, ref)
    if kind == "gauss":  # same distance from the Shannon bound (test_quantization.py:55-60)
        sh = m["shannon_distortion"]
        assert ours >= sh * 0.999, (ours, sh)
        assert abs((ours - sh) - (ref - sh)) <= 0.25 * (ref - sh) + 0.005, (ours, ref, sh)


def test_reference_trained_dim256_quantizer_codes():
    """The 4 x 256 quantizer the reference trained at dim 256 (golden_trainer_quality.npz) loaded into ours: codes of
    2,048 fresh frames equal the reference's, or differ only on adjudicated fp32 near-ties."""
    g, m = load_npz("golden_trainer_quality.npz")
    sd = {k[len("mlp/state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("mlp/state/")}
    q = Quantizer(dim=m["dim"], codebook_size=256, num_codebooks=m["bytes_per_frame"])
    q.load_state_dict(sd)
    q = q.to(DEV)
    x = torch.from_numpy(g["mlp/x_eval"])
    ours = q.encode(x.to(DEV)).cpu().numpy().astype(np.int64)
    ref = g["mlp/codes"].astype(np.int64)
    with torch.no_grad():
        cs = q.get_centers().cpu().numpy()
    rows, ratios = disagreement(ours, ref, x.numpy(), cs)
    assert len(rows) <= 1, f"{len(rows)}/2048 frames differ"
    if len(rows):
        _, margin = oracle.compute_indexes(x.numpy()[rows], sd["centers"].numpy(), sd["to_logits.weight"].numpy(),
                                           sd["to_logits.bias"].numpy(), float(sd["centers_scale"]),
                                           float(sd["logits_scale"]), iters=5, return_margin=True)
        assert np.all(margin <= 1e-6) and _width(ratios) <= 0.1, (margin, ratios)


# ---------------------------------------------------------------------------------------------------------------------
# Regressions from the round-1 review

def test_trainer_graph_and_eager_steps_interleave():
    """after a CUDA-graph capture / replay `p.grad` pointed at the graph's buffers, and the next EAGER step
    (diagnostics iteration, warm-up of the other pass count) accumulated onto the previous replay's gradients.
    (i) invariant, with two_iter_prob = 0.5 across a diagnostics iteration: every update that is launched eagerly
    starts with no gradient on any parameter, and none is left behind after any step; (ii) numbers, with one pass count:
    a graphed and an eager trainer stay together step by step THROUGH the eager diagnostics step that follows replays
    (training on discrete codes amplifies rounding differences, so the comparison runs over 14 steps, like
    test_trainer_cuda_graph_equals_eager; a doubled gradient at the diagnostics step shows as ~2e-3)."""
    dim, B = 64, 2048
    xs = [synth.synth_x(B, dim, 900 + i).to(DEV) for i in range(4)]

    def make(use_graph, prob):
        torch.manual_seed(3)
        random.seed(3)
        tr = QuantizerTrainer(dim=dim, bytes_per_frame=2, device=DEV, phase_one_iters=10000, phase_two_iters=10000)
        tr._use_graph = use_graph
        tr.two_iter_prob = prob
        return tr
    # (i)
    tr = make(True, 0.5)
    tr.cur_iter = 170
    seen = []
    orig = tr._loss_and_update

    def spy(x, n):
        seen.append((tr.cur_iter, torch.cuda.is_current_stream_capturing(),
                     all(p.grad is None for p in tr.quantizer.parameters())))
        return orig(x, n)
    tr._loss_and_update = spy
    for i in range(60):
        tr.step(xs[i % len(xs)])
        assert all(p.grad is None for p in tr.quantizer.parameters())
    assert len(tr._graphs) == 2, "both pass counts should have been captured"
    assert all(clean for _, _, clean in seen), [s for s in seen if not s[2]]
    eager_after_capture = [it for it, cap, _ in seen if not cap and it >= 200]
    assert 200 in eager_after_capture, "the diagnostics step at iteration 200 must have run eagerly after replays"
    # (ii)
    eager, graphed = make(False, 0.0), make(True, 0.0)
    for t in (eager, graphed):
        t.cur_iter = 190
    for i in range(14):  # 190..192 eager, 193 capture + replay, ..., 200 eager (diagnostics), 201.. replays
        for t in (eager, graphed):
            t.step(xs[i % len(xs)])
        for (k, a), (_, b) in zip(eager.quantizer.state_dict().items(), graphed.quantizer.state_dict().items()):
            if a.dtype.is_floating_point:
                assert (a - b).abs().max().item() <= 1e-4 * (a.abs().max().item() + 1e-6), (i, k)
    assert len(graphed._graphs) == 1


def test_prepared_state_follows_fused_optimizer_steps():
    """torch.optim.Adam(fused=True) -- what QuantizerTrainer uses -- updates parameters WITHOUT bumping their version
    counters, which the prepared-state cache (scaled centers, Gram table, operand splits) is keyed on.  Found in round 2:
    eager trainer steps and the no_grad diagnostics ran on tables up to 200 steps old.  After training steps with a
    fused optimiser, encode() must equal the encode() of a fresh module holding the same state_dict."""
    dim, B = 64, 1024
    torch.manual_seed(11)
    q = Quantizer(dim=dim, codebook_size=16, num_codebooks=4).to(DEV)
    opt = torch.optim.Adam(q.parameters(), lr=0.01, fused=True)
    x = synth.synth_x(B, dim, 77).to(DEV)
    versions = [p._version for p in q.parameters()]
    for step in range(6):
        losses = q.compute_loss(x, 1)
        (losses[0] + losses[1] + 0.01 * losses[2]).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        with torch.no_grad():  # like the trainer's diagnostics: right after the optimiser step
            codes = q.encode(x, refine_indexes_iters=2)
            rel = float(q.compute_loss(x, 2)[0])
        fresh = Quantizer(dim=dim, codebook_size=16, num_codebooks=4).to(DEV)
        fresh.load_state_dict(q.state_dict())
        with torch.no_grad():
            assert torch.equal(codes, fresh.encode(x, refine_indexes_iters=2)), f"stale prepared state at step {step}"
            assert rel == float(fresh.compute_loss(x, 2)[0])
    if [p._version for p in q.parameters()] != versions:
        pytest.skip("this PyTorch bumps version counters in fused Adam: the scenario is not reproduced")


_TAIL_CHUNK_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from quantization_b200 import synth
from helpers import make_quantizer
dev = torch.device("cuda:0")
D, N, K = 64, 4, 256
B = 148 * 128 + 1000          # one full 18,944-frame chunk (MCQ_CHUNK_WAVES=1) + a tail that is not a multiple of it
p = synth.synth_params(D, N, K, 5)
x = synth.synth_x(B, D, 31).to(dev)
def losses(gemm):
    os.environ["MCQ_GEMM"] = gemm
    q = make_quantizer(D, N, K, p, dev)
    out = q.compute_loss(x, 1)
    (out[1] + out[2]).backward()
    return [float(v) for v in out], q.to_logits.weight.grad.clone(), q.to_logits.bias.grad.clone()
a, gwa, gba = losses("tc")
b, gwb, gbb = losses("ffma")
print("LOSSES", a, b)
assert all(abs(u - v) <= 2e-6 * max(1.0, abs(v)) for u, v in zip(a, b)), (a, b)
assert torch.allclose(gwa, gwb, rtol=1e-3, atol=1e-7) and torch.allclose(gba, gbb, rtol=1e-3, atol=1e-7)
print("TAIL_OK")
"""


def test_class_loss_tail_chunk_uses_the_right_split_plane():
    """in the last (shorter) chunk of a batch larger than one chunk the classifier GEMM read its second fp16
    plane at the wrong row offset (logits accurate to ~2^-11 only).  The chunk size is fixed per process, hence the
    subprocess with MCQ_CHUNK_WAVES=1; the tcgen05 path must agree with the CUDA-core fp32 GEMM to fp32 accuracy."""
    env = dict(os.environ, MCQ_CHUNK_WAVES="1")
    r = subprocess.run([sys.executable, "-c", _TAIL_CHUNK_SCRIPT.format(root=_ROOT)], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "TAIL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_product_quantizer_and_correlations_match_reference():
    """get_product_quantizer (quantization.py:81-112) and compute_codebook_correlations (:150-181) against
    outputs of the reference itself (golden_misc.npz), on the device."""
    g, meta = load_npz("golden_misc.npz")
    for name, m in meta.items():
        p = synth.synth_params(m["D"], m["N"], m["K"], m["seed"])
        q = make_quantizer(m["D"], m["N"], m["K"], p, DEV, m["centers_scale"], m["logits_scale"])
        corr = q.compute_codebook_correlations().cpu().numpy()
        ref = g[f"{name}/correlations"]
        assert corr.shape == ref.shape and np.allclose(corr, ref, rtol=2e-4, atol=2e-6), np.abs(corr - ref).max()
        if m["K"] == 16:
            pq = q.get_product_quantizer()
            assert (pq.codebook_size, pq.num_codebooks) == (256, m["N"] // 2)
            for what, t in (("weight", pq.to_logits.weight), ("bias", pq.to_logits.bias), ("centers", pq.centers)):
                assert synth.sha256_of(t) == bytes(g[f"{name}/pq_{what}_sha"]).decode(), (name, what)
            assert np.allclose([float(pq.logits_scale), float(pq.centers_scale)], g[f"{name}/pq_scales"])
            x = synth.synth_x(256, m["D"], 77).to(DEV)
            codes = pq.encode(x, refine_indexes_iters=2).cpu().numpy()
            assert int((codes != g[f"{name}/pq_codes"]).any(1).sum()) == 0


def test_search_stats_count_executed_passes():
    """mcq_search_stats: the counters the roofline accounting reads (passes actually executed <= passes requested)."""
    D, N, K, B = 256, 8, 256, 4096
    p = synth.synth_params(D, N, K, 0)
    q = make_quantizer(D, N, K, p, DEV)
    x = synth.synth_x(B, D, 5).to(DEV)
    q.encode(x)
    ws = q._workspace(B)
    _lib.search_stats(ws, reset=True, read=False)
    q.encode(x, refine_indexes_iters=5)
    passes, frames = _lib.search_stats(ws)
    assert frames == B and 2 * B <= passes <= 5 * B, (passes, frames)
    _lib.search_stats(ws, reset=True, read=False)
    q.encode(x, refine_indexes_iters=1)
    assert _lib.search_stats(ws, reset=True) == (B, B)
    assert _lib.search_stats(ws) == (0, 0)


# ---------------------------------------------------------------------------------------------------------------------
# Slab decode (decode.cu: decode_slab_kernel, used for >= 16,384 frames of byte codes): must equal the row-gather
# kernel (which int64 indexes still take) bit for bit, and the CPU oracle on a prefix.
@pytest.mark.parametrize("N,K,D,B", [(8, 256, 512, 70003), (8, 256, 768, 65536), (4, 256, 256, 131075),
                                     (8, 64, 128, 65537), (8, 256, 1024, 66000), (4, 128, 96, 65540),
                                     (4, 64, 4800, 16400)])  # the last: 150 slabs > 148 CTAs, a CTA walks two slabs
def test_decode_slab_matches_row_gather(N, K, D, B):
    p = synth.synth_params(D, N, K, 3)
    q = make_quantizer(D, N, K, p, DEV)
    g = torch.Generator().manual_seed(99)
    codes = torch.randint(0, K, (B, N), generator=g, dtype=torch.int64).to(torch.uint8)
    if K < 256:
        codes[5, 1] = 255  # out-of-range byte: both kernels decode it as entry 0
    cd = codes.to(DEV)
    c64 = cd.to(torch.int64)
    idx64 = torch.where(c64 >= K, torch.zeros_like(c64), c64).contiguous()
    L = _lib.lib()
    blob = q._prepared()
    with torch.no_grad():
        slab = q.decode(cd)
        narrowed = q.decode(idx64)  # large int64 batches are narrowed to bytes and take the slab kernel too
        cs = q.get_centers().cpu().numpy()
    rows = torch.empty(B, D, dtype=torch.float32, device=DEV)  # int64 codes through the C ABI: the row-gather kernel
    _lib.check(L.mcq_decode(idx64.data_ptr(), _lib.I64, B, N, N, K, D, blob.data_ptr(), rows.data_ptr(), _lib.F32,
                            _lib.stream_ptr(DEV)), "mcq_decode")
    torch.cuda.synchronize()
    assert torch.equal(slab, rows) and torch.equal(narrowed, rows)
    bad = c64.clone()
    bad[7, 0] = K + 5
    bad[9, N - 1] = -3
    with torch.no_grad():  # out-of-range indexes decode as entry 0 on either path
        fixed = c64.clone()
        fixed[7, 0] = 0
        fixed[9, N - 1] = 0
        assert torch.equal(q.decode(bad), q.decode(torch.where(fixed >= K, torch.zeros_like(fixed), fixed)))
    ref = oracle.decode(idx64[:1024].cpu().numpy(), cs)
    assert np.array_equal(slab[:1024].cpu().numpy(), ref)
    # half / bfloat16 outputs through the C ABI (Quantizer.decode returns fp32)
    for dt, code in ((torch.float16, _lib.F16), (torch.bfloat16, _lib.BF16)):
        a = torch.empty(B, D, dtype=dt, device=DEV)
        b = torch.empty(B, D, dtype=dt, device=DEV)
        _lib.check(L.mcq_decode(cd.data_ptr(), _lib.U8, B, N, N, K, D, blob.data_ptr(), a.data_ptr(), code,
                                _lib.stream_ptr(DEV)), "mcq_decode")
        _lib.check(L.mcq_decode(idx64.data_ptr(), _lib.I64, B, N, N, K, D, blob.data_ptr(), b.data_ptr(), code,
                                _lib.stream_ptr(DEV)), "mcq_decode")
        torch.cuda.synchronize()
        assert torch.equal(a, b)


def test_decode_slab_under_cuda_graph_capture():
    """The slab decode launch (function attribute + launch, no allocation, no synchronisation) can be captured."""
    N, K, D, B = 8, 256, 512, 20000
    q = make_quantizer(D, N, K, synth.synth_params(D, N, K, 4), DEV)
    codes = torch.randint(0, K, (B, N), dtype=torch.uint8, device=DEV)
    L = _lib.lib()
    blob = q._prepared()
    out = torch.zeros(B, D, dtype=torch.float32, device=DEV)
    with torch.no_grad():
        want = q.decode(codes)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(DEV)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            _lib.check(L.mcq_decode(codes.data_ptr(), _lib.U8, B, N, N, K, D, blob.data_ptr(), out.data_ptr(), _lib.F32,
                                    _lib.stream_ptr(DEV)), "mcq_decode")
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)


# ---------------------------------------------------------------------------------------------------------------------
# The other BASELINE shapes against the CPU oracle at sizes it finishes in seconds (test_large_batch_against_oracle
# covers config 2): config 4 (dim 1024, 16 codebooks: the 16-codebook search kernel), config 5 (dim 768, fp16 frames),
# config 3 phase 2 (dim 256, 4 codebooks, bf16 frames) and phase 1 (16-entry codebooks, packed codes).
@pytest.mark.parametrize("name,D,N,K,B,dtype", [("c4", 1024, 16, 256, 2048, torch.float32),
                                                ("c5", 768, 8, 256, 4096, torch.float16),
                                                ("c3p2", 256, 4, 256, 16384, torch.bfloat16),
                                                ("c3p1", 256, 8, 16, 16384, torch.bfloat16)])
def test_other_configs_against_oracle(name, D, N, K, B, dtype):
    p = synth.synth_params(D, N, K, 0)
    x = synth.synth_x(B, D, 4242, dtype)
    q = make_quantizer(D, N, K, p, DEV)
    idx = q.encode(x.to(DEV), as_bytes=False).cpu().numpy()
    xf = x.float().numpy()  # fp16 / bf16 frames are up-converted exactly: the oracle sees the same values
    ref, margin = oracle.compute_indexes(xf, p["centers"].numpy(), p["weight"].numpy(), p["bias"].numpy(), iters=5,
                                         return_margin=True)
    bad = (idx != ref).any(1)
    rows, ratios = disagreement(idx, ref, xf, p["centers"].numpy())
    _record("other_configs", name, {"frames": B, "differing_frames": int(bad.sum()),
                                    "margins": [float(v) for v in margin[bad]], "err_ratio": ratios.tolist()})
    # the fp32 noise floor grows with the number of selection decisions per frame: <= 1e-4 (+1 frame) up to 8
    # codebooks, <= 5e-4 (+1 frame) at 16
    limit = (5e-4 if N >= 16 else 1e-4) + 1.0 / B
    assert bad.mean() <= limit, f"{int(bad.sum())}/{B} frames differ"
    if bad.any():  # every differing frame an adjudicated fp32 near-tie of the same quality
        assert np.all(margin[bad] <= 1e-6), margin[bad]
        assert np.all(np.abs(np.log(ratios)) <= 0.1), ratios
    # packed byte codes agree with the indexes (config 3 phase 1 packs two 4-bit codes per byte)
    codes = q.encode(x.to(DEV)).cpu()
    with torch.no_grad():
        assert torch.equal(q.decode(codes.to(DEV)), q.decode(torch.from_numpy(idx).to(DEV)))


@pytest.mark.parametrize("B,N,K", [(65536, 8, 16), (70001, 4, 256), (300, 16, 256), (5, 2, 32), (4096, 32, 256)])
def test_index_counts_and_column_sums(B, N, K):
    """mcq_index_counts against a bincount, mcq_column_sums against a float64 column sum (and reproducible)."""
    g = torch.Generator().manual_seed(B + N)
    idx = torch.randint(0, K, (B, N), generator=g, dtype=torch.int64)
    idx[0, 0] = -1  # entries outside [0, K) are not counted
    idx[B - 1, N - 1] = K
    want = torch.zeros(N, K)
    for n in range(N):
        col = idx[:, n]
        col = col[(col >= 0) & (col < K)]
        want[n] = torch.bincount(col, minlength=K).float()
    got = _lib.index_counts(idx.to(DEV), N, K).cpu()
    assert torch.equal(got, want)
    x = torch.randn(B, N * K, generator=g).to(DEV)
    s1 = _lib.column_sums(x)
    s2 = _lib.column_sums(x)
    assert torch.equal(s1, s2)
    ref = x.double().sum(0)
    scale = x.double().abs().sum(0)
    assert float(((s1.double() - ref).abs() / scale).max()) < 1e-6
