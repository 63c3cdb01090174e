"""Pins the CPU oracle (oracle/mcq_oracle.c) against outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

import oracle
from quantization_b200 import synth
import helpers
from helpers import case_inputs, golden_case_names, trained_params


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_matches_reference_codes(golden_cases, name):
    g, meta = golden_cases
    m = meta[name]
    x, p = case_inputs(m)
    xf = x.float().numpy()
    c, w, b = p["centers"].numpy(), p["weight"].numpy(), p["bias"].numpy()
    idx = oracle.compute_indexes(xf, c, w, b, m["centers_scale"], m["logits_scale"], iters=m["iters"])
    ref = g[name + "/idx"].astype(np.int64)
    assert np.array_equal(idx, ref), f"{int((idx != ref).any(1).sum())} frames differ from the reference"
    assert np.array_equal(oracle.pack(idx, m["K"]), g[name + "/codes"])  # quantization.py:266-272
    # one _refine_indexes call from given indexes (quantization.py:308-547)
    idx0 = synth.synth_indexes(m["B"], m["N"], m["K"], m["seed_i"]).numpy()
    r1 = oracle.compute_indexes(xf, c, w, b, m["centers_scale"], m["logits_scale"], iters=1, idx_in=idx0)
    assert np.array_equal(r1, g[name + "/refine1"].astype(np.int64))


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_decode_matches_reference(golden_cases, name):
    g, meta = golden_cases
    m = meta[name]
    p = synth.synth_params(m["D"], m["N"], m["K"], m["seed_p"])
    dec = oracle.decode(g[name + "/codes"], p["centers"].numpy(), m["centers_scale"])  # packed codes in
    head = g[name + "/decode_head"]
    if m["N"] <= 16:
        # sequential n = 0..N-1 fp32 sum == torch's sum(dim=0) bit for bit
        assert synth.sha256_of(dec) == m["sha_decode"]
        assert np.array_equal(dec[:8], head)
    else:
        # torch switches reduction order for >= 32 addends; north_star tolerance is 1e-5 relative
        assert np.abs(dec[:8] - head).max() <= 1e-5 * np.abs(head).max()
        assert abs(float(dec.astype(np.float64).sum()) - m["decode_sum"]) <= 1e-5 * np.sqrt(m["decode_sumsq"])


def test_oracle_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    for K, N in ((16, 8), (4, 16), (2, 32), (256, 4), (16, 2)):
        idx = rng.integers(0, K, size=(37, N), dtype=np.int64)
        packed = oracle.pack(idx, K)
        kk, cols = K, N
        while kk * kk <= 256:
            kk, cols = kk * kk, cols // 2
        assert packed.shape == (37, cols) and packed.dtype == np.uint8
        assert np.array_equal(oracle.unpack(packed, N, K), idx)


@pytest.mark.parametrize("tag", ["p1", "p2"])
def test_oracle_matches_reference_trained(golden_trained, tag):
    gt = golden_trained
    p = trained_params(gt, tag)
    x = gt["x_eval"]
    idx = oracle.compute_indexes(x, p["centers"].numpy(), p["weight"].numpy(), p["bias"].numpy(),
                                 p["centers_scale"], p["logits_scale"], iters=5)
    ref = gt[f"{tag}/idx"].astype(np.int64)
    nbad = int((idx != ref).any(1).sum())
    # non-trivial scale parameters: the oracle's correctly rounded exp equals torch's fp32 exp on these values
    # (checked below), so the codes must be identical
    for nm in ("centers_scale", "logits_scale"):
        import math
        import torch
        t = float((torch.tensor(p[nm], dtype=torch.float32) * 10.0).exp())
        assert t == float(np.float32(math.exp(float(np.float32(np.float32(p[nm]) * np.float32(10.0))))))
    assert nbad == 0, f"{nbad} of {len(ref)} frames differ"
    K = p["centers"].shape[1]
    dec = oracle.decode(gt[f"{tag}/codes"], p["centers"].numpy(), p["centers_scale"])
    head = gt[f"{tag}/decode_head"]
    assert np.abs(dec[:8] - head).max() <= 1e-5 * np.abs(head).max()
    assert oracle.pack(ref, K).shape == gt[f"{tag}/codes"].shape


def test_oracle_rejects_what_the_reference_rejects():
    x = np.zeros((4, 8), np.float32)
    with pytest.raises(oracle.OracleError):  # K < 16 with N > 1: reference raises UnboundLocalError
        oracle.compute_indexes(x, np.zeros((2, 4, 8), np.float32), np.zeros((8, 8), np.float32),
                               np.zeros((8,), np.float32))
    with pytest.raises(oracle.OracleError):  # not a power of two (quantization.py:33-36)
        oracle.compute_indexes(x, np.zeros((3, 16, 8), np.float32), np.zeros((48, 8), np.float32),
                               np.zeros((48,), np.float32))


def test_oracle_empty_batch():
    p = synth.synth_params(32, 2, 16, 0)
    idx = oracle.compute_indexes(np.zeros((0, 32), np.float32), p["centers"].numpy(), p["weight"].numpy(),
                                 p["bias"].numpy())
    assert idx.shape == (0, 2)
    assert oracle.decode(np.zeros((0, 2), np.int64), p["centers"].numpy()).shape == (0, 32)


# ---- JointCodebookLoss: the numpy restatement against the reference's own outputs ----

@pytest.mark.parametrize("name", helpers.jcl_case_names())
def test_jcl_oracle_matches_reference_golden(name):
    from oracle import jcl_oracle as jo
    g, meta = helpers.jcl_golden()
    m, pred, codes, par, grads = helpers.jcl_case(g, meta, name)
    args = (pred, codes, par["linear1.weight"], par["linear1.bias"], par["codebook_embedding.weight"],
            par["linear2_weight"], par["linear2b_weight"], par["linear2_bias"])
    loss = jo.joint_codebook_loss(*args, ignore_index=-100, reduction=m["reduction"])
    ref = g[name + "/loss"].astype(np.float64)
    assert np.allclose(loss, ref.reshape(np.shape(loss)), rtol=2e-6, atol=1e-6)
    up = g[name + "/upstream"]
    gr = jo.joint_codebook_loss_grads(*args, ignore_index=-100, reduction=m["reduction"], upstream=up)
    pairs = [("pred", g[name + "/g_pred"].reshape(pred.shape)), ("w1", grads["linear1.weight"]),
             ("b1", grads["linear1.bias"]), ("emb", grads["codebook_embedding.weight"]), ("w2", grads["linear2_weight"]),
             ("w2b", grads["linear2b_weight"]), ("bias2", grads["linear2_bias"])]
    for k, ref_g in pairs:
        scale = np.abs(ref_g).max() + 1e-30
        assert np.abs(gr[k] - ref_g).max() <= 2e-5 * scale, k


def test_jcl_hidden_stage_is_the_reference_order():
    """The fp32 stage restatement (what the CUDA kernel is compared with bit for bit) against the fp64 one."""
    from oracle import jcl_oracle as jo
    g, meta = helpers.jcl_golden()
    for name in helpers.jcl_case_names():
        m, pred, codes, par, _ = helpers.jcl_case(g, meta, name)
        hidden = (pred.astype(np.float64) @ par["linear1.weight"].astype(np.float64).T
                  + par["linear1.bias"]).astype(np.float32)
        act = jo.hidden_stage(hidden, codes, par["codebook_embedding.weight"], m["K"])
        # same stage in fp64 started from the same hidden vector
        first = np.maximum(codes[:, :-1], 0) + np.arange(m["N"] - 1) * m["K"]
        e = par["codebook_embedding.weight"].astype(np.float64)[first] * jo.embedding_scale(m["H"], m["N"])
        pre = np.cumsum(np.concatenate([hidden.astype(np.float64)[:, None, :], e], axis=1), axis=1)
        ref = np.maximum(pre, 0).transpose(1, 0, 2)
        assert act.shape == ref.shape and act.dtype == np.float32
        assert np.abs(act - ref).max() <= 1e-5 * (np.abs(ref).max() + 1e-30)
