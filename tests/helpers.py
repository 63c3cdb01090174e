"""Shared helpers for the tests: re-create golden inputs from seeds, call the oracle."""
import json
import os

import numpy as np
import torch

from quantization_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_case_names():
    g = np.load(os.path.join(GOLDEN_DIR, "golden_cases.npz"))
    return list(json.loads(bytes(g["meta_json"]).decode()).keys())


def case_inputs(m):
    """(x in its stored dtype, params dict) for a golden case; verifies the SHA-256 stored with the fixture."""
    p = synth.synth_params(m["D"], m["N"], m["K"], m["seed_p"])
    x = synth.synth_x(m["B"], m["D"], m["seed_x"], getattr(torch, m["x_dtype"]))
    assert synth.sha256_of(x) == m["sha_x"], "synthetic frames drifted from the fixture"
    assert synth.sha256_of(p["centers"], p["weight"], p["bias"]) == m["sha_params"], "synthetic params drifted"
    return x, p


def trained_params(gt, tag):
    """Parameters of the reference-trained quantizer stored in golden_trained.npz (tag 'p1' or 'p2')."""
    return dict(
        centers=torch.from_numpy(gt[f"{tag}/centers"]),
        weight=torch.from_numpy(gt[f"{tag}/to_logits.weight"]),
        bias=torch.from_numpy(gt[f"{tag}/to_logits.bias"]),
        centers_scale=float(gt[f"{tag}/centers_scale"]),
        logits_scale=float(gt[f"{tag}/logits_scale"]),
        id_buf=torch.from_numpy(gt[f"{tag}/id_buf"]),
    )


def make_quantizer(D, N, K, params, device, centers_scale=0.0, logits_scale=0.0):
    """A quantization_b200.Quantizer on `device` holding the given parameter tensors."""
    from quantization_b200 import Quantizer
    q = Quantizer(dim=D, codebook_size=K, num_codebooks=N)
    with torch.no_grad():
        q.centers.copy_(params["centers"])
        q.to_logits.weight.copy_(params["weight"])
        q.to_logits.bias.copy_(params["bias"])
        q.centers_scale.fill_(centers_scale)
        q.logits_scale.fill_(logits_scale)
    return q.to(device)


def search_supported(N, K):
    """(K, N) pairs the CUDA search kernels are built for: everything the reference itself supports up to
    codebook_size 256 and 64 codebooks (K < 16 with N > 1 crashes in the reference, quantization.py:453,470,504-507)."""
    return K <= 256 and N <= 64 and (N == 1 or K >= 16)


_JCL_PARAM_ORDER = ("linear1.weight", "linear1.bias", "codebook_embedding.weight", "linear2_weight", "linear2b_weight",
                    "linear2_bias")


def jcl_golden():
    """(arrays, meta) of tests/golden/golden_jcl.npz (made by tests/golden/make_golden_jcl.py from the reference)."""
    import json
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_jcl.npz"))
    meta = json.loads(bytes(g["meta_json"]).decode())
    return g, meta


def jcl_case_names():
    return list(jcl_golden()[1].keys())


def jcl_case(g, meta, name):
    """Inputs of one golden JointCodebookLoss case as numpy arrays, flattened to (B, .)."""
    m = meta[name]
    par = {pn: g[f"{name}/param/{pn}"] for pn in _JCL_PARAM_ORDER}
    grads = {pn: g[f"{name}/grad/{pn}"] for pn in _JCL_PARAM_ORDER}
    pred = g[name + "/pred"].reshape(-1, m["P"])
    codes = g[name + "/codes"].reshape(-1, m["N"])
    return m, pred, codes, par, grads
