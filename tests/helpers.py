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


# ---- round 2 ---------------------------------------------------------------------------------------------------------

def load_npz(name):
    """(arrays, meta) of a tests/golden/*.npz fixture that carries a `meta_json` entry."""
    g = np.load(os.path.join(GOLDEN_DIR, name))
    return g, json.loads(bytes(g["meta_json"]).decode())


def reference_package():
    """The unmodified reference package, pip-installed into the git-ignored baseline/_ref (it travels to the GPU box
    with the snapshot; /root/reference itself does not exist there), or None.  Only used to run the reference LIVE on
    the GPU beside the product in the full-size parity tests."""
    import sys
    import types
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "quantization")):
        return None
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # imported by the reference for read_hdf5_data only
    sys.path.insert(0, ref_dir)
    try:
        import quantization as refq
        return refq
    except Exception:
        return None
    finally:
        sys.path.remove(ref_dir)


def fp64_errors(idx, x, scaled_centers, rows):
    """Per-frame squared reconstruction error, evaluated in float64, of the int codes idx[rows] (B', N)."""
    c64 = np.asarray(scaled_centers, dtype=np.float64)
    x64 = np.asarray(x, dtype=np.float64)[rows]
    rec = sum(c64[n, np.asarray(idx)[rows, n]] for n in range(c64.shape[0]))
    return ((rec - x64) ** 2).sum(1)


def disagreement(idx_a, idx_b, x, scaled_centers):
    """Frames on which two code sets differ and, on those, err64(a) / err64(b).  Returns (rows, ratios)."""
    rows = np.nonzero((np.asarray(idx_a) != np.asarray(idx_b)).any(1))[0]
    if rows.size == 0:
        return rows, np.zeros(0)
    ea = fp64_errors(idx_a, x, scaled_centers, rows)
    eb = fp64_errors(idx_b, x, scaled_centers, rows)
    return rows, ea / np.maximum(eb, 1e-300)


def trainer_quality_data(kind, dim):
    """The batch generator of tests/golden/make_golden_r2.py::trainer_data (same seeds, same order of RNG use):
    'mlp' = the data of the reference's own trainer test (test_quantization.py:15-22, 31-33), 'gauss' = :66-67."""
    from torch import nn
    gen = torch.Generator().manual_seed(4242)
    if kind == "mlp":
        model = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(), nn.Linear(dim, dim), nn.ReLU(), nn.LayerNorm(dim),
                              nn.Linear(dim, dim))

        def f(b):
            with torch.no_grad():
                x = torch.randn(b, dim, generator=gen)
                return model(x) + 0.05 * x
        return f
    return lambda b: torch.randn(b, dim, generator=gen)
