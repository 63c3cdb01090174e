#!/usr/bin/env python3
"""Generates tests/golden/golden_jcl.npz by running the UNMODIFIED reference `JointCodebookLoss`
(/root/reference/quantization/prediction.py:86-197) on the CPU in this build container: parameters, inputs, the loss
and the autograd gradients of every parameter and of the predictor, for a handful of small shapes (padding via
ignore_index, every reduction, a codebook_size that is not a power of two, more than 33 codebooks, a hidden size
that is not a multiple of 128).  The reference cannot travel to the GPU box; these fixtures do.

    python tests/golden/make_golden_jcl.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference")
import quantization as refq  # noqa: E402  (the reference package)

# name, predictor_channels, hidden_channels, num_codebooks, codebook_size, frames shape, reduction, padded frames, codes dtype
CASES = [
    ("sum_pad", 24, 32, 4, 16, (5, 10), "sum", 7, "int64"),
    ("mean_k256", 40, 64, 8, 256, (33,), "mean", 0, "uint8"),
    ("none_k10", 16, 20, 2, 10, (17,), "none", 3, "int32"),
    ("n40", 8, 8, 40, 4, (9,), "sum", 0, "int64"),
    ("mean_pad", 12, 132, 3, 32, (4, 6), "mean", 5, "int64"),
]


def main():
    out, meta = {}, {}
    for ci, (name, P, H, N, K, shape, reduction, npad, cdt) in enumerate(CASES):
        torch.manual_seed(100 + ci)
        m = refq.JointCodebookLoss(predictor_channels=P, num_codebooks=N, hidden_channels=H, codebook_size=K,
                                   reduction=reduction, checkpoint=False)
        with torch.no_grad():
            m.linear2_bias.copy_(0.3 * torch.randn(N, K))  # the reference initialises it to zero: make it matter
        B = int(np.prod(shape))
        pred = torch.randn(*shape, P, requires_grad=True)
        codes = torch.randint(0, K, (*shape, N))
        if npad:
            flat = codes.reshape(-1, N)
            flat[torch.randperm(B)[:npad]] = -100  # whole frames, as the reference requires (:171-175)
        loss = m(pred, codes)
        up = torch.randn(loss.shape) if reduction == "none" else torch.tensor(0.7)
        (loss * up).sum().backward()
        out[name + "/pred"] = pred.detach().numpy()
        out[name + "/codes"] = codes.numpy().astype(np.int64)
        out[name + "/loss"] = loss.detach().numpy()
        out[name + "/upstream"] = up.numpy()
        out[name + "/g_pred"] = pred.grad.numpy()
        for pn, p in m.named_parameters():
            out[f"{name}/param/{pn}"] = p.detach().numpy()
            out[f"{name}/grad/{pn}"] = p.grad.numpy()
        meta[name] = dict(P=P, H=H, N=N, K=K, shape=list(shape), reduction=reduction, npad=npad, codes_dtype=cdt)
        print(name, "loss", loss.detach().reshape(-1)[:3].tolist(), flush=True)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_jcl.npz"), **out)


if __name__ == "__main__":
    main()
