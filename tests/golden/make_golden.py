#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (danpovey/quantization,
/root/reference) on the CPU in this build container.  The reference cannot travel to the GPU box,
so its outputs are committed as small fixtures; the inputs are re-created from seeds by
quantization_b200.synth (their SHA-256 is stored so drift is detected, not silently accepted).

    python tests/golden/make_golden.py            # writes golden_cases.npz, golden_trained.npz

The reference has no golden vectors of its own, which is why they are made here.
The only shim needed is an empty `h5py` module (reference quantization.py:2 imports it; it is only
used by read_hdf5_data, which is off the hot path).
"""
import json
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference")
import quantization as refq  # noqa: E402  (the reference package)

from quantization_b200 import synth  # noqa: E402

torch.set_num_threads(8)

# name, D, N, K, B, iters, centers_scale, logits_scale, x dtype
CASES = [
    ("c1_d256_n4_b4096", 256, 4, 256, 4096, 5, 0.0, 0.0, "float32"),  # BASELINE config 1 in full
    ("c2s_d512_n8_b1024", 512, 8, 256, 1024, 5, 0.0, 0.0, "float32"),  # config 2 shape
    ("c4s_d1024_n16_b256", 1024, 16, 256, 256, 5, 0.0, 0.0, "float32"),  # config 4 shape
    ("c5s_d768_n8_b512_f16", 768, 8, 256, 512, 5, 0.0, 0.0, "float16"),  # config 5 shape (x upcast)
    ("c3p1_d256_k16_n8_b1024_bf16", 256, 8, 16, 1024, 2, 0.0, 0.0, "bfloat16"),  # config 3 phase 1
    ("c3p2_d256_n4_b1024_bf16", 256, 4, 256, 1024, 1, 0.0, 0.0, "bfloat16"),  # config 3 phase 2
    ("n1_d128_b512", 128, 1, 256, 512, 3, 0.0, 0.0, "float32"),
    ("n2_d128_b512", 128, 2, 256, 512, 5, 0.0, 0.0, "float32"),
    ("n32_d256_b64", 256, 32, 256, 64, 2, 0.0, 0.0, "float32"),
    ("k16_n16_d128_b512", 128, 16, 16, 512, 5, 0.0, 0.0, "float32"),
    ("k16_n64_d192_b64", 192, 64, 16, 64, 2, 0.0, 0.0, "float32"),  # trainer phase 1 at bytes_per_frame=32
    ("k64_n4_d96_b512", 96, 4, 64, 512, 5, 0.0, 0.0, "float32"),
    ("k32_n2_d40_b256", 40, 2, 32, 256, 4, 0.0, 0.0, "float32"),  # dim not a multiple of 16
    ("scaled_d256_n8_b512", 256, 8, 256, 512, 5, 0.03, -0.02, "float32"),  # non-trivial scale parameters
    ("iters0_d256_n8_b512", 256, 8, 256, 512, 0, 0.0, 0.0, "float32"),  # classifier arg-max only
]


def make_ref_quantizer(D, N, K, seed, centers_scale, logits_scale):
    q = refq.Quantizer(dim=D, codebook_size=K, num_codebooks=N)
    p = synth.synth_params(D, N, K, seed)
    with torch.no_grad():
        q.centers.copy_(p["centers"])
        q.to_logits.weight.copy_(p["weight"])
        q.to_logits.bias.copy_(p["bias"])
        q.centers_scale.fill_(centers_scale)
        q.logits_scale.fill_(logits_scale)
    return q, p


def main():
    out = {}
    meta = {}
    for ci, (name, D, N, K, B, iters, cs, ls, xdt) in enumerate(CASES):
        seed_x, seed_p, seed_i = 1234 + ci, 100 + ci, 500 + ci
        q, p = make_ref_quantizer(D, N, K, seed_p, cs, ls)
        x_in = synth.synth_x(B, D, seed_x, getattr(torch, xdt))
        x = x_in.float()  # the reference raises on non-fp32 x; its own usage upcasts
        with torch.no_grad():
            codes = q.encode(x, refine_indexes_iters=iters, as_bytes=True)
            idx = q.encode(x, refine_indexes_iters=iters, as_bytes=False)
            dec = q.decode(codes)
            # one _refine_indexes call from random starting indexes
            idx0 = synth.synth_indexes(B, N, K, seed_i)
            idx1 = q._refine_indexes(x, idx0)
            nb = min(B, 256)
            losses = [float(v) for v in q.compute_loss(x[:nb], min(iters, 2))]
        meta[name] = dict(
            D=D, N=N, K=K, B=B, iters=iters, centers_scale=cs, logits_scale=ls, x_dtype=xdt,
            seed_x=seed_x, seed_p=seed_p, seed_i=seed_i,
            sha_x=synth.sha256_of(x_in), sha_params=synth.sha256_of(p["centers"], p["weight"], p["bias"]),
            sha_decode=synth.sha256_of(dec), decode_sum=float(dec.double().sum()),
            decode_sumsq=float((dec.double() ** 2).sum()),
            rel_err=float(((dec - x) ** 2).sum() / (x ** 2).sum()),
            loss_frames=nb, loss_iters=min(iters, 2), losses=losses,
        )
        out[name + "/codes"] = codes.numpy()
        out[name + "/idx"] = idx.numpy().astype(np.int16)
        out[name + "/decode_head"] = dec[:8].numpy()
        out[name + "/refine1"] = idx1.numpy().astype(np.int16)
        print(name, "codes", tuple(codes.shape), "rel_err %.4f" % meta[name]["rel_err"], flush=True)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_cases.npz"), **out)

    # ---- a reference-TRAINED quantizer (both phases), so parity is also checked on non-synthetic state
    torch.manual_seed(1)
    random.seed(1)
    dim, bpf = 64, 4
    trainer = refq.QuantizerTrainer(dim=dim, bytes_per_frame=bpf, device=torch.device("cpu"),
                                    phase_one_iters=300, phase_two_iters=300)
    mix = torch.randn(dim, dim) / dim ** 0.5
    phase1_state = None
    gen = torch.Generator().manual_seed(7)

    def gen_x(b):
        z = torch.randn(b, dim, generator=gen)
        return torch.tanh(z @ mix) + 0.1 * z

    while not trainer.done():
        if trainer.cur_iter == trainer.phase_one_iters and phase1_state is None:
            phase1_state = {k: v.clone() for k, v in trainer.quantizer.state_dict().items()}
        trainer.step(gen_x(512))
    q2 = trainer.get_quantizer()
    tr = {}
    x_eval = gen_x(2048)
    tr["x_eval"] = x_eval.numpy()
    for tag, state, (N, K) in (("p1", phase1_state, (2 * bpf, 16)), ("p2", q2.state_dict(), (bpf, 256))):
        q = refq.Quantizer(dim=dim, codebook_size=K, num_codebooks=N)
        q.load_state_dict(state)
        for k, v in state.items():
            tr[f"{tag}/{k}"] = v.detach().numpy()
        with torch.no_grad():
            tr[f"{tag}/codes"] = q.encode(x_eval).numpy()
            tr[f"{tag}/idx"] = q.encode(x_eval, as_bytes=False).numpy().astype(np.int16)
            dec = q.decode(q.encode(x_eval))
            tr[f"{tag}/decode_head"] = dec[:8].numpy()
            tr[f"{tag}/sha_decode"] = np.frombuffer(synth.sha256_of(dec).encode(), dtype=np.uint8)
            tr[f"{tag}/losses"] = np.array([float(v) for v in q.compute_loss(x_eval[:256], 2)])
            rel = float(((dec - x_eval) ** 2).sum() / ((x_eval - q.get_data_mean()) ** 2).sum())
        print("trained", tag, "N,K", N, K, "rel err %.4f" % rel, "scales", float(q.centers_scale),
              float(q.logits_scale), flush=True)
    np.savez_compressed(os.path.join(HERE, "golden_trained.npz"), **tr)


if __name__ == "__main__":
    main()
