#!/usr/bin/env python3
"""Round-2 fixtures, again produced by running the UNMODIFIED reference (/root/reference) on the CPU of the build
container (it cannot travel to the GPU box):

    python tests/golden/make_golden_r2.py control     # golden_control.npz   (~6 min on 8 cores)
    python tests/golden/make_golden_r2.py trainer     # golden_trainer_quality.npz  (~15 min)
    python tests/golden/make_golden_r2.py misc        # golden_misc.npz (seconds)

control -- the reference against ITSELF with the feature dimension of x / centers / to_logits.weight permuted
    consistently (SURVEY.md appendix B "fp32 re-association floor"): mathematically the same function, a different
    fp32 summation order.  65,536 frames of BASELINE config 2 (dim 512, 8 x 256, 5 passes).  The frames on which the
    two disagree, and the fp64 reconstruction-error ratios on them, are the yardstick the GPU path's own differences
    from the reference are held to (tests/test_gpu_parity.py::test_parity_against_reference_with_control).
trainer -- the reference QuantizerTrainer at the setting of its own test (test_quantization.py:11-48: dim 256,
    bytes_per_frame 4, batches of 600, torch.manual_seed(1)), 2,000 + 2,000 iterations, on (i) the MLP-shaped data of
    that test and (ii) Gaussian data (test_quantization.py:51-84, compared with the Shannon bound): final relative
    reconstruction error over 30 fresh batches, the logged per-200-iteration losses, and the final state_dict of (i).
misc -- compute_codebook_correlations / get_product_quantizer outputs of the reference on seeded quantizers.
"""
import json
import os
import random
import sys
import time
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, "/root/reference")
import quantization as refq  # noqa: E402  (the reference package)

from quantization_b200 import synth  # noqa: E402

CONTROL = dict(D=512, N=8, K=256, B=65536, iters=5, seed_x=1234 + 1, seed_p=0, perm_seeds=[11, 12, 13])
TRAINER = dict(dim=256, bytes_per_frame=4, B=600, phase_one_iters=2000, phase_two_iters=2000, eval_batches=30)


def ref_quantizer(D, N, K, p):
    q = refq.Quantizer(dim=D, codebook_size=K, num_codebooks=N)
    with torch.no_grad():
        q.centers.copy_(p["centers"])
        q.to_logits.weight.copy_(p["weight"])
        q.to_logits.bias.copy_(p["bias"])
    return q


def encode_chunks(q, x, iters, chunk=4096):
    out = []
    with torch.no_grad():
        for i in range(0, x.shape[0], chunk):
            out.append(q.encode(x[i:i + chunk], refine_indexes_iters=iters, as_bytes=True))
    return torch.cat(out)


def make_control():
    c = CONTROL
    D, N, K, B = c["D"], c["N"], c["K"], c["B"]
    p = synth.synth_params(D, N, K, c["seed_p"])
    x = synth.synth_x(B, D, c["seed_x"])
    t0 = time.time()
    codes = encode_chunks(ref_quantizer(D, N, K, p), x, c["iters"])
    print("reference: %.0f s" % (time.time() - t0), flush=True)
    out = {"codes_ref": codes.numpy()}
    for ps in c["perm_seeds"]:
        perm = torch.randperm(D, generator=torch.Generator().manual_seed(ps))
        pp = dict(centers=p["centers"][:, :, perm].contiguous(), weight=p["weight"][:, perm].contiguous(),
                  bias=p["bias"])
        cp = encode_chunks(ref_quantizer(D, N, K, pp), x[:, perm].contiguous(), c["iters"])
        nd = int((cp != codes).any(1).sum())
        print(f"perm seed {ps}: {nd} / {B} frames differ from the unpermuted reference", flush=True)
        out[f"codes_perm{ps}"] = cp.numpy()
    meta = dict(c, sha_x=synth.sha256_of(x), sha_params=synth.sha256_of(p["centers"], p["weight"], p["bias"]))
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_control.npz"), **out)


def trainer_data(kind, dim):
    """Batch generator shared with the GPU test (tests/helpers.py::trainer_quality_data mirrors it): a dedicated CPU
    generator, so both implementations train on the same frames whatever else consumes torch's global RNG."""
    gen = torch.Generator().manual_seed(4242)
    if kind == "mlp":  # test_quantization.py:15-22, 31-33
        model = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(), nn.Linear(dim, dim), nn.ReLU(), nn.LayerNorm(dim),
                              nn.Linear(dim, dim))

        def f(b):
            with torch.no_grad():
                x = torch.randn(b, dim, generator=gen)
                return model(x) + 0.05 * x
        return f
    return lambda b: torch.randn(b, dim, generator=gen)  # test_quantization.py:66-67


def make_trainer():
    import logging
    t = TRAINER
    out, meta = {}, dict(t)
    for kind in ("mlp", "gauss"):
        torch.manual_seed(1)
        random.seed(1)
        gen_x = trainer_data(kind, t["dim"])  # (the MLP is initialised first, like test_quantization.py:12-23)
        trainer = refq.QuantizerTrainer(dim=t["dim"], bytes_per_frame=t["bytes_per_frame"],
                                        device=torch.device("cpu"), phase_one_iters=t["phase_one_iters"],
                                        phase_two_iters=t["phase_two_iters"])
        logged = []

        class Grab(logging.Handler):
            def emit(self, rec):
                m = rec.getMessage()
                if "loss_per_iter=" in m:
                    logged.append(json.loads(m.split("loss_per_iter=")[1].split("]")[0] + "]"))
        h = Grab()
        logging.getLogger().addHandler(h)
        logging.getLogger().setLevel(logging.INFO)
        t0 = time.time()
        while not trainer.done():
            trainer.step(gen_x(t["B"]))
            if trainer.cur_iter % 500 == 0:
                print(kind, "iter", trainer.cur_iter, "%.0f s" % (time.time() - t0), flush=True)
        logging.getLogger().removeHandler(h)
        q = trainer.get_quantizer()
        x_mean = q.get_data_mean()
        errs = []
        with torch.no_grad():
            for _ in range(t["eval_batches"]):  # test_quantization.py:41-46
                x = gen_x(t["B"])
                xa = q.decode(q.encode(x))
                errs.append(float(((x - xa) ** 2).sum() / ((x - x_mean) ** 2).sum()))
        meta[f"{kind}_avg_rel_err"] = float(np.mean(errs))
        meta[f"{kind}_rel_err_std_over_batches"] = float(np.std(errs))
        out[f"{kind}/loss_per_iter"] = np.array(logged, dtype=np.float32)
        print(kind, "avg rel err", meta[f"{kind}_avg_rel_err"], "logged rows", len(logged), flush=True)
        if kind == "mlp":
            for k, v in q.state_dict().items():
                out[f"mlp/state/{k}"] = v.detach().numpy()
            xe = gen_x(2048)
            with torch.no_grad():
                out["mlp/x_eval"] = xe.numpy()
                out["mlp/codes"] = q.encode(xe).numpy()
    rate = t["bytes_per_frame"] * 8 / t["dim"]
    meta["shannon_distortion"] = 2 ** -(2 * rate)  # test_quantization.py:55-60
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_trainer_quality.npz"), **out)


def make_misc():
    out, meta = {}, {}
    for name, (D, N, K, seed) in {"k16_n8_d64": (64, 8, 16, 3), "k16_n4_d40": (40, 4, 16, 4),
                                  "k256_n4_d96": (96, 4, 256, 5)}.items():
        p = synth.synth_params(D, N, K, seed)
        q = ref_quantizer(D, N, K, p)
        with torch.no_grad():
            q.centers_scale.fill_(0.02)
            q.logits_scale.fill_(-0.01)
            out[f"{name}/correlations"] = q.compute_codebook_correlations().numpy()
            if K == 16:
                pq = q.get_product_quantizer()
                out[f"{name}/pq_weight_sha"] = np.frombuffer(synth.sha256_of(pq.to_logits.weight).encode(), dtype=np.uint8)
                out[f"{name}/pq_bias_sha"] = np.frombuffer(synth.sha256_of(pq.to_logits.bias).encode(), dtype=np.uint8)
                out[f"{name}/pq_centers_sha"] = np.frombuffer(synth.sha256_of(pq.centers).encode(), dtype=np.uint8)
                out[f"{name}/pq_scales"] = np.array([float(pq.logits_scale), float(pq.centers_scale)], dtype=np.float32)
                x = synth.synth_x(256, D, 77)
                out[f"{name}/pq_codes"] = pq.encode(x, refine_indexes_iters=2).numpy()
        meta[name] = dict(D=D, N=N, K=K, seed=seed, centers_scale=0.02, logits_scale=-0.01)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_misc.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", "8")))
    if what in ("misc", "all"):
        make_misc()
    if what in ("control", "all"):
        make_control()
    if what in ("trainer", "all"):
        make_trainer()
