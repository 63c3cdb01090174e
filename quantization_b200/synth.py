"""Deterministic synthetic inputs shared by bench.py, the tests and the golden-fixture generator.

Follows BASELINE.md section 3: frames are `torch.randn(B, dim, generator=Generator().manual_seed(seed))`
generated on the CPU in fp32 (then cast), so that the CPU oracle and the GPU see identical bits;
the synthetic quantizer is centers ~ N(0, 1/N), to_logits.weight = 2 * centers, bias = -|c|^2.
"""
import hashlib
import math

import torch


def synth_x(num_frames: int, dim: int, seed: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(num_frames, dim, generator=g, dtype=torch.float32)
    return x.to(dtype)


def synth_params(dim: int, num_codebooks: int, codebook_size: int, seed: int = 0):
    """Returns dict(centers (N,K,D), weight (N*K,D), bias (N*K,)) float32 CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(num_codebooks, codebook_size, dim, generator=g, dtype=torch.float32)
    centers = centers * (1.0 / math.sqrt(num_codebooks))
    weight = (2.0 * centers).reshape(num_codebooks * codebook_size, dim).contiguous()
    bias = -(centers.double() ** 2).sum(-1).float().reshape(-1).contiguous()
    return {"centers": centers.contiguous(), "weight": weight, "bias": bias}


def synth_indexes(num_frames: int, num_codebooks: int, codebook_size: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, codebook_size, (num_frames, num_codebooks), generator=g, dtype=torch.int64)


def sha256_of(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        if isinstance(t, torch.Tensor):
            t = t.detach().cpu().contiguous()
            if t.dtype == torch.bfloat16:
                t = t.view(torch.int16)
            t = t.numpy()
        h.update(t.tobytes())
    return h.hexdigest()
