"""Frame-sharded encode across the GPUs of one box (one process per GPU, torch.distributed / NCCL).

Frames are independent (no cross-frame term anywhere in quantization.py:244-547), so rank r encodes the contiguous
row block [floor(r*B/G), floor((r+1)*B/G)) with its own replica of the (1-64 MB) prepared tables, and the only
exchange is one all-gather of the uint8 codes (B/G x N bytes per rank) -- no data-path collective on the frames.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(num_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Row range [start, stop) of rank `rank` out of `num_frames` rows split over `world_size` ranks."""
    assert 0 <= rank < world_size
    return (rank * num_frames) // world_size, ((rank + 1) * num_frames) // world_size


def all_gather_codes(codes_local: torch.Tensor, num_frames_total: int, group=None) -> torch.Tensor:
    """Concatenates every rank's (rows_r, ncols) uint8/int64 codes in rank order into (num_frames_total, ncols).
    Shards may be ragged (B not divisible by the world size): they are padded to the largest shard for the
    collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    ncols = codes_local.shape[1]
    sizes = [shard_rows(num_frames_total, world, r) for r in range(world)]
    rows = [b - a for a, b in sizes]
    assert codes_local.shape[0] == rows[dist.get_rank(group)], "local shard does not match shard_rows()"
    mx = max(rows)
    if all(r == mx for r in rows):
        out = torch.empty(world * mx, ncols, dtype=codes_local.dtype, device=codes_local.device)
        dist.all_gather_into_tensor(out, codes_local.contiguous(), group=group)
        return out
    padded = torch.zeros(mx, ncols, dtype=codes_local.dtype, device=codes_local.device)
    padded[:codes_local.shape[0]] = codes_local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, rows)], dim=0)


def sharded_encode(quantizer, x_local: torch.Tensor, num_frames_total: int, refine_indexes_iters: int = 5,
                   as_bytes: bool = True, group=None,
                   encode_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> torch.Tensor:
    """Every rank passes ITS row block of x (see shard_rows) and receives the codes of ALL frames.
    `encode_fn` defaults to quantizer.encode (the CUDA path); tests inject a stand-in to exercise the
    sharding / gather logic on CPU with the gloo backend."""
    if encode_fn is None:
        def encode_fn(t):
            return quantizer.encode(t, refine_indexes_iters=refine_indexes_iters, as_bytes=as_bytes)
    x2 = x_local.reshape(-1, x_local.shape[-1])
    codes_local = encode_fn(x2)
    return all_gather_codes(codes_local, num_frames_total, group=group)


def sharded_round_trip_error(quantizer, x_local: torch.Tensor, refine_indexes_iters: int = 5, group=None,
                             encode_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                             decode_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
    """encode -> decode round trip of every rank's row block and the job-wide relative reconstruction error
    sum((x - x_hat)^2) / sum(x^2) (BASELINE config 5: the MSE sweep over 1/2/4/8 GPUs).  The only exchange is one
    all-reduce of two float64 scalars; frames and codes stay on their rank.  Returns (rel_error, codes_local).
    `encode_fn` / `decode_fn` default to the quantizer's CUDA path; the CPU tests inject stand-ins."""
    if encode_fn is None:
        def encode_fn(t):
            return quantizer.encode(t, refine_indexes_iters=refine_indexes_iters, as_bytes=True)
    if decode_fn is None:
        def decode_fn(c):
            with torch.no_grad():
                return quantizer.decode(c)
    x2 = x_local.reshape(-1, x_local.shape[-1])
    codes = encode_fn(x2)
    xf = x2.to(torch.float32)
    err = decode_fn(codes).to(torch.float32) - xf
    # float64 partial sums: the result does not depend on how the frames are split over the ranks beyond 1e-12
    sums = torch.stack([(err.double() ** 2).sum(), (xf.double() ** 2).sum()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return float(sums[0] / sums[1]), codes
