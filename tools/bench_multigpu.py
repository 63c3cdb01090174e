#!/usr/bin/env python3
"""BASELINE configs[3] and configs[4] on N GPUs of one box (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/bench_multigpu.py --config c4|c5 [--frames TOTAL]
    python tools/bench_multigpu.py --config c5          # single GPU

c4: dim=1024, 16 codebooks, TOTAL (default 8M) fp32 frames sharded by rows, encode + NCCL all-gather of the uint8 codes.
c5: dim=768, 8 codebooks, TOTAL (default 262,144) fp16 frames: encode + decode round trip, job-wide relative
    reconstruction error through one 2-scalar all-reduce, compared on a 4,096-frame sample with the CPU oracle.
Strong scaling (TOTAL is fixed).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

from quantization_b200 import dist as qdist
from quantization_b200 import synth
from helpers import make_quantizer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c4", "c5"])
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.config == "c4":
        D, N, dt, total = 1024, 16, torch.float32, args.frames or (8 << 20)
    else:
        D, N, dt, total = 768, 8, torch.float16, args.frames or 262144
    p = synth.synth_params(D, N, 256, 0)
    q = make_quantizer(D, N, 256, p, dev)
    a, b = qdist.shard_rows(total, world, rank)
    # shard r of the job: rows [a, b) of the seeded batch, generated in blocks so that every world size sees the
    # same frames (block i of 65,536 rows has seed 4000 + i)
    blk = 65536
    parts = []
    for i in range(a // blk, (b + blk - 1) // blk):
        xb = synth.synth_x(min(blk, total - i * blk), D, 4000 + i, dt)
        lo, hi = max(a, i * blk) - i * blk, min(b, (i + 1) * blk) - i * blk
        parts.append(xb[lo:hi])
    x = torch.cat(parts).to(dev) if parts else torch.empty(0, D, dtype=dt, device=dev)
    q._prepared()

    def step():
        if args.config == "c4":
            return qdist.sharded_encode(q, x, total) if world > 1 else q.encode(x)
        return qdist.sharded_round_trip_error(q, x)

    out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    line = {"config": args.config, "dim": D, "num_codebooks": N, "dtype": str(dt).replace("torch.", ""),
            "total_frames": total, "n_gpus": world, "ms_per_pass": ms, "Mvectors_per_s": total / ms / 1e3,
            "scaling": "strong"}
    if args.config == "c4":
        codes = out
        line["codes_shape"] = list(codes.shape)
        line["codes_checksum"] = int(codes.to(torch.int64).sum().item())
    else:
        rel, codes = out
        line["round_trip_rel_error"] = rel
        if rank == 0:
            import numpy as np

            import oracle
            n = min(4096, x.shape[0])
            xs = x[:n].float().cpu().numpy()
            ref = oracle.encode(xs, p["centers"].numpy(), p["weight"].numpy(), p["bias"].numpy(), iters=5)
            dref = oracle.decode(ref, p["centers"].numpy())
            e_ref = float(((dref - xs).astype(np.float64) ** 2).sum() / (xs.astype(np.float64) ** 2).sum())
            with torch.no_grad():
                d_ours = q.decode(codes[:n]).cpu().numpy()
            e_ours = float(((d_ours - xs).astype(np.float64) ** 2).sum() / (xs.astype(np.float64) ** 2).sum())
            line["sample_4096"] = {"rel_error_ours": e_ours, "rel_error_cpu_oracle": e_ref,
                                   "rel_difference": abs(e_ours - e_ref) / e_ref,
                                   "frames_with_different_codes": int((codes[:n].cpu().numpy() != ref).any(1).sum())}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
