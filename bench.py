#!/usr/bin/env python3
"""bench.py -- encode throughput of the multi-codebook hot path (BASELINE.json metric:
"encode Mvectors/sec at dim=512, 8 codebooks; reconstruction MSE vs ref").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one Quantizer.encode (classifier arg-max + 5 refinement passes, uint8 codes out) over one batch of
2^20 synthetic Gaussian frames per GPU (BASELINE configs[1]: dim=512, bytes_per_frame=8, batch=1M fp32).
`value` times the step with the frames already resident in HBM; `e2e` times Quantizer.encode_host() on pinned HOST
buffers (H2D of the frames and D2H of the codes inside the timed region).  With N > 1 every rank encodes its own
1M-frame shard (weak scaling) and the step ends with the NCCL all-gather of the uint8 codes.
The line also carries: `roofline` (search kernel, SURVEY 8d flops on the passes actually executed) and `roofline_l1`
(its real bound), `cpu_baseline` (the unmodified reference on the host cores; the C port beside it as `cpu_port`),
`reference_gpu_pytorch` (the unmodified reference on the same GPU, best of a sub-batch sweep) and a `configs` block
with the other BASELINE configs (C1 ms/call, C3 trainer steps, C4 8M sharded encode, C5 fp16 round trip) at every N.
--impl reference times the reference's own CPU path (baseline/_ref PyTorch, else the C port in oracle/) instead.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "encode Mvectors/sec at dim=512, 8 codebooks"
UNIT = "Mvectors/s"
DIM, NCB, KSZ, ITERS = 512, 8, 256, 5
FRAMES = 1 << 20
CPU_SAMPLE = 4096      # frames of the C-port CPU sample
# frames per step of the unmodified PyTorch reference on the host cores (~0.5-1 kvec/s)
REF_CPU_SAMPLE = int(os.environ.get("MCQ_BENCH_REF_SAMPLE", "1024"))
# reference-formulation flops per frame at this config:
#   classifier GEMM 2*D*N*K = 2.097 MFLOP once; per refinement pass: dense scoring 2*D*N*K = 2.097 MFLOP plus the
#   per-frame pair products D * sum_levels 2*(N/2^l)*Kc_l^2 = 512 * 2*(4*256 + 2*256 + 1*1024) = 2.621 MFLOP
#   (5 passes: 12.583 + 13.107 = 25.69 MFLOP)
FLOP_PER_FRAME_INIT = 2.0 * DIM * NCB * KSZ  # classifier GEMM (not part of the search kernel)
FLOP_DENSE_PER_PASS = 2.0 * DIM * NCB * KSZ
FLOP_PAIR_PER_PASS = DIM * 2.0 * (4 * 256 + 2 * 256 + 1 * 1024)
BYTES_PER_FRAME_ENCODE = DIM * 4 + NCB       # algorithmic HBM bytes per frame (x in, codes out)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sust=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in a thread (nvidia_ml_py: the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints, every 20 ms, no process start-up
    inside the window); if NVML cannot be loaded, an `nvidia-smi -lms 100` loop over the same window."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = None
        self.p = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()
        self.source = None

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            if self.stop_flag:
                break
            time.sleep(0.02)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.idx
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi"
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=5)
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            self.f.seek(0)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self.f.read().splitlines():
                c = [v.strip() for v in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    self.sm.append(float(c[1]))
                    self.mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            try:
                os.unlink(self.f.name)
            except OSError:
                pass
        if self.sm:
            out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons),
                   "samples": len(self.sm), "source": self.source}
        return out


def host_threads():
    """Threads the CPU arm uses: every core this process may run on.  (torchrun exports OMP_NUM_THREADS=1, which would
    silently make the 'all host threads' baseline single-threaded, so the count is passed explicitly.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def import_reference():
    """The UNMODIFIED reference package (pip-installed into the git-ignored baseline/_ref), or None.
    The only shim is an empty `h5py` module: the reference imports it for read_hdf5_data alone."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "quantization")):
        return None
    import types
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, ref_dir)
    try:
        import quantization as refq
        return refq
    except Exception:  # pragma: no cover
        return None
    finally:
        sys.path.remove(ref_dir)


def reference_quantizer(refq, params, D, N, K, dev):
    import torch
    q = refq.Quantizer(dim=D, codebook_size=K, num_codebooks=N)
    with torch.no_grad():
        q.centers.copy_(params["centers"])
        q.to_logits.weight.copy_(params["weight"])
        q.to_logits.bias.copy_(params["bias"])
    return q.to(dev)


def cpu_port_run(x_np, params, threads=0):
    """One pass of the CPU port (oracle) over x_np.  Returns (seconds, indexes)."""
    import oracle
    t0 = time.perf_counter()
    idx = oracle.compute_indexes(x_np, params["centers"].numpy(), params["weight"].numpy(), params["bias"].numpy(),
                                 iters=ITERS, threads=threads)
    return time.perf_counter() - t0, idx


def reference_cpu_pytorch_run(refq, x_host, params, threads):
    """One Quantizer.encode of the unmodified reference on the host cores (fp32, no_grad).  (seconds, codes)."""
    import torch
    torch.set_num_threads(threads)
    q = reference_quantizer(refq, params, DIM, NCB, KSZ, torch.device("cpu"))
    with torch.no_grad():
        t0 = time.perf_counter()
        codes = q.encode(x_host, refine_indexes_iters=ITERS)
        return time.perf_counter() - t0, codes


def reference_gpu_leg(refq, x_host, params, dev, frames=65536):
    """The unmodified reference on the same GPU: its own Quantizer.encode, fp32, TF32 off (PyTorch default),
    torch.no_grad(), `frames` frames in sub-batches of 16,384 / 32,768 / 65,536 (its (B, N, 16, dim) fp32 temporaries
    do not allow the full 2^20-frame batch), 3 timed repetitions each; the BEST sub-batch size is the denominator of
    north_star's ">= 10x the reference GPU PyTorch path"."""
    import torch
    q = reference_quantizer(refq, params, DIM, NCB, KSZ, dev)
    x = x_host[:frames].to(dev)
    sweep, best, codes = {}, None, None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for sub in (16384, 32768, 65536):
        try:
            with torch.no_grad():
                def enc():
                    return torch.cat([q.encode(x[i:i + sub], refine_indexes_iters=ITERS) for i in range(0, frames, sub)])
                c = enc()
                torch.cuda.synchronize()
                reps = 3
                e0.record()
                for _ in range(reps):
                    c = enc()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            sweep[str(sub)] = {"ms": ms, "Mvectors_per_s": frames / (ms * 1e-3) / 1e6}
            if best is None or ms < best[1]:
                best, codes = (sub, ms), c
        except Exception as e:  # e.g. out of memory at the largest sub-batch: the smaller ones stand
            sweep[str(sub)] = {"failed": f"{type(e).__name__}: {e}"[:160]}
            torch.cuda.empty_cache()
    if best is None:
        return {"unavailable": "every sub-batch size failed", "sweep": sweep}
    return {"value": frames / (best[1] * 1e-3) / 1e6, "unit": UNIT, "frames": frames, "ms": best[1],
            "best_sub_batch": best[0], "sweep": sweep, "reps": 3,
            "how": "reference Quantizer.encode from baseline/_ref on cuda, fp32, no_grad; best of the sub-batch sweep",
            "codes": codes}


def run_reference(args, rank, world):
    """The reference arm: the reference's OWN implementation of the path on the host cores -- the unmodified PyTorch
    package from baseline/_ref on CPU (kind "reference") when it is there, else the C port in oracle/ (kind "port").
    Each step is a bounded sample (a prefix of the 2^20-frame batch) so that K + W steps end within minutes."""
    if rank != 0:
        return
    from quantization_b200 import synth
    params = synth.synth_params(DIM, NCB, KSZ, 0)
    cores = host_threads()
    refq = import_reference()
    if refq is not None:
        import torch
        kind, sample = "reference", REF_CPU_SAMPLE
        x = synth.synth_x(sample, DIM, 1234 + 1)
        for _ in range(args.warmup):
            reference_cpu_pytorch_run(refq, x, params, cores)
        times = [reference_cpu_pytorch_run(refq, x, params, cores)[0] for _ in range(args.steps)]
        how = (f"unmodified reference Quantizer.encode (baseline/_ref, PyTorch {torch.__version__} CPU, "
               f"torch.set_num_threads({cores}))")
        # the C port beside it, same sample (it is the faster CPU implementation; reported, not the denominator)
        psec = cpu_port_run(x.numpy(), params, cores)[0]
        port = {"value": sample / psec / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{sample}-frame prefix, one pass of oracle/mcq_oracle.c (OpenMP)"}
    else:
        kind, sample = "port", CPU_SAMPLE
        x = synth.synth_x(sample, DIM, 1234 + 1).numpy()
        for _ in range(args.warmup):
            cpu_port_run(x, params, cores)
        times = [cpu_port_run(x, params, cores)[0] for _ in range(args.steps)]
        how, port = "oracle/mcq_oracle.c (OpenMP); baseline/_ref is not importable here", None
    sec = sum(times) / len(times)
    value = sample / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: dim=512, bytes_per_frame=8 (8 x 256), refine_indexes_iters=5, fp32; "
                               f"each step = a {sample}-frame sample of the 2^20-frame batch",
                   "frames_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample}-frame prefix of the 2^20-frame batch, {args.steps} repetitions, {how}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if port is not None:
        line["cpu_port"] = port
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs (configs[0], [2], [3], [4]), measured in the same run and carried in the JSON line's
# `configs` block at every N.  The headline (`value`, `e2e`, `roofline`) stays configs[1].

def _timed(fn, reps, world, dev):
    """ms per call of fn over `reps` calls: CUDA events, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(reps):
        out = fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


def make_ours(D, N, K, params, dev):
    import torch
    from quantization_b200 import Quantizer
    q = Quantizer(dim=D, codebook_size=K, num_codebooks=N)
    with torch.no_grad():
        q.centers.copy_(params["centers"])
        q.to_logits.weight.copy_(params["weight"])
        q.to_logits.bias.copy_(params["bias"])
    return q.to(dev)


def config_c1(dev):
    """configs[0]: dim=256, bytes_per_frame=4, batch=4096 fp32 -- the reference's own CPU-runnable case.  ms per
    encode call (launch-latency bound at this size) and the code-exact check against the CPU oracle."""
    import numpy as np
    import torch

    import oracle
    from quantization_b200 import synth
    D, N, K, B = 256, 4, 256, 4096
    p = synth.synth_params(D, N, K, 0)
    q = make_ours(D, N, K, p, dev)
    x_host = synth.synth_x(B, D, 1234)
    x = x_host.to(dev)
    for _ in range(5):
        codes = q.encode(x)
    ms, codes = _timed(lambda: q.encode(x), 50, 1, dev)
    with torch.no_grad():
        dms, dec = _timed(lambda: q.decode(codes), 50, 1, dev)
    ref = oracle.encode(x_host.numpy(), p["centers"].numpy(), p["weight"].numpy(), p["bias"].numpy(), iters=ITERS)
    dref = oracle.decode(ref, p["centers"].numpy())
    return {"workload": "configs[0]: dim=256, bytes_per_frame=4, batch=4096 fp32", "encode_ms_per_call": ms,
            "encode_Mvectors_per_s": B / ms / 1e3, "decode_ms_per_call": dms,
            "frames_with_different_codes_vs_cpu_oracle": int((codes.cpu().numpy() != ref).any(1).sum()),
            "decode_bit_exact_vs_cpu_oracle": bool(np.array_equal(dec.cpu().numpy(), dref))}


def _trainer_phase_times(tr, x, steps, warm):
    """(ms per regular step, ms of one diagnostics step) of the trainer's current phase.  cur_iter is kept clear of the
    every-200-iterations diagnostics during the regular steps (it is stepped over a multiple of 200); the diagnostics
    step is run once untimed first -- its first evaluation in a phase pays one-time costs (new workspaces, lazily
    loaded kernels) that a 20,001-step run pays once, not 50 times -- and then timed."""
    import torch
    base = tr.cur_iter - tr.cur_iter % 200
    tr.cur_iter = base + 1
    for _ in range(warm):
        tr.step(x)
    tr.cur_iter = base + 400
    tr.step(x)  # diagnostics warm-up
    tr.cur_iter = base + 201
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(steps):
        if tr.cur_iter % 200 == 0:
            tr.cur_iter += 1
        tr.step(x)
    e1.record()
    tr.cur_iter = base + 400
    tr.step(x)  # cur_iter % 200 == 0: six extra compute_loss evaluations + the log line (reference :655-675)
    e2.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, e1.elapsed_time(e2)


def _trainer_run(trainer_cls, x, dev, steps, warm):
    import random

    import torch
    torch.manual_seed(1)
    random.seed(1)
    tr = trainer_cls(dim=256, bytes_per_frame=4, device=dev, phase_one_iters=10000, phase_two_iters=10000)
    p1, d1 = _trainer_phase_times(tr, x, steps, warm)
    tr.cur_iter = tr.phase_one_iters
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.step(x)  # the phase switch (get_product_quantizer, new optimiser)
    torch.cuda.synchronize()
    switch_ms = (time.perf_counter() - t0) * 1e3
    p2, d2 = _trainer_phase_times(tr, x, steps, warm)
    # iterations 0..20000: 10,001 in phase 1 (51 with diagnostics), 10,000 in phase 2 (50 with diagnostics)
    total_s = (9950 * p1 + 51 * d1 + 9950 * p2 + 50 * d2 + switch_ms) / 1e3
    return {"phase1_ms_per_step": p1, "phase1_diagnostics_step_ms": d1, "phase2_ms_per_step": p2,
            "phase2_diagnostics_step_ms": d2, "phase_switch_ms": switch_ms, "steps_timed_per_phase": steps,
            "est_seconds_for_20001_steps": total_s}


def config_c3(dev, with_reference):
    """configs[2]: dim=256, bytes_per_frame=4, batch=65536 bf16, QuantizerTrainer.step: 200 steps per phase (replayed
    from CUDA graphs after the warm-up) + one diagnostics step per phase + the phase switch, scaled to the 20,001-step
    loop.  The unmodified reference trainer from baseline/_ref is timed the same way (fewer steps: it is ~30x slower;
    it needs the frames up-cast to fp32, quantization.py:277-279 raises on bf16 with fp32 parameters)."""
    import logging

    import torch
    from quantization_b200 import QuantizerTrainer, synth
    x = synth.synth_x(65536, 256, 1234 + 2, torch.bfloat16).to(dev)
    logging.getLogger().setLevel(logging.WARNING)
    out = {"workload": "configs[2]: dim=256, bytes_per_frame=4, batch=65536 bf16, QuantizerTrainer.step "
                       "(phase 1: 8 x 16, phase 2: 4 x 256); the same batch every step"}
    out["ours"] = _trainer_run(QuantizerTrainer, x, dev, 200, 24)
    if with_reference:
        refq = import_reference()
        if refq is not None:
            try:
                out["reference_gpu_pytorch"] = _trainer_run(refq.QuantizerTrainer, x.float(), dev, 10, 2)
                out["speedup_over_reference_gpu"] = (out["reference_gpu_pytorch"]["est_seconds_for_20001_steps"] /
                                                     out["ours"]["est_seconds_for_20001_steps"])
            except Exception as e:
                out["reference_gpu_pytorch"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    return out


def config_c4(dev, rank, world):
    """configs[3]: dim=1024, bytes_per_frame=16, batch=8M fp32 sharded over the ranks (rank r encodes its 8M/N rows),
    one NCCL all-gather of the uint8 codes.  One timed pass (strong scaling: the total is fixed).  The frames are
    generated on the device (32 GiB at N=1)."""
    import torch
    from quantization_b200 import dist as qdist
    from quantization_b200 import synth
    D, N, K, total = 1024, 16, 256, 8 << 20
    p = synth.synth_params(D, N, K, 0)
    q = make_ours(D, N, K, p, dev)
    a, b = qdist.shard_rows(total, world, rank)
    g = torch.Generator(device=dev).manual_seed(4000 + rank)
    x = torch.randn(b - a, D, device=dev, generator=g, dtype=torch.float32)
    q._prepared()
    q.encode(x[:75776])  # warm-up on one chunk (the pass below is 3 s at N=1)
    if world > 1:
        qdist.sharded_encode(q, x[:4096], 4096 * world)  # and of the all-gather

    def step():
        return qdist.sharded_encode(q, x, total) if world > 1 else q.encode(x)
    # one timed pass at N = 1 (3 s); with more ranks the shard is short, so the first full-size pass (the first
    # all-gather of the full 128 MB, allocator growth) is run untimed and the second one is the measurement
    if world > 1:
        step()
    ms, codes = _timed(step, 1, world, dev)
    out = {"workload": "configs[3]: dim=1024, bytes_per_frame=16 (16 x 256), batch=8M fp32 sharded by rows over the "
                       "ranks + NCCL all-gather of the uint8 codes; frames generated on the device",
           "total_frames": total, "frames_per_gpu": b - a, "ms_per_pass": ms, "Mvectors_per_s": total / ms / 1e3,
           "scaling": "strong", "codes_shape": list(codes.shape),
           "codes_checksum_local_shard": int(codes[a:b].to(torch.int64).sum().item()) if world > 1
           else int(codes.to(torch.int64).sum().item())}
    del x, codes
    torch.cuda.empty_cache()
    return out


def config_c5(dev, rank, world):
    """configs[4]: dim=768, bytes_per_frame=8, batch=262144 fp16: encode + decode round trip of every rank's rows and
    the job-wide relative reconstruction error through one 2-scalar all-reduce; on a 4,096-frame sample the same error
    from the CPU oracle's codes (rank 0)."""
    import numpy as np
    import torch

    from quantization_b200 import dist as qdist
    from quantization_b200 import synth
    D, N, K, total = 768, 8, 256, 262144
    p = synth.synth_params(D, N, K, 0)
    q = make_ours(D, N, K, p, dev)
    a, b = qdist.shard_rows(total, world, rank)
    blk = 65536  # block i of 65,536 rows has seed 4000 + i: every world size sees the same frames
    parts = []
    for i in range(a // blk, (b + blk - 1) // blk):
        xb = synth.synth_x(min(blk, total - i * blk), D, 4000 + i, torch.float16)
        parts.append(xb[max(a, i * blk) - i * blk:min(b, (i + 1) * blk) - i * blk])
    x = torch.cat(parts).to(dev)
    q._prepared()
    for _ in range(2):
        qdist.sharded_round_trip_error(q, x)
    ms, (rel, codes) = _timed(lambda: qdist.sharded_round_trip_error(q, x), 5, world, dev)
    out = {"workload": "configs[4]: dim=768, bytes_per_frame=8, batch=262144 fp16, encode + decode round trip + "
                       "2-scalar all-reduce of the error sums", "total_frames": total, "frames_per_gpu": b - a,
           "ms_per_round_trip": ms, "Mvectors_per_s": total / ms / 1e3, "scaling": "strong", "rel_mse_ours": rel}
    if rank == 0:
        import oracle
        n = min(4096, x.shape[0])
        xs = x[:n].float().cpu().numpy()
        ref = oracle.encode(xs, p["centers"].numpy(), p["weight"].numpy(), p["bias"].numpy(), iters=ITERS)
        dref = oracle.decode(ref, p["centers"].numpy())
        x64 = xs.astype(np.float64)
        e_ref = float(((dref.astype(np.float64) - x64) ** 2).sum() / (x64 ** 2).sum())
        with torch.no_grad():
            d_ours = q.decode(codes[:n]).cpu().numpy()
        e_ours = float(((d_ours.astype(np.float64) - x64) ** 2).sum() / (x64 ** 2).sum())
        out["sample_4096"] = {"rel_mse_ours": e_ours, "rel_mse_cpu_oracle": e_ref,
                              "rel_difference": abs(e_ours - e_ref) / e_ref,
                              "frames_with_different_codes": int((codes[:n].cpu().numpy() != ref).any(1).sum())}
    return out


def other_configs(dev, rank, world):
    import torch.distributed as dist
    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # a failing side config must not take the headline down; it is reported as failed
            out[name] = {"failed": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:  # single-GPU configs: rank 0 measures them, the other ranks wait at the barrier below
        guarded("c1", lambda: config_c1(dev))
        guarded("c3", lambda: config_c3(dev, with_reference=(world == 1)))
    if world > 1:
        dist.barrier()
    guarded("c4", lambda: config_c4(dev, rank, world))
    guarded("c5", lambda: config_c5(dev, rank, world))
    return out


def search_kernel_profile():
    """Numbers of this round's `ncu --set full` capture of the search kernel (profiles/r02_search_ncu.json, written by
    tools/ncu_search_summary.py from the .ncu-rep): L1 wavefronts per frame-pass and dram bytes per launch."""
    for name in ("r02_search_ncu.json", "search_kernel_ncu.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
                d["file"] = "profiles/" + name
                return d
        except Exception:
            continue
    return {}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from quantization_b200 import _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B = args.frames
    params = synth.synth_params(DIM, NCB, KSZ, 0)
    q = make_ours(DIM, NCB, KSZ, params, dev)
    x_host = synth.synth_x(B, DIM, 1234 + 1 + rank)  # shard r of the job: its own seeded 2^20 frames
    x = x_host.to(dev)
    gathered = torch.empty(world * B, NCB, dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        codes = q.encode(x, refine_indexes_iters=ITERS, as_bytes=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, codes)  # the path's one exchange: uint8 codes only
        return codes

    q._prepared()  # parameter-version work (scaled centers, Gram table, operand splits) is not part of a step
    for _ in range(max(args.warmup, 3)):
        codes = step()
    torch.cuda.synchronize()
    ws = q._workspace(B)
    _lib.search_stats(ws, reset=True, read=False)

    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.profile(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        codes = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = _lib.profile_read()
    _lib.profile(False)
    clocks = sampler.stop()
    passes, frames_searched = _lib.search_stats(ws)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B / (ms_max * 1e-3) / 1e6

    # ---- end to end through the public API with HOST buffers (H2D + kernels + D2H inside the timed region)
    xp = x_host.pin_memory()
    out_host = torch.empty(B, NCB, dtype=torch.uint8, pin_memory=True)
    q.encode_host(xp, ITERS, True, out=out_host)
    e2e_steps = max(1, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(e2e_steps):
        q.encode_host(xp, ITERS, True, out=out_host)  # returns when the codes are in host memory
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * B / (e2e_ms * 1e-3) / 1e6
    same_codes = bool(torch.equal(out_host, codes.cpu()))
    # the host -> device copy of a step's frames ALONE, all ranks at once: what the host side of the box can feed.
    # (e2e can be no faster than this; at 8 ranks the eight 2 GiB streams share one host memory system.)
    xd = torch.empty_like(x)
    xd.copy_(xp, non_blocking=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        xd.copy_(xp, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d_alone_ms = float(t.item())
    del xp, xd

    # ---- decode (HBM-bound leg of the path), reported beside the encode number
    with torch.no_grad():
        for _ in range(3):
            dec = q.decode(codes)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            dec = q.decode(codes)
        e1.record()
        torch.cuda.synchronize()
    dec_ms = e0.elapsed_time(e1) / 10
    dec_bytes = B * (NCB + DIM * 4)
    del dec

    # ---- roofline of the dominant kernel (the search kernel), from the events recorded inside the timed region.
    # algorithmic flops are the REFERENCE formulation's, per pass ACTUALLY EXECUTED (the kernel counts
    # them: converged frames skip their remaining passes): dense scoring 2*D*N*K + per-frame pair products.
    s_ms, s_n = prof["search"]
    g_ms, g_n = prof["gemm"]
    o_ms, o_n = prof["other"]
    launches = s_n + g_n + o_n
    frames_per_launch = B * args.steps / max(s_n, 1)
    search_avg_ms = s_ms / max(s_n, 1)
    passes_per_frame = passes / max(frames_searched, 1)
    flop_search_frame = (FLOP_DENSE_PER_PASS + FLOP_PAIR_PER_PASS) * passes_per_frame
    achieved_tf = flop_search_frame * frames_per_launch / (search_avg_ms * 1e-3) / 1e12
    kp = search_kernel_profile()
    roofline = {"bound": "tensor", "kernel": "search2_kernel<8>", "achieved": achieved_tf, "peak": peaks["tf_sust"],
                "unit": "TFLOP/s", "frac": achieved_tf / peaks["tf_sust"],
                "traffic": kp.get("dram_bytes_per_launch"), "traffic_source": kp.get("file"),
                "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['src']})",
                "algorithmic_flop_per_frame_pass": FLOP_DENSE_PER_PASS + FLOP_PAIR_PER_PASS,
                "passes_executed_per_frame": passes_per_frame, "passes_requested": ITERS,
                "frame_passes_counted": passes, "frames_counted": frames_searched,
                "algorithmic_flop_per_frame": flop_search_frame,
                "frames_per_launch": frames_per_launch, "avg_launch_ms": search_avg_ms,
                "share_of_step": s_ms / args.steps / ms,
                "algorithmic_bytes_per_launch": BYTES_PER_FRAME_ENCODE * frames_per_launch,
                "formula": "frac = algorithmic_flop_per_frame_pass * passes_executed_per_frame * frames_per_launch "
                           "/ avg_launch_ms / peak",
                "note": "algorithmic flops are the reference formulation's (SURVEY 8d: dense 2DNK + pair products per "
                        "executed pass); the kernel replaces them by Gram-table look-ups and executes no MMA, so its "
                        "own bound is the L1 data pipe: see roofline_l1"}
    # the same step on the flops the Gram reformulation actually needs (two GEMMs per frame: 4*D*N*K), whole step
    step_reduced_tf = 4.0 * DIM * NCB * KSZ * B / (ms * 1e-3) / 1e12
    step_ref_tf = (FLOP_PER_FRAME_INIT + flop_search_frame) * B / (ms * 1e-3) / 1e12
    roofline["frac_reduced"] = step_reduced_tf / peaks["tf_sust"]
    roofline["whole_step"] = {"reference_formulation_tflops": step_ref_tf,
                              "frac_reference_formulation": step_ref_tf / peaks["tf_sust"],
                              "reduced_count_tflops": step_reduced_tf, "frac_reduced": step_reduced_tf / peaks["tf_sust"],
                              "reduced_flop_per_frame": 4.0 * DIM * NCB * KSZ}
    roofline_l1 = None
    wf = kp.get("l1_wavefronts_per_frame_pass")
    if wf and clocks.get("sm_mhz"):
        frame_passes_per_launch = passes / max(s_n, 1)
        ach = wf * frame_passes_per_launch / (search_avg_ms * 1e-3)
        peak = 148 * clocks["sm_mhz"] * 1e6  # one 128-byte wavefront per cycle per SM
        roofline_l1 = {"bound": "l1", "kernel": "search2_kernel<8>", "achieved": ach / 1e9, "peak": peak / 1e9,
                       "unit": "Gwavefront/s", "frac": ach / peak, "wavefronts_per_frame_pass": wf,
                       "wavefronts_source": kp.get("file"), "frame_passes_per_launch": frame_passes_per_launch,
                       "note": "LSU data-pipe wavefronts (global + shared) per frame-pass from this round's ncu "
                               "capture x frame-passes counted in this run / launch time, against 148 SMs x the SM "
                               "clock sampled in the timed region"}
    gemm_flop_exec = 3 * 2.0 * DIM * NCB * KSZ * frames_per_launch  # three fp16 products per GEMM launch
    extra = {
        "gemm": {"kernel": "gemm_fp16x2_kernel<128> (tcgen05, 3 MMAs per K step; the logits GEMM carries the fused arg-max epilogue)", "launches": g_n, "avg_launch_ms": g_ms / max(g_n, 1),
                 "executed_tflops": gemm_flop_exec / (g_ms / max(g_n, 1) * 1e-3) / 1e12 if g_n else None,
                 "frac_of_bf16_sustained": (gemm_flop_exec / (g_ms / max(g_n, 1) * 1e-3) / 1e12 / peaks["tf_sust"])
                 if g_n else None, "share_of_step": g_ms / args.steps / ms},
        "other_kernels": {"launches": o_n, "share_of_step": o_ms / args.steps / ms},
        "encode_hbm": {"algorithmic_GBps": B * BYTES_PER_FRAME_ENCODE / (ms * 1e-3) / 1e9,
                       "frac_of_hbm_peak": B * BYTES_PER_FRAME_ENCODE / (ms * 1e-3) / 1e9 / peaks["hbm"]},
        "decode": {"Mvectors_per_s": B / (dec_ms * 1e-3) / 1e6, "bound": "hbm",
                   "achieved_GBps": dec_bytes / (dec_ms * 1e-3) / 1e9,
                   "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / peaks["hbm"], "peak_GBps": peaks["hbm"]},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: dim=512, bytes_per_frame=8 (8 codebooks x 256), batch=2^20 Gaussian fp32 "
                               "frames per GPU, refine_indexes_iters=5, as_bytes=True; synthetic codebooks "
                               "centers~N(0,1/N), to_logits = (2c, -|c|^2)",
                   "frames_per_gpu": B, "l2": "inputs (2 GiB of frames per step) exceed the 126 MB L2; no flush needed",
                   "exchange": "NCCL all-gather of the uint8 codes" if world > 1 else "none (1 GPU)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * DIM * 4, "d2h_bytes_per_step": B * NCB,
                "ms_per_step": e2e_ms, "api": "Quantizer.encode_host -> mcq_encode_host_ws (pinned host buffers, PyTorch-owned device staging, current stream)",
                "codes_equal_device_path": same_codes,
                "h2d_alone_ms": h2d_alone_ms,
                "h2d_alone_GBps_per_rank": B * DIM * 4 / (h2d_alone_ms * 1e-3) / 1e9,
                "h2d_note": "h2d_alone_ms = the step's 2 GiB host->device copy with nothing else running, all ranks "
                            "at once (max over ranks): the floor the host memory system / PCIe sets for e2e"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": extra,
    }
    if roofline_l1 is not None:
        line["roofline_l1"] = roofline_l1

    if rank == 0 and world == 1:
        # CPU baselines + reconstruction error on a bounded sample of the same workload
        import numpy as np
        cores = host_threads()
        xs = x_host[:CPU_SAMPLE].numpy()
        cpu_port_run(xs[:512], params, cores)
        sec, ref_idx = cpu_port_run(xs, params, cores)
        port = {"value": CPU_SAMPLE / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{CPU_SAMPLE}-frame prefix of the batch, one pass of oracle/mcq_oracle.c (OpenMP, all host "
                          "threads)"}
        refq = import_reference()
        line["cpu_baseline"] = port
        if refq is not None:
            try:
                reference_cpu_pytorch_run(refq, x_host[:256], params, cores)
                rsec, rcodes = reference_cpu_pytorch_run(refq, x_host[:REF_CPU_SAMPLE], params, cores)
                line["cpu_baseline"] = {
                    "value": REF_CPU_SAMPLE / rsec / 1e6, "unit": UNIT, "cores": cores, "kind": "reference",
                    "sample": f"{REF_CPU_SAMPLE}-frame prefix of the batch, one Quantizer.encode of the unmodified "
                              f"reference (baseline/_ref, PyTorch {torch.__version__} on CPU, "
                              f"torch.get_num_threads()={torch.get_num_threads()}, os.cpu_count()={os.cpu_count()})",
                    "frames_with_different_codes_vs_ours": int(
                        (rcodes.numpy() != codes[:REF_CPU_SAMPLE].cpu().numpy()).any(1).sum())}
                line["cpu_port"] = port
            except Exception as e:
                line["cpu_port"] = {"note": f"reference CPU leg failed: {type(e).__name__}: {e}"[:200]}
        ours = codes[:CPU_SAMPLE].cpu().numpy().astype(np.int64)
        c64 = params["centers"].numpy().astype(np.float64)
        x64 = xs.astype(np.float64)

        def rel_err(ix):
            rec = sum(c64[n, ix[:, n]] for n in range(NCB))
            return float(((rec - x64) ** 2).sum() / (x64 ** 2).sum())
        if refq is not None:
            try:
                rg = reference_gpu_leg(refq, x_host, params, dev)
            except Exception as e:  # the reference's own failure must not take the bench line down
                rg = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            rc = rg.pop("codes", None)
            if rc is not None:
                n = rc.shape[0]
                rg["frames_with_different_codes_vs_ours"] = int((rc != codes[:n]).any(1).sum().item())
                rg["speedup_of_value_over_reference_gpu"] = value / rg["value"]
            line["reference_gpu_pytorch"] = rg
            torch.cuda.empty_cache()
        line["parity"] = {"sample_frames": CPU_SAMPLE,
                          "frames_with_different_codes": int((ours != ref_idx).any(1).sum()),
                          "rel_reconstruction_mse_ours": rel_err(ours), "rel_reconstruction_mse_ref": rel_err(ref_idx)}
    del x, codes, gathered
    torch.cuda.empty_cache()
    if not args.no_configs:
        line["configs"] = other_configs(dev, rank, world)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _route_stdout_to_stderr():
    """Everything libraries write to fd 1 (e.g. NCCL's "NCCL version ..." banner) goes to stderr, so that stdout
    carries exactly the one JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _route_stdout_to_stderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES, help="frames per GPU per step (default 2^20 = configs[1])")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
