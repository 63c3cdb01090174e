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
--impl reference times the CPU port of the reference algorithm (oracle/) on the host cores instead.
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
CPU_SAMPLE = 4096
# reference-formulation flops per frame at this config (5 refine passes):
#   dense scoring 2*D*N*K*(1+5) = 12.583 MFLOP, per-frame pair products 13.107 MFLOP  -> 25.69 MFLOP
FLOP_PER_FRAME_TOTAL = 25.69e6
FLOP_PER_FRAME_INIT = 2.0 * DIM * NCB * KSZ  # classifier GEMM (not part of the search kernel)
BYTES_PER_FRAME_ENCODE = DIM * 4 + NCB       # algorithmic HBM bytes per frame (x in, codes out)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sust=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "search_kernel_ncu.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


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


def cpu_port_run(x_np, params, threads=0):
    """One pass of the CPU port (oracle) over x_np.  Returns (seconds, indexes)."""
    import oracle
    t0 = time.perf_counter()
    idx = oracle.compute_indexes(x_np, params["centers"].numpy(), params["weight"].numpy(), params["bias"].numpy(),
                                 iters=ITERS, threads=threads)
    return time.perf_counter() - t0, idx


def reference_gpu_leg(x_host, params, dev, frames=65536):
    """The UNMODIFIED reference (pip-installed into the git-ignored baseline/_ref) on the same GPU:
    its own Quantizer.encode, fp32, TF32 off (PyTorch default), torch.no_grad(), one 65,536-frame chunk (its
    (B, N, 16, dim) fp32 temporaries do not allow the full 2^20-frame batch).  Informational: the denominator of
    north_star's "10x the reference GPU PyTorch path".  Returns None when baseline/_ref is not there."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "quantization")):
        return None
    import types

    import torch
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))  # the reference imports h5py only for read_hdf5_data
    sys.path.insert(0, ref_dir)
    try:
        import quantization as refq
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"import failed: {e}"}
    finally:
        sys.path.remove(ref_dir)
    q = refq.Quantizer(dim=DIM, codebook_size=KSZ, num_codebooks=NCB)
    with torch.no_grad():
        q.centers.copy_(params["centers"])
        q.to_logits.weight.copy_(params["weight"])
        q.to_logits.bias.copy_(params["bias"])
    q = q.to(dev)
    x = x_host[:frames].to(dev)
    sub = 16384  # keeps the reference's temporaries (256 KiB per frame, several live copies) well inside HBM
    with torch.no_grad():
        def enc():
            return torch.cat([q.encode(x[i:i + sub], refine_indexes_iters=ITERS) for i in range(0, frames, sub)])
        codes = enc()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 2
        for _ in range(reps):
            codes = enc()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": frames / (ms * 1e-3) / 1e6, "unit": UNIT, "frames": frames, "ms": ms,
            "how": "reference Quantizer.encode from baseline/_ref on cuda, fp32, no_grad, 16,384-frame sub-batches",
            "codes": codes}


def run_reference(args, rank, world):
    """The reference arm: the reference's algorithm on the host cores (CPU port in oracle/, all threads)."""
    if rank != 0:
        return
    import oracle
    from quantization_b200 import synth
    params = synth.synth_params(DIM, NCB, KSZ, 0)
    x = synth.synth_x(CPU_SAMPLE, DIM, 1234 + 1).numpy()
    cores = host_threads()
    for _ in range(args.warmup):
        cpu_port_run(x, params, cores)
    times = [cpu_port_run(x, params, cores)[0] for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = CPU_SAMPLE / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: dim=512, bytes_per_frame=8 (8 x 256), refine_indexes_iters=5, fp32; "
                               f"each step = a {CPU_SAMPLE}-frame sample of the 2^20-frame batch",
                   "frames_per_step": CPU_SAMPLE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{CPU_SAMPLE}-frame prefix of the 2^20-frame batch, {args.steps} repetitions, "
                                   "oracle/mcq_oracle.c (OpenMP)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from quantization_b200 import Quantizer, _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    B = args.frames
    params = synth.synth_params(DIM, NCB, KSZ, 0)
    q = Quantizer(dim=DIM, codebook_size=KSZ, num_codebooks=NCB)
    with torch.no_grad():
        q.centers.copy_(params["centers"])
        q.to_logits.weight.copy_(params["weight"])
        q.to_logits.bias.copy_(params["bias"])
    q = q.to(dev)
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

    # ---- roofline of the dominant kernel (the search kernel), from the events recorded inside the timed region
    s_ms, s_n = prof["search"]
    g_ms, g_n = prof["gemm"]
    o_ms, o_n = prof["other"]
    launches = s_n + g_n + o_n
    frames_per_launch = B * args.steps / max(s_n, 1)
    search_avg_ms = s_ms / max(s_n, 1)
    flop_search = (FLOP_PER_FRAME_TOTAL - FLOP_PER_FRAME_INIT) * frames_per_launch
    achieved_tf = flop_search / (search_avg_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "search2_kernel<8> (search.cu generic kernel for other shapes)", "achieved": achieved_tf, "peak": peaks["tf_sust"],
                "unit": "TFLOP/s", "frac": achieved_tf / peaks["tf_sust"], "traffic": load_traffic(),
                "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['src']})",
                "algorithmic_flop_per_frame": FLOP_PER_FRAME_TOTAL - FLOP_PER_FRAME_INIT,
                "frames_per_launch": frames_per_launch, "avg_launch_ms": search_avg_ms,
                "share_of_step": s_ms / args.steps / ms,
                "note": "algorithmic flops are the reference formulation's (SURVEY 8d); the kernel replaces them by "
                        "Gram-table lookups, so its own bound is the L1 data pipe (3.9 k wavefronts per frame-pass = "
                        "4.0 ms per 75,776-frame launch at 100 %; ncu: 77 % busy), not the tensor pipe (DESIGN.md 3)"}
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
                "ms_per_step": e2e_ms, "api": "Quantizer.encode_host -> mcq_encode_host (pinned host buffers)",
                "codes_equal_device_path": same_codes},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": extra,
    }

    if rank == 0 and world == 1:
        # CPU baseline + reconstruction error on a bounded sample of the same workload
        import numpy as np
        xs = x_host[:CPU_SAMPLE].numpy()
        import oracle
        cores = host_threads()
        cpu_port_run(xs[:512], params, cores)
        sec, ref_idx = cpu_port_run(xs, params, cores)
        ours = codes[:CPU_SAMPLE].cpu().numpy().astype(np.int64)
        c64 = params["centers"].numpy().astype(np.float64)
        x64 = xs.astype(np.float64)

        def rel_err(ix):
            rec = sum(c64[n, ix[:, n]] for n in range(NCB))
            return float(((rec - x64) ** 2).sum() / (x64 ** 2).sum())
        line["cpu_baseline"] = {"value": CPU_SAMPLE / sec / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{CPU_SAMPLE}-frame prefix of the batch, one pass of oracle/mcq_oracle.c "
                                          "(OpenMP, all host threads)"}
        try:
            rg = reference_gpu_leg(x_host, params, dev)
        except Exception as e:  # the reference's own failure must not take the bench line down
            rg = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        if rg is not None:
            rc = rg.pop("codes", None)
            if rc is not None:
                n = rc.shape[0]
                rg["frames_with_different_codes_vs_ours"] = int((rc != codes[:n]).any(1).sum().item())
                rg["speedup_of_value_over_reference_gpu"] = value / rg["value"]
            line["reference_gpu_pytorch"] = rg
        line["parity"] = {"sample_frames": CPU_SAMPLE,
                          "frames_with_different_codes": int((ours != ref_idx).any(1).sum()),
                          "rel_reconstruction_mse_ours": rel_err(ours), "rel_reconstruction_mse_ref": rel_err(ref_idx)}
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
