"""ctypes binding of oracle/libmcq_oracle.so (plain-C restatement of the reference path).

TEST INFRASTRUCTURE ONLY -- see oracle/mcq_oracle.c for the parity statement and the
reference lines each function follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmcq_oracle.so")
_lib = None


class OracleError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make).  Returns the path of the shared library."""
    src = os.path.join(_HERE, "mcq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libmcq_oracle.so"], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(_SO)
    f32p = ctypes.POINTER(ctypes.c_float)
    i64p = ctypes.POINTER(ctypes.c_int64)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    lib.mcq_oracle_compute_indexes.argtypes = [
        f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, ctypes.c_float, f32p, f32p,
        ctypes.c_float, ctypes.c_float, ctypes.c_int, i64p, i64p, f32p, ctypes.c_int]
    lib.mcq_oracle_compute_indexes.restype = ctypes.c_int
    lib.mcq_oracle_decode.argtypes = [i64p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p,
                                      ctypes.c_float, ctypes.c_float, f32p, ctypes.c_int]
    lib.mcq_oracle_decode.restype = ctypes.c_int
    lib.mcq_oracle_pack.argtypes = [i64p, ctypes.c_long, ctypes.c_int, ctypes.c_int, u8p]
    lib.mcq_oracle_pack.restype = ctypes.c_int
    lib.mcq_oracle_unpack.argtypes = [i64p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, i64p]
    lib.mcq_oracle_unpack.restype = ctypes.c_int
    lib.mcq_oracle_max_threads.restype = ctypes.c_int
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _check(rc, what):
    if rc < 0:
        names = {-1: "invalid argument", -2: "unsupported (K < 16 with N > 1: the reference crashes too)",
                 -3: "out of memory"}
        raise OracleError(f"{what}: {names.get(rc, rc)}")


def max_threads() -> int:
    return int(_load().mcq_oracle_max_threads())


def compute_indexes(x, centers, weight, bias, centers_scale=0.0, logits_scale=0.0, scale_speed=10.0, iters=5,
                    idx_in=None, return_margin=False, threads=0):
    """Quantizer._compute_indexes (idx_in=None) or `iters` x Quantizer._refine_indexes (idx_in given).

    x (B,D) float32; centers (N,K,D); weight (N*K,D); bias (N*K).  Returns int64 (B,N)
    [and the per-frame minimum relative decision margin when return_margin]."""
    lib = _load()
    x = _f32(x)
    centers = _f32(centers)
    weight = _f32(weight)
    bias = _f32(bias)
    B, D = x.shape
    N, K, D2 = centers.shape
    assert D2 == D and weight.shape == (N * K, D) and bias.shape == (N * K,)
    out = np.empty((B, N), dtype=np.int64)
    margin = np.empty((B,), dtype=np.float32) if return_margin else None
    if idx_in is not None:
        idx_in = np.ascontiguousarray(np.asarray(idx_in, dtype=np.int64))
        assert idx_in.shape == (B, N)
    rc = lib.mcq_oracle_compute_indexes(
        _p(x, ctypes.c_float), B, D, N, K, _p(centers, ctypes.c_float), float(centers_scale),
        _p(weight, ctypes.c_float), _p(bias, ctypes.c_float), float(logits_scale), float(scale_speed), int(iters),
        _p(idx_in, ctypes.c_int64) if idx_in is not None else None, _p(out, ctypes.c_int64),
        _p(margin, ctypes.c_float) if margin is not None else None, int(threads))
    _check(rc, "mcq_oracle_compute_indexes")
    return (out, margin) if return_margin else out


def pack(idx, codebook_size):
    """The as_bytes=True branch of Quantizer.encode: (B,N) int64 -> (B,N_packed) uint8."""
    lib = _load()
    idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int64))
    B, N = idx.shape
    out = np.empty((B, N), dtype=np.uint8)
    cols = lib.mcq_oracle_pack(_p(idx, ctypes.c_int64), B, N, int(codebook_size), _p(out, ctypes.c_uint8))
    _check(cols, "mcq_oracle_pack")
    return out.reshape(-1)[: B * cols].reshape(B, cols).copy()


def unpack(packed, num_codebooks, codebook_size):
    """Quantizer._maybe_separate_indexes: (B,n) -> (B,N) int64."""
    lib = _load()
    packed = np.ascontiguousarray(np.asarray(packed, dtype=np.int64))
    B, n = packed.shape
    out = np.empty((B, num_codebooks), dtype=np.int64)
    rc = lib.mcq_oracle_unpack(_p(packed, ctypes.c_int64), B, n, int(num_codebooks), int(codebook_size),
                               _p(out, ctypes.c_int64))
    _check(rc, "mcq_oracle_unpack")
    return out


def encode(x, centers, weight, bias, centers_scale=0.0, logits_scale=0.0, scale_speed=10.0, iters=5,
           as_bytes=True, threads=0):
    """Quantizer.encode on a (B,D) batch."""
    idx = compute_indexes(x, centers, weight, bias, centers_scale, logits_scale, scale_speed, iters, threads=threads)
    return pack(idx, centers.shape[1]) if as_bytes else idx


def decode(indexes, centers, centers_scale=0.0, scale_speed=10.0, threads=0):
    """Quantizer.decode on (B,n) codes (packed or not).  Returns float32 (B,D)."""
    lib = _load()
    centers = _f32(centers)
    N, K, D = centers.shape
    idx = np.ascontiguousarray(np.asarray(indexes, dtype=np.int64))
    if idx.shape[1] != N:
        idx = unpack(idx, N, K)
    B = idx.shape[0]
    out = np.empty((B, D), dtype=np.float32)
    rc = lib.mcq_oracle_decode(_p(idx, ctypes.c_int64), B, D, N, K, _p(centers, ctypes.c_float), float(centers_scale),
                               float(scale_speed), _p(out, ctypes.c_float), int(threads))
    _check(rc, "mcq_oracle_decode")
    return out
