"""ctypes binding of oracle/libmcq_gram_model.so -- the bit-level CPU model of the CUDA search
kernel's arithmetic (see mcq_gram_model.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmcq_gram_model.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mcq_gram_model.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libmcq_gram_model.so"], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.mcq_gm_gram.argtypes = [f32p, ctypes.c_int, ctypes.c_int, f32p, ctypes.c_int]
        lib.mcq_gm_xct.argtypes = [f32p, ctypes.c_long, ctypes.c_int, f32p, ctypes.c_int, f32p, ctypes.c_int]
        lib.mcq_gm_search.argtypes = [f32p, f32p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, i64p, i64p,
                                      ctypes.c_int]
        for f in (lib.mcq_gm_gram, lib.mcq_gm_xct, lib.mcq_gm_search):
            f.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def gram(scaled_centers, threads=0):
    """(N,K,D) scaled centers -> G (N*K, N*K) float32, double accumulation rounded once."""
    c = _f32(scaled_centers)
    N, K, D = c.shape
    G = np.empty((N * K, N * K), np.float32)
    rc = _load().mcq_gm_gram(_p(c, ctypes.c_float), N * K, D, _p(G, ctypes.c_float), threads)
    assert rc == 0
    return G


def xct(x, scaled_centers, threads=0):
    """x (B,D), centers (N,K,D) -> P (B, N*K) float32."""
    x = _f32(x)
    c = _f32(scaled_centers)
    N, K, D = c.shape
    P = np.empty((x.shape[0], N * K), np.float32)
    rc = _load().mcq_gm_xct(_p(x, ctypes.c_float), x.shape[0], D, _p(c, ctypes.c_float), N * K, _p(P, ctypes.c_float),
                            threads)
    assert rc == 0
    return P


def search(P, G, idx_in, num_codebooks, codebook_size, iters, threads=0):
    """`iters` passes of the table-driven refinement.  Returns int64 (B,N)."""
    P = _f32(P)
    G = _f32(G)
    idx_in = np.ascontiguousarray(np.asarray(idx_in, dtype=np.int64))
    B = P.shape[0]
    N, K = int(num_codebooks), int(codebook_size)
    assert P.shape == (B, N * K) and G.shape == (N * K, N * K) and idx_in.shape == (B, N)
    out = np.empty((B, N), np.int64)
    rc = _load().mcq_gm_search(_p(P, ctypes.c_float), _p(G, ctypes.c_float), B, N, K, int(iters),
                               _p(idx_in, ctypes.c_int64), _p(out, ctypes.c_int64), threads)
    if rc:
        raise RuntimeError(f"mcq_gm_search failed: {rc}")
    return out
