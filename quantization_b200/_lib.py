"""ctypes binding of quantization_b200/libmcq.so (C ABI declared in include/mcq.h).

There is no CPU fallback and no alternative backend: if the library is missing this module raises, and every
entry point raises when a call fails.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCQ_LIB") or os.path.join(_HERE, "libmcq.so")  # MCQ_LIB: A/B builds of the same ABI

F32, F16, BF16 = 0, 1, 2
U8, I64, I32 = 0, 1, 2

_X_DTYPES = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}
_IDX_DTYPES = {torch.uint8: U8, torch.int64: I64, torch.int32: I32}


class McqError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m quantization_b200.build` "
            "(nvcc, sm_100a).  quantization_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, sz, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float
    sigs = {
        "mcq_version": (i32, []),
        "mcq_last_error": (ctypes.c_char_p, []),
        "mcq_packed_cols": (i32, [i32, i32]),
        "mcq_prepared_bytes": (sz, [i32, i32, i32]),
        "mcq_workspace_bytes": (sz, [i64, i32, i32, i32]),
        "mcq_prepare": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, sz, vp]),
        "mcq_encode": (i32, [vp, i32, i64, i32, i32, i32, vp, i32, vp, i32, vp, sz, vp]),
        "mcq_refine": (i32, [vp, i32, i64, i32, i32, i32, vp, i32, vp, vp, vp, sz, vp]),
        "mcq_decode": (i32, [vp, i32, i64, i32, i32, i32, i32, vp, vp, i32, vp]),
        "mcq_decode_centers": (i32, [vp, i32, i64, i32, i32, i32, i32, vp, vp, i32, vp]),
        "mcq_decode_backward": (i32, [vp, vp, i64, i32, i32, i32, vp, vp]),
        "mcq_class_loss_forward": (i32, [vp, i32, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
        "mcq_class_loss_partials": (i32, []),
        "mcq_class_loss_backward": (i32, [vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
        "mcq_index_counts": (i32, [vp, i64, i32, i32, vp, vp, vp]),
        "mcq_column_sum_partials": (i32, [i32]),
        "mcq_column_sums": (i32, [vp, i64, i32, vp, vp, vp]),
        "mcq_prepared_scaled_centers": (vp, [vp, i32, i32, i32]),
        "mcq_prepared_gram": (vp, [vp, i32, i32, i32]),
        "mcq_xct": (i32, [vp, i32, i64, i32, i32, i32, vp, vp, vp, sz, vp]),
        "mcq_search": (i32, [vp, vp, i64, i32, i32, i32, vp, vp, vp]),
        "mcq_encode_host": (i32, [vp, i32, i64, i32, i32, i32, vp, i32, vp, i32, i32]),
        "mcq_encode_host_ws_bytes": (sz, [i64, i32, i32, i32, i32, i32]),
        "mcq_encode_host_ws": (i32, [vp, i32, i64, i32, i32, i32, vp, i32, vp, i32, vp, sz, vp]),
        "mcq_recon_loss_partials": (i32, []),
        "mcq_recon_loss_forward": (i32, [vp, i32, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp]),
        "mcq_recon_loss_backward": (i32, [vp, i32, vp, i64, i32, i32, i32, vp, vp, vp, vp]),
        "mcq_gemm_tn_workspace_bytes": (sz, [i64, i32, i32]),
        "mcq_gemm_tn": (i32, [vp, i64, vp, i32, i64, i64, i32, i32, vp, vp, sz, vp]),
        "mcq_gemm_nt_workspace_bytes": (sz, [i64, i32, i32]),
        "mcq_gemm_nt": (i32, [vp, i64, vp, i64, i64, i32, i32, vp, i64, i32, vp, sz, vp]),
        "mcq_jcl_hidden_forward": (i32, [vp, vp, i32, i64, i32, i32, i32, vp, f32, vp, vp]),
        "mcq_jcl_hidden_backward": (i32, [vp, vp, vp, i32, i64, i32, i32, i32, f32, vp, vp, vp]),
        "mcq_jcl_partials": (i32, []),
        "mcq_jcl_cross_entropy": (i32, [vp, vp, vp, i32, i64, i32, i32, i64, i32, vp, vp, vp, vp]),
        "mcq_profile": (i32, [i32]),
        "mcq_profile_read": (i32, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
        "mcq_search_stats": (i32, [vp, i32, ctypes.POINTER(ctypes.c_uint64), vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError here == the library does not export what mcq.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTS = ["mcq_version", "mcq_last_error", "mcq_packed_cols", "mcq_prepared_bytes", "mcq_workspace_bytes",
           "mcq_prepare", "mcq_encode", "mcq_refine", "mcq_decode", "mcq_decode_centers", "mcq_decode_backward",
           "mcq_class_loss_forward", "mcq_class_loss_backward", "mcq_class_loss_partials",
           "mcq_index_counts", "mcq_column_sum_partials", "mcq_column_sums",
           "mcq_prepared_scaled_centers", "mcq_prepared_gram", "mcq_xct", "mcq_search", "mcq_encode_host",
           "mcq_encode_host_ws_bytes", "mcq_encode_host_ws",
           "mcq_recon_loss_partials", "mcq_recon_loss_forward", "mcq_recon_loss_backward",
           "mcq_gemm_tn_workspace_bytes", "mcq_gemm_tn", "mcq_gemm_nt_workspace_bytes", "mcq_gemm_nt",
           "mcq_jcl_hidden_forward", "mcq_jcl_hidden_backward", "mcq_jcl_partials", "mcq_jcl_cross_entropy",
           "mcq_profile", "mcq_profile_read", "mcq_search_stats"]


def check(rc, what):
    if rc != 0:
        msg = lib().mcq_last_error().decode(errors="replace")
        raise McqError(f"{what} failed ({rc}): {msg}")


def x_dtype_code(t: torch.Tensor) -> int:
    try:
        return _X_DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported dtype {t.dtype}: expected float32, float16 or bfloat16") from None


def idx_dtype_code(t: torch.Tensor) -> int:
    try:
        return _IDX_DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported index dtype {t.dtype}: expected uint8, int32 or int64") from None


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} is on {t.device}: quantization_b200 runs on CUDA (sm_100a) only and has no CPU fallback")


PROF_KINDS = ("other", "gemm", "search", "decode")


def profile(enable: bool):
    check(lib().mcq_profile(1 if enable else 0), "mcq_profile")


def profile_read():
    """{kind: (total_ms, launches)} of the kernels launched since profile(True)."""
    ms = (ctypes.c_double * 4)()
    n = (ctypes.c_int64 * 4)()
    check(lib().mcq_profile_read(ms, n), "mcq_profile_read")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(PROF_KINDS)}


def search_stats(workspace: torch.Tensor, reset: bool = False, read: bool = True):
    """(refinement passes executed, frames searched) accumulated in `workspace` since the last reset (mcq.h)."""
    out = (ctypes.c_uint64 * 2)()
    with torch.cuda.device(workspace.device):
        check(lib().mcq_search_stats(workspace.data_ptr(), 1 if reset else 0, out if read else None,
                                     stream_ptr(workspace.device)), "mcq_search_stats")
    return (int(out[0]), int(out[1])) if read else None


def gemm_tn(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a^T . b on the tensor cores (include/mcq.h: mcq_gemm_tn): a (R, C1) fp32 (a column slice of a row-major matrix is
    fine: stride(1) == 1), b (R, C2) fp32 / fp16 / bf16 -> (C1, C2) fp32."""
    L = lib()
    assert a.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2 and a.shape[0] == b.shape[0]
    assert a.stride(1) == 1 and b.stride(1) == 1
    R, C1 = a.shape
    C2 = b.shape[1]
    out = torch.empty(C1, C2, dtype=torch.float32, device=a.device)
    if R == 0:
        return out.zero_()
    nbytes = L.mcq_gemm_tn_workspace_bytes(R, C1, C2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        check(L.mcq_gemm_tn(a.data_ptr(), a.stride(0), b.data_ptr(), x_dtype_code(b), b.stride(0), R, C1, C2,
                            out.data_ptr(), ws.data_ptr(), nbytes, stream_ptr(a.device)), "mcq_gemm_tn")
    return out


def index_counts(idx: torch.Tensor, N: int, K: int) -> torch.Tensor:
    """(N, K) float32 histogram of the int64 indexes (B, N) per codebook (include/mcq.h: mcq_index_counts)."""
    L = lib()
    assert idx.dtype == torch.int64 and idx.is_contiguous() and idx.shape[1] == N
    counts = torch.empty(N, K, dtype=torch.float32, device=idx.device)
    scratch = torch.empty(N * K, dtype=torch.int32, device=idx.device)
    with torch.cuda.device(idx.device):
        check(L.mcq_index_counts(idx.data_ptr(), idx.shape[0], N, K, counts.data_ptr(), scratch.data_ptr(),
                                 stream_ptr(idx.device)), "mcq_index_counts")
    return counts


def column_sums(x: torch.Tensor) -> torch.Tensor:
    """x.sum(0) of a contiguous fp32 (R, C) matrix, C a multiple of 4, in a fixed order (mcq_column_sums)."""
    L = lib()
    assert x.dtype == torch.float32 and x.is_contiguous() and x.ndim == 2 and x.shape[1] % 4 == 0
    R, C = x.shape
    out = torch.empty(C, dtype=torch.float32, device=x.device)
    part = torch.empty(L.mcq_column_sum_partials(C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(L.mcq_column_sums(x.data_ptr(), R, C, out.data_ptr(), part.data_ptr(), stream_ptr(x.device)),
              "mcq_column_sums")
    return out


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None, accumulate: bool = False) -> torch.Tensor:
    """a . b^T on the tensor cores (include/mcq.h: mcq_gemm_nt): a (M, K), b (N, K) fp32 with unit inner stride (row
    slices / column blocks of wider matrices are fine) -> out (M, N) fp32, which may itself be a column block."""
    L = lib()
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.shape[1] == b.shape[1]
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == torch.float32
    nbytes = L.mcq_gemm_nt_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        check(L.mcq_gemm_nt(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, out.data_ptr(), out.stride(0),
                            int(accumulate), ws.data_ptr(), nbytes, stream_ptr(a.device)), "mcq_gemm_nt")
    return out
