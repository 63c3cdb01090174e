"""`Quantizer`: the reference's additive multi-codebook quantizer (danpovey/quantization,
quantization/quantization.py:16-573) with the hot path -- `encode`, `decode`, `_compute_indexes`,
`_refine_indexes` -- executed by the sm_100a kernels of libmcq.so instead of ATen ops.

The class keeps the reference's constructor, parameter names / state_dict keys (`centers`, `logits_scale`,
`centers_scale`, `id_buf`, `to_logits.weight`, `to_logits.bias`), RNG call order and method signatures, so it is a
drop-in for that path.  `compute_loss` runs in the library too (reconstruction term, classifier losses and their
gradients: recon.cu, loss.cu, gemm_tn.cu); only scalar / (N, K)-sized arithmetic and the diagnostics stay in PyTorch.
Extensions over the reference (documented in DESIGN.md): float16 / bfloat16 `x` is accepted (the reference raises on
mixed dtypes; the kernels up-convert exactly, so the result equals the reference on `x.float()`), and
`encode_host()` takes host tensors.
"""
import binascii
import ctypes
import math
import os
from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib


def _is_power_of_two(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


# smallest codebook_size that takes the fused classifier-loss kernels (below it: the PyTorch formulation)
_FUSED_LOSS_MIN_K = int(os.environ.get("MCQ_FUSED_LOSS_MIN_K", "16"))
_CHECK_INDEXES = os.environ.get("MCQ_CHECK_INDEXES", "0") == "1"
_SLAB_MIN_FRAMES = 16384  # csrc/decode.cu SLAB_MIN_FRAMES: from here on byte codes take the slab decode kernel
# largest shapes the search kernels cover (csrc/api.cu check_shape); the reference itself has no such limit
MAX_CODEBOOK_SIZE = 256
MAX_NUM_CODEBOOKS = 64


class _DecodeFn(torch.autograd.Function):
    """decode as a differentiable function of the scaled centers (reference: gather + sum, quantization.py:138-147)."""

    @staticmethod
    def forward(ctx, scaled_centers: Tensor, indexes: Tensor) -> Tensor:
        N, K, D = scaled_centers.shape
        B = indexes.shape[0]
        cs = scaled_centers.detach().contiguous()
        out = torch.empty(B, D, dtype=torch.float32, device=cs.device)
        L = _lib.lib()
        rc = L.mcq_decode_centers(indexes.data_ptr(), _lib.I64, B, N, N, K, D, cs.data_ptr(), out.data_ptr(),
                                  _lib.F32, _lib.stream_ptr(cs.device))
        _lib.check(rc, "mcq_decode_centers")
        ctx.save_for_backward(indexes)
        ctx.shape = (N, K, D)
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        (indexes,) = ctx.saved_tensors
        N, K, D = ctx.shape
        g = grad_out.contiguous().float()
        grad = torch.zeros(N, K, D, dtype=torch.float32, device=g.device)
        L = _lib.lib()
        rc = L.mcq_decode_backward(g.data_ptr(), indexes.data_ptr(), indexes.shape[0], N, K, D, grad.data_ptr(),
                                   _lib.stream_ptr(g.device))
        _lib.check(rc, "mcq_decode_backward")
        return grad, None


class _ReconLossFn(torch.autograd.Function):
    """(sum (x_hat - x)^2, sum (x - mean)^2) with x_hat = decode(indexes), differentiable in the scaled centers
    (reference: quantization.py:209-216).  One kernel forward (`mcq_recon_loss_forward`: no x_hat, no difference, no
    squares in memory), one kernel backward (`mcq_recon_loss_backward`: recomputes x_hat - x and scatter-adds)."""

    @staticmethod
    def forward(ctx, scaled_centers: Tensor, indexes: Tensor, x: Tensor, mean: Tensor) -> Tensor:
        N, K, D = scaled_centers.shape
        B = indexes.shape[0]
        L = _lib.lib()
        cs = scaled_centers.detach().contiguous()
        sums = torch.empty(2, dtype=torch.float32, device=cs.device)
        partials = torch.empty(L.mcq_recon_loss_partials(), dtype=torch.float32, device=cs.device)
        mean = mean.detach().to(torch.float32).contiguous()
        with torch.cuda.device(cs.device):
            rc = L.mcq_recon_loss_forward(x.data_ptr(), _lib.x_dtype_code(x), indexes.data_ptr(), B, N, K, D,
                                          cs.data_ptr(), mean.data_ptr(), sums.data_ptr(), partials.data_ptr(),
                                          _lib.stream_ptr(cs.device))
        _lib.check(rc, "mcq_recon_loss_forward")
        ctx.save_for_backward(cs, indexes, x)
        ctx.mark_non_differentiable(indexes)
        return sums

    @staticmethod
    def backward(ctx, g: Tensor):
        cs, indexes, x = ctx.saved_tensors
        N, K, D = cs.shape
        L = _lib.lib()
        coef = (2.0 * g[0]).to(torch.float32).reshape(1).contiguous()  # d sums[0] / d x_hat = 2 (x_hat - x); sums[1]: no grad
        grad = torch.zeros(N, K, D, dtype=torch.float32, device=cs.device)
        with torch.cuda.device(cs.device):
            rc = L.mcq_recon_loss_backward(x.data_ptr(), _lib.x_dtype_code(x), indexes.data_ptr(), indexes.shape[0], N, K,
                                           D, cs.data_ptr(), coef.data_ptr(), grad.data_ptr(), _lib.stream_ptr(cs.device))
        _lib.check(rc, "mcq_recon_loss_backward")
        return grad, None, None, None


class _ClassLossFn(torch.autograd.Function):
    """(logprob_sum, prob_sum) of the classifier logits as a differentiable function of the classifier parameters
    (reference: `_logits` -> log_softmax -> gather / exp -> mean, quantization.py:220-235).  The forward GEMM and the
    softmax reductions run in libmcq.so (`mcq_class_loss_forward`) without materialising log-softmax, probabilities
    or one-hot tensors; the backward forms d loss / d logits in one kernel (`mcq_class_loss_backward`) and the
    parameter gradients as the same two fp32 matrix products autograd forms for `nn.Linear`."""

    @staticmethod
    def forward(ctx, quantizer, x: Tensor, indexes: Tensor, weight: Tensor, bias: Tensor, logits_scale: Tensor):
        N, K, D = quantizer.num_codebooks, quantizer.codebook_size, quantizer.dim
        B = x.shape[0]
        L = _lib.lib()
        blob = quantizer._prepared()
        ws = quantizer._workspace(B)
        Bp = (B + 127) // 128 * 128
        xw = torch.empty(Bp, N * K, dtype=torch.float32, device=x.device)
        out = torch.empty(1 + N * K, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = L.mcq_class_loss_forward(x.data_ptr(), _lib.x_dtype_code(x), B, D, N, K, blob.data_ptr(),
                                          indexes.data_ptr(), xw.data_ptr(), out.data_ptr(), out[1:].data_ptr(),
                                          ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "mcq_class_loss_forward")
        ctx.quantizer = quantizer
        ctx.blob = blob
        ctx.save_for_backward(x, indexes, xw, weight, logits_scale)
        return out[0], out[1:].reshape(N, K)

    @staticmethod
    def backward(ctx, g_lp: Tensor, g_prob: Tensor):
        x, indexes, xw, weight, logits_scale = ctx.saved_tensors
        q = ctx.quantizer
        N, K, D = q.num_codebooks, q.codebook_size, q.dim
        B = x.shape[0]
        L = _lib.lib()
        g_lp = g_lp.reshape(1).float().contiguous()
        g_prob = g_prob.reshape(N * K).float().contiguous()
        grad_logits = torch.empty(B, N * K, dtype=torch.float32, device=x.device)
        part_gx = torch.empty(L.mcq_class_loss_partials(), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = L.mcq_class_loss_backward(xw.data_ptr(), B, D, N, K, ctx.blob.data_ptr(), indexes.data_ptr(),
                                           g_lp.data_ptr(), g_prob.data_ptr(), grad_logits.data_ptr(),
                                           part_gx.data_ptr(), _lib.stream_ptr(x.device))
        _lib.check(rc, "mcq_class_loss_backward")
        scale = (logits_scale.detach() * q.scale_speed).exp()
        # d loss / d weight = grad_logits^T . (scale * x)  (backward of quantization.py:278-279): split-K tcgen05
        # product over the frames (mcq_gemm_tn), the scalar applied to the (N*K, D) result
        grad_w = _lib.gemm_tn(grad_logits, x) * scale
        grad_b = _lib.column_sums(grad_logits)  # fixed-order column sums (the bias gradient of nn.Linear)
        # d/d logits_scale: logits = xs W^T + b with d xs / d logits_scale = speed * xs, so the gradient is
        # speed * sum(grad_logits * (xs W^T)); xs W^T is the saved forward product and the kernel above already
        # summed the products: no second GEMM, no extra pass
        grad_ls = part_gx.sum() * q.scale_speed
        return None, None, None, grad_w, grad_b, grad_ls.reshape(logits_scale.shape)


class Quantizer(nn.Module):
    def __init__(self, dim: int, codebook_size: int, num_codebooks: int):
        """Trainable quantizer encoding a `dim`-vector as `num_codebooks` integers in [0, codebook_size)
        (reference quantization.py:20-55; same parameters, same torch-RNG consumption)."""
        super().__init__()
        self.dim = dim
        self.codebook_size = codebook_size
        self.num_codebooks = num_codebooks
        assert _is_power_of_two(codebook_size)
        assert _is_power_of_two(num_codebooks)

        self.to_logits = nn.Linear(dim, codebook_size * num_codebooks)
        self.centers = nn.Parameter(
            self.to_logits.weight.detach().clone().reshape(num_codebooks, codebook_size, dim))
        self.logits_scale = nn.Parameter(torch.zeros(()))
        self.centers_scale = nn.Parameter(torch.zeros(()))
        self.scale_speed = 10.0

        id_bytes = binascii.b2a_hex(os.urandom(4))
        self.id_str = id_bytes.decode("utf-8")
        self.register_buffer("id_buf", torch.tensor(list(id_bytes), dtype=torch.uint8))

        self._prep_key = None
        self._prep_blob: Optional[Tensor] = None
        self._ws: Optional[Tensor] = None

    # ------------------------------------------------------------------ bookkeeping (reference :57-79)
    def load_state_dict(self, *args, **kwargs):
        ret = super().load_state_dict(*args, **kwargs)
        self.id_str = bytes(self.id_buf.tolist()).decode("utf-8")
        return ret

    def get_id(self) -> str:
        return self.id_str

    def show_init_invocation(self) -> str:
        return (f"quantization.Quantizer(dim={self.dim}, codebook_size={self.codebook_size}, "
                f"num_codebooks={self.num_codebooks})")

    def get_data_mean(self) -> Tensor:
        return self.get_centers().mean(dim=1).sum(dim=0).detach()

    def get_centers(self) -> Tensor:
        scale = (self.centers_scale * self.scale_speed).exp()
        return scale * self.centers

    # ------------------------------------------------------------------ device-side prepared state
    def _params(self):
        return (self.centers, self.centers_scale, self.to_logits.weight, self.to_logits.bias, self.logits_scale)

    def invalidate_prepared(self) -> None:
        """Forget the prepared state.  Needed only after changing parameter VALUES in a way PyTorch's version counters
        do not see (a fused optimiser step outside compute_loss, writes through raw pointers)."""
        self._prep_key = None

    def _prepared(self) -> Tensor:
        """The prepared blob (scaled centers, Gram table, operand splits), rebuilt when any parameter changes (as seen
        through data_ptr / _version; compute_loss additionally drops it around every training step)."""
        if self.codebook_size > MAX_CODEBOOK_SIZE or self.num_codebooks > MAX_NUM_CODEBOOKS:
            # e.g. get_product_quantizer() of a codebook_size-256 quantizer: constructible (as in the reference), but
            # there is no kernel for it and no PyTorch fallback by design
            raise NotImplementedError(
                f"codebook_size={self.codebook_size} / num_codebooks={self.num_codebooks}: the CUDA search kernels "
                f"cover codebook_size <= {MAX_CODEBOOK_SIZE} and num_codebooks <= {MAX_NUM_CODEBOOKS} "
                "(every shape QuantizerTrainer produces); there is no CPU / PyTorch fallback")
        ps = self._params()
        dev = self.centers.device
        _lib.require_cuda(self.centers, "Quantizer parameters")
        key = (dev, float(self.scale_speed)) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._prep_key or self._prep_blob is None:
            for p in ps:
                if p.dtype != torch.float32:
                    raise RuntimeError("Quantizer parameters must be float32 (the reference's .half()/.bfloat16() "
                                       "quantizer is a different numerical function and is not supported)")
            L = _lib.lib()
            N, K, D = self.num_codebooks, self.codebook_size, self.dim
            nbytes = L.mcq_prepared_bytes(N, K, D)
            if nbytes == 0:
                _lib.check(-1, "mcq_prepared_bytes")
            if self._prep_blob is None or self._prep_blob.numel() < nbytes or self._prep_blob.device != dev:
                self._prep_blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            c, cs, w, b, ls = (p.detach().contiguous() for p in ps)
            with torch.cuda.device(dev):
                rc = L.mcq_prepare(c.data_ptr(), cs.data_ptr(), w.data_ptr(), b.data_ptr(), ls.data_ptr(),
                                   float(self.scale_speed), N, K, D, self._prep_blob.data_ptr(), nbytes,
                                   _lib.stream_ptr(dev))
            _lib.check(rc, "mcq_prepare")
            self._prep_key = key
        return self._prep_blob

    def _workspace(self, num_frames: int) -> Tensor:
        L = _lib.lib()
        need = L.mcq_workspace_bytes(max(int(num_frames), 1), self.dim, self.num_codebooks, self.codebook_size)
        dev = self.centers.device
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        return self._ws

    def _check_x(self, x: Tensor) -> Tensor:
        _lib.require_cuda(x, "x")
        if x.device != self.centers.device:
            raise RuntimeError(f"x is on {x.device} but the quantizer is on {self.centers.device}")
        _lib.x_dtype_code(x)
        return x.contiguous()

    # ------------------------------------------------------------------ hot path
    def _encode_2d(self, x: Tensor, iters: int, codes_dtype: int) -> Tensor:
        x = self._check_x(x)
        B = x.shape[0]
        N, K, D = self.num_codebooks, self.codebook_size, self.dim
        L = _lib.lib()
        blob = self._prepared()
        if codes_dtype == _lib.U8:
            out = torch.empty(B, L.mcq_packed_cols(N, K), dtype=torch.uint8, device=x.device)
        else:
            out = torch.empty(B, N, dtype=torch.int64, device=x.device)
        if B == 0:
            return out
        ws = self._workspace(B)
        with torch.cuda.device(x.device):
            rc = L.mcq_encode(x.data_ptr(), _lib.x_dtype_code(x), B, D, N, K, blob.data_ptr(), int(iters),
                              out.data_ptr(), codes_dtype, ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "mcq_encode")
        return out

    def encode(self, x: Tensor, refine_indexes_iters: int = 5, as_bytes: bool = True) -> Tensor:
        """Reference quantization.py:244-275.  x (*, dim) -> uint8 (*, N_packed) if as_bytes else int64 (*, N)."""
        x2 = x.reshape(-1, self.dim)
        codes = self._encode_2d(x2, refine_indexes_iters, _lib.U8 if as_bytes else _lib.I64)
        return codes.reshape(*x.shape[:-1], codes.shape[-1])

    def encode_host(self, x: Tensor, refine_indexes_iters: int = 5, as_bytes: bool = True,
                    out: Optional[Tensor] = None, library_buffers: bool = False) -> Tensor:
        """Extension: `x` (*, dim) lives in HOST memory (ideally pinned); returns host codes.  Chunks are streamed
        through the device with the copies overlapping the kernels.  Default: mcq_encode_host_ws -- the staging
        buffer is a PyTorch tensor cached on the quantizer, the work is ordered on the current stream, which is
        synchronised before returning.  library_buffers=True: mcq_encode_host (the library's own cached device
        buffers; what a host without CUDA plumbing of its own binds)."""
        if x.is_cuda:
            raise RuntimeError("encode_host expects a host tensor; use encode() for device tensors")
        x2 = x.reshape(-1, self.dim).contiguous()
        B = x2.shape[0]
        N, K, D = self.num_codebooks, self.codebook_size, self.dim
        L = _lib.lib()
        blob = self._prepared()
        dev = self.centers.device
        cols = L.mcq_packed_cols(N, K) if as_bytes else N
        dt = torch.uint8 if as_bytes else torch.int64
        cdt = _lib.U8 if as_bytes else _lib.I64
        if out is None:
            out = torch.empty(B, cols, dtype=dt, pin_memory=True)
        assert out.shape == (B, cols) and out.dtype == dt and out.is_contiguous() and not out.is_cuda
        if B == 0:
            return out.reshape(*x.shape[:-1], cols)
        if library_buffers:
            torch.cuda.current_stream(dev).synchronize()  # the blob may just have been rebuilt on this stream
            rc = L.mcq_encode_host(x2.data_ptr(), _lib.x_dtype_code(x2), B, D, N, K, blob.data_ptr(),
                                   int(refine_indexes_iters), out.data_ptr(), cdt,
                                   dev.index if dev.index is not None else torch.cuda.current_device())
            _lib.check(rc, "mcq_encode_host")
            return out.reshape(*x.shape[:-1], cols)
        need = int(L.mcq_encode_host_ws_bytes(B, D, N, K, _lib.x_dtype_code(x2), cdt))
        st = getattr(self, "_host_staging", None)
        if st is None or st.device != dev or st.numel() < need:
            st = torch.empty(need, dtype=torch.uint8, device=dev)
            self._host_staging = st
        with torch.cuda.device(dev):
            rc = L.mcq_encode_host_ws(x2.data_ptr(), _lib.x_dtype_code(x2), B, D, N, K, blob.data_ptr(),
                                      int(refine_indexes_iters), out.data_ptr(), cdt, st.data_ptr(), st.numel(),
                                      _lib.stream_ptr(dev))
        _lib.check(rc, "mcq_encode_host_ws")
        torch.cuda.current_stream(dev).synchronize()  # the codes are in host memory when the stream has drained
        return out.reshape(*x.shape[:-1], cols)

    def _compute_indexes(self, x: Tensor, refine_indexes_iters: int = 3) -> Tensor:
        """Reference quantization.py:281-305: classifier arg-max, then `refine_indexes_iters` refinement passes."""
        assert x.ndim == 2 and x.shape[1] == self.dim
        return self._encode_2d(x, refine_indexes_iters, _lib.I64)

    def _refine_indexes(self, x: Tensor, indexes: Tensor) -> Tensor:
        """One pass of the reference's hierarchical pairwise search (quantization.py:308-547)."""
        return self._refine(x, indexes, 1)

    def _refine(self, x: Tensor, indexes: Tensor, iters: int) -> Tensor:
        x = self._check_x(x)
        assert x.ndim == 2 and x.shape[1] == self.dim
        B = x.shape[0]
        N, K, D = self.num_codebooks, self.codebook_size, self.dim
        assert indexes.shape == (B, N)
        idx = indexes.to(dtype=torch.int64, device=x.device).contiguous()
        out = torch.empty_like(idx)
        if B == 0:
            return out
        L = _lib.lib()
        blob = self._prepared()
        ws = self._workspace(B)
        with torch.cuda.device(x.device):
            rc = L.mcq_refine(x.data_ptr(), _lib.x_dtype_code(x), B, D, N, K, blob.data_ptr(), int(iters),
                              idx.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device))
        _lib.check(rc, "mcq_refine")
        return out

    def decode(self, indexes: Tensor) -> Tensor:
        """Reference quantization.py:117-148.  indexes (*, n) with n == num_codebooks or a packed column count
        -> (*, dim) float32: the sum of the selected scaled centers."""
        orig_shape = indexes.shape
        idx = indexes.reshape(-1, indexes.shape[-1])
        _lib.require_cuda(idx, "indexes")
        if idx.dtype not in (torch.uint8, torch.int32, torch.int64):
            idx = idx.to(torch.int64)
        idx = idx.contiguous()
        B, ncols = idx.shape
        N, K, D = self.num_codebooks, self.codebook_size, self.dim
        if ncols != N:
            r = N // max(ncols, 1)
            assert ncols > 0 and N % ncols == 0 and r in (2, 4, 8, 16)  # reference :566
        if _CHECK_INDEXES and idx.dtype != torch.uint8 and B > 0:
            # debugging aid only (MCQ_CHECK_INDEXES=1): a blocking host round trip on an HBM-bound path, and it
            # would make decode uncapturable in a CUDA graph.  The kernel clamps out-of-range codes to entry 0.
            hi = K ** (N // ncols)
            if bool(((idx < 0) | (idx >= hi)).any()):
                raise IndexError(f"decode: indexes out of range [0, {hi})")
        needs_grad = torch.is_grad_enabled() and (self.centers.requires_grad or self.centers_scale.requires_grad)
        if needs_grad and ncols == N:
            out = _DecodeFn.apply(self.get_centers(), idx.to(torch.int64))
        else:
            L = _lib.lib()
            blob = self._prepared()
            if (idx.dtype != torch.uint8 and ncols == N and N in (4, 8) and D % 32 == 0 and B >= _SLAB_MIN_FRAMES):
                # large batches of int32 / int64 indexes: one narrowing pass (out-of-range indexes become entry 0, as
                # the kernels treat them) buys the shared-memory slab kernel, which takes byte codes
                idx = torch.where((idx < 0) | (idx >= K), torch.zeros_like(idx), idx).to(torch.uint8)
            out = torch.empty(B, D, dtype=torch.float32, device=idx.device)
            if B > 0:
                with torch.cuda.device(idx.device):
                    rc = L.mcq_decode(idx.data_ptr(), _lib.idx_dtype_code(idx), B, ncols, N, K, D, blob.data_ptr(),
                                      out.data_ptr(), _lib.F32, _lib.stream_ptr(idx.device))
                _lib.check(rc, "mcq_decode")
        return out.reshape(*orig_shape[:-1], D)

    def _maybe_separate_indexes(self, indexes: Tensor) -> Tensor:
        """Reference quantization.py:551-573 (kept for API completeness; decode unpacks inside its kernel)."""
        n = indexes.shape[1]
        if n == self.num_codebooks:
            return indexes
        r = self.num_codebooks // n
        assert r in (2, 4, 8, 16)
        K = self.codebook_size
        powers = K ** torch.arange(r, device=indexes.device)
        return ((indexes.unsqueeze(2).to(torch.int64) // powers) % K).reshape(indexes.shape[0], self.num_codebooks)

    # ------------------------------------------------------------------ training-side (reference :184-242)
    def compute_loss(self, x: Tensor, refine_indexes_iters: int = 0):
        """Returns (rel_reconstruction_loss, logprob_loss, logits_entropy_loss, index_entropy_loss) exactly as the
        reference defines them (quantization.py:184-242); only the index search and the decode gather run in the
        CUDA library (the search carries no gradient in the reference either)."""
        x = x.reshape(-1, self.dim)
        # A training step: the parameters have probably just been changed by an optimiser, and fused optimisers
        # (torch.optim.Adam(fused=True), which QuantizerTrainer uses) do NOT bump the tensors' version counters the
        # prepared-state cache is keyed on.  So the cache is dropped on entry (this step prepares from the current
        # values, once) and again on exit (whatever runs after the optimiser step -- the trainer's no_grad
        # diagnostics, an encode() -- must not see this step's tables).
        training_call = torch.is_grad_enabled() and any(p.requires_grad for p in self._params())
        if training_call:
            self._prep_key = None
        try:
            return self._compute_loss(x, refine_indexes_iters)
        finally:
            if training_call:
                self._prep_key = None

    def _compute_loss(self, x: Tensor, refine_indexes_iters: int):
        with torch.no_grad():
            indexes = self._compute_indexes(x, refine_indexes_iters)
        if self.dim <= 1024 and x.shape[0] > 0:
            # :209-216 fused: both sums from one pass over the frames, gradient w.r.t. the scaled centers from another
            sums = _ReconLossFn.apply(self.get_centers(), indexes, self._check_x(x), self.get_data_mean())
            rel_reconstruction_loss = sums[0] / (sums[1] + 1.0e-20)
        else:
            xf = x.float() if x.dtype != torch.float32 else x
            needs_grad = torch.is_grad_enabled() and (self.centers.requires_grad or self.centers_scale.requires_grad)
            if needs_grad:
                x_approx = _DecodeFn.apply(self.get_centers(), indexes)
            else:
                x_approx = self.decode(indexes)
            tot_error = x_approx - xf
            rel_reconstruction_loss = (tot_error ** 2).sum() / (((xf - self.get_data_mean()) ** 2).sum() + 1.0e-20)

        N, K = self.num_codebooks, self.codebook_size
        B = x.shape[0]
        if B == 0 or K < _FUSED_LOSS_MIN_K:
            # (K = 16, trainer phase 1, also takes the fused kernels: 1.49 vs 1.89 ms of kernels per step at 65,536 frames)
            xf = x.float() if x.dtype != torch.float32 else x
            return self._compute_loss_tail_torch(xf, indexes, rel_reconstruction_loss)
        # logprob / entropy terms (reference :218-240) from two fused reductions over the logits
        xc = self._check_x(x)
        logprob_sum, prob_sum = _ClassLossFn.apply(self, xc, indexes, self.to_logits.weight, self.to_logits.bias,
                                                   self.logits_scale)
        logprob_loss = -(logprob_sum / (B * N))

        # histogram of the chosen entries (:227-231): shared-memory integer counting in the library (exact, no host
        # synchronisation: the trainer captures this function in a CUDA graph); a scatter_add of ones costs 92 us of
        # contended atomics at 65,536 frames x 8 codebooks of 16 entries
        if N * K <= 12288:
            counts = _lib.index_counts(indexes.contiguous(), N, K)
        else:
            flat = (indexes + torch.arange(N, device=indexes.device) * K).reshape(-1)
            counts = torch.zeros(N * K, dtype=torch.float32, device=indexes.device).scatter_add_(
                0, flat, torch.ones(1, dtype=torch.float32, device=indexes.device).expand(flat.numel())).reshape(N, K)
        avg_counts = counts / B + 1.0e-20
        index_entropy = -(avg_counts * avg_counts.log()).sum(dim=1).mean()

        probs = prob_sum / B + 1.0e-20
        logits_entropy = -(probs * probs.log()).sum(dim=1).mean()
        ref_entropy = math.log(K)
        logits_entropy_loss = (ref_entropy - logits_entropy) / ref_entropy
        index_entropy_loss = (ref_entropy - index_entropy) / ref_entropy
        return rel_reconstruction_loss, logprob_loss, logits_entropy_loss, index_entropy_loss

    def _compute_loss_tail_torch(self, xf: Tensor, indexes: Tensor, rel_reconstruction_loss: Tensor):
        """The reference's own formulation of the classifier-side losses (quantization.py:218-240), used for shapes
        the fused kernels do not pay off for (empty batches, a single codebook of fewer than 16 entries).  Plain
        PyTorch on device tensors."""
        N, K = self.num_codebooks, self.codebook_size
        logits = self._logits(xf).reshape(-1, N, K).log_softmax(dim=2)
        chosen = torch.gather(logits, dim=2, index=indexes.unsqueeze(2))
        logprob_loss = -chosen.mean()
        B = xf.shape[0]
        counts = torch.zeros(B, N, K, device=xf.device)
        counts.scatter_(dim=2, index=indexes.unsqueeze(2), value=1.0)
        avg_counts = counts.mean(dim=0) + 1.0e-20
        index_entropy = -(avg_counts * avg_counts.log()).sum(dim=1).mean()
        probs = logits.exp().mean(dim=0) + 1.0e-20
        logits_entropy = -(probs * probs.log()).sum(dim=1).mean()
        ref_entropy = math.log(K)
        return (rel_reconstruction_loss, logprob_loss, (ref_entropy - logits_entropy) / ref_entropy,
                (ref_entropy - index_entropy) / ref_entropy)

    def _logits(self, x: Tensor) -> Tensor:
        x = (self.logits_scale * self.scale_speed).exp() * x
        return self.to_logits(x)

    def compute_codebook_correlations(self) -> Tensor:
        """Diagnostic of reference quantization.py:150-181: normalised tr(S_i S_j) of the per-codebook
        (mean-removed) second-moment matrices."""
        c = self.get_centers().detach()
        c = c - c.mean(dim=1, keepdim=True)
        var = torch.matmul(c.transpose(1, 2), c).reshape(self.num_codebooks, self.dim * self.dim)
        cross = torch.matmul(var, var.t())
        norm = cross.diag() ** -0.5
        return cross * (norm.unsqueeze(0) * norm.unsqueeze(1))

    def get_product_quantizer(self) -> "Quantizer":
        """Reference quantization.py:81-112: K -> K*K, N -> N/2, entry k1*K + k2 of output codebook c is the sum of
        entry k1 of input codebook 2c and entry k2 of input codebook 2c+1 (weights, biases and centers alike).
        Vectorised; the sums are the same fp32 additions the reference's triple loop performs."""
        N, K, D = self.num_codebooks, self.codebook_size, self.dim
        ans = Quantizer(D, K * K, N // 2).to(self.centers.device)
        ans.apply_mask = False
        with torch.no_grad():
            ans.logits_scale.fill_(self.logits_scale.item())
            ans.centers_scale.fill_(self.centers_scale.item())
            ans.scale_speed = self.scale_speed
            W = self.to_logits.weight.reshape(N, K, D)
            b = self.to_logits.bias.reshape(N, K)
            C = self.centers
            ans.to_logits.weight.copy_((W[0::2, :, None, :] + W[1::2, None, :, :]).reshape(N // 2 * K * K, D))
            ans.to_logits.bias.copy_((b[0::2, :, None] + b[1::2, None, :]).reshape(N // 2 * K * K))
            ans.centers.copy_((C[0::2, :, None, :] + C[1::2, None, :, :]).reshape(N // 2, K * K, D))
        return ans
