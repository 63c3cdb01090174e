"""JointCodebookLoss -- host-side mirror of the reference's `quantization.prediction.JointCodebookLoss`
(prediction.py:86-197; functional form `joint_codebook_loss` :9-82): predicts the codes `Quantizer.encode` produced
from a predictor vector, each codebook regressed on the previous ones.

Same constructor arguments, parameter names (`linear1`, `codebook_embedding`, `linear2_weight`, `linear2b_weight`,
`linear2_bias` -- state_dicts interchange with the reference) and return value.  What differs is where the work runs:

  * the embedding gather + concat + cumsum + ReLU (:47-68) is one kernel of libmcq.so (`mcq_jcl_hidden_forward`) that
    writes the (N, B, hidden) operand of the per-codebook product directly; the reference materialises four (B, N, hidden)
    tensors for it;
  * the cross entropy (:79-82) is one kernel (`mcq_jcl_cross_entropy`) that also leaves softmax - onehot in place of the
    logits, so the backward pass starts from it with no log-softmax graph;
  * the backward of the gather/cumsum/ReLU stage is one kernel (`mcq_jcl_hidden_backward`);
  * the dense products -- fp32 SGEMMs in the reference -- are fp32-faithful tcgen05 products (fp16x2 operand split, fp32
    accumulation): `mcq_gemm_nt` for the forward and input-gradient products, split-K `mcq_gemm_tn` for the weight
    gradients (reductions over all frames).  Shapes whose widths are not multiples of 64 use the library GEMM.

`checkpoint=True` keeps the reference's meaning (prediction.py:113-115: recompute in backward, store only the inputs).
The codes may be the uint8 tensor `Quantizer.encode` returns (no int64 copy is made), int32 or int64 (negative =
padding, ignored when equal to `ignore_index`).  There is no CPU fallback.
"""
import torch
from torch import Tensor, nn

from . import _lib


def _use_tc(B: int, P: int, H: int, K: int) -> bool:
    """The tcgen05 products take output widths that are multiples of 64 (mcq_gemm_nt); tiny or odd shapes go to the
    library GEMM (identical mathematics, fp32)."""
    return B >= 256 and P % 64 == 0 and H % 64 == 0 and K % 64 == 0


def _stages_forward(pred: Tensor, codes: Tensor, w1: Tensor, b1, emb: Tensor, w2: Tensor, w2b: Tensor, bias2: Tensor,
                    ignore_index: int, want_grad: bool):
    """Returns (row_loss (B, N), sums (2,), act (N, B, H), dlogits (B, N*K) or None)."""
    L = _lib.lib()
    B, P = pred.shape
    N, K, H = w2.shape
    dev = pred.device
    stream = _lib.stream_ptr(dev)
    tc = _use_tc(B, P, H, K)
    if tc:  # prediction.py:56 as an fp32-faithful tcgen05 product
        hidden = _lib.gemm_nt(pred, w1)
        if b1 is not None:
            hidden.add_(b1)
    else:
        hidden = torch.addmm(b1, pred, w1.t()) if b1 is not None else pred.mm(w1.t())  # (B, H)
    act = torch.empty(N, B, H, dtype=torch.float32, device=dev)
    scale = 0.5 * ((H / N) ** 0.5)  # prediction.py:51
    with torch.cuda.device(dev):
        _lib.check(L.mcq_jcl_hidden_forward(hidden.data_ptr(), codes.data_ptr(), _lib.idx_dtype_code(codes), B, N, K, H,
                                            emb.data_ptr(), scale, act.data_ptr(), stream), "mcq_jcl_hidden_forward")
    # logits (B, N, K): predictor part as ONE product against all codebooks (:74-76), hidden part per codebook (:70-72)
    if tc:
        logits = _lib.gemm_nt(pred, w2b.reshape(N * K, P))
        for n in range(N):
            _lib.gemm_nt(act[n], w2[n], out=logits[:, n * K:(n + 1) * K], accumulate=True)
    else:
        logits = pred.mm(w2b.reshape(N * K, P).t())
        lv = logits.view(B, N, K).transpose(0, 1)  # (N, B, K) view
        lv.baddbmm_(act, w2.transpose(1, 2))
    row_loss = torch.empty(B, N, dtype=torch.float32, device=dev)
    sums = torch.empty(2, dtype=torch.float32, device=dev)
    partials = torch.empty(L.mcq_jcl_partials(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.mcq_jcl_cross_entropy(logits.data_ptr(), bias2.data_ptr(), codes.data_ptr(),
                                           _lib.idx_dtype_code(codes), B, N, K, int(ignore_index), int(want_grad),
                                           row_loss.data_ptr(), sums.data_ptr(), partials.data_ptr(), stream),
                   "mcq_jcl_cross_entropy")
    return row_loss, sums, act, (logits if want_grad else None)


class _JointCodebookLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, codes, w1, b1, emb, w2, w2b, bias2, ignore_index, reduction, checkpoint):
        need = any(ctx.needs_input_grad)
        row_loss, sums, act, dlogits = _stages_forward(pred, codes, w1, b1, emb, w2, w2b, bias2, ignore_index,
                                                       need and not checkpoint)
        ctx.cfg = (ignore_index, reduction, checkpoint, b1 is not None)
        if checkpoint or not need:
            ctx.save_for_backward(pred, codes, w1, b1, emb, w2, w2b, bias2, sums)
        else:
            ctx.save_for_backward(pred, codes, w1, b1, emb, w2, w2b, bias2, sums, act, dlogits)
        if reduction == "sum":
            return sums[0].clone()
        if reduction == "mean":
            return sums[0] / sums[1]
        return row_loss.reshape(-1)  # 'none': one loss per (frame, codebook), like cross_entropy on the flattened rows

    @staticmethod
    def backward(ctx, g):
        ignore_index, reduction, checkpoint, has_b1 = ctx.cfg
        saved = ctx.saved_tensors
        pred, codes, w1, b1, emb, w2, w2b, bias2, sums = saved[:9]
        if checkpoint:
            _, _, act, dl = _stages_forward(pred, codes, w1, b1, emb, w2, w2b, bias2, ignore_index, True)
        else:
            act, dl = saved[9], saved[10]
        L = _lib.lib()
        B, P = pred.shape
        N, K, H = w2.shape
        dev = pred.device
        # dl (B, N*K) holds d(sum of row losses) / d logits; fold the upstream gradient in
        if reduction == "none":
            dl = dl.view(B, N, K) * g.reshape(B, N, 1).to(torch.float32)
            dl = dl.view(B, N * K)
            gs = None
        else:
            gs = g.to(torch.float32) if reduction == "sum" else g.to(torch.float32) / sums[1]
        dlv = dl.view(B, N, K).transpose(0, 1)  # (N, B, K) view
        tc = _use_tc(B, P, H, K)
        if tc:
            grad_act = torch.empty(N, B, H, dtype=torch.float32, device=dev)
            w2t = w2.transpose(1, 2).contiguous()  # (N, H, K)
            for n in range(N):
                _lib.gemm_nt(dl[:, n * K:(n + 1) * K], w2t[n], out=grad_act[n])
        else:
            grad_act = torch.bmm(dlv, w2)  # (N, B, H)
        grad_hidden = torch.empty(B, H, dtype=torch.float32, device=dev)
        grad_emb = torch.zeros_like(emb)
        scale = 0.5 * ((H / N) ** 0.5)
        with torch.cuda.device(dev):
            _lib.check(L.mcq_jcl_hidden_backward(grad_act.data_ptr(), act.data_ptr(), codes.data_ptr(),
                                                 _lib.idx_dtype_code(codes), B, N, K, H, scale, grad_hidden.data_ptr(),
                                                 grad_emb.data_ptr(), _lib.stream_ptr(dev)), "mcq_jcl_hidden_backward")
        # weight gradients: reductions over the frames -> split-K tcgen05 products (mcq_gemm_tn)
        grad_w2 = torch.stack([_lib.gemm_tn(dl[:, n * K:(n + 1) * K], act[n]) for n in range(N)])  # (N, K, H)
        grad_w2b = _lib.gemm_tn(dl, pred).view(N, K, P)
        grad_bias2 = dl.sum(dim=0).view(N, K)
        if tc:
            grad_pred = _lib.gemm_nt(dl, w2b.reshape(N * K, P).t().contiguous())
            _lib.gemm_nt(grad_hidden, w1.t().contiguous(), out=grad_pred, accumulate=True)
        else:
            grad_pred = dl.mm(w2b.reshape(N * K, P))
            grad_pred.addmm_(grad_hidden, w1)
        grad_w1 = _lib.gemm_tn(grad_hidden, pred)
        grad_b1 = grad_hidden.sum(dim=0) if has_b1 else None
        if gs is not None:  # everything above is linear in dl: scale the (small) results instead of the (B, N*K) tensor
            grad_pred, grad_w1, grad_emb, grad_w2, grad_w2b, grad_bias2 = (
                t * gs for t in (grad_pred, grad_w1, grad_emb, grad_w2, grad_w2b, grad_bias2))
            if grad_b1 is not None:
                grad_b1 = grad_b1 * gs
        return grad_pred, None, grad_w1, grad_b1, grad_emb, grad_w2, grad_w2b, grad_bias2, None, None, None


def joint_codebook_loss(predictor: Tensor, codebook_indexes: Tensor, linear1_weight: Tensor, linear1_bias,
                        codebook_embedding_weight: Tensor, linear2_weight: Tensor, linear2b_weight: Tensor,
                        linear2_bias: Tensor, ignore_index: int, reduction: str, checkpoint: bool = False) -> Tensor:
    """Functional form with the reference's argument order (prediction.py:9-18)."""
    if reduction not in ("sum", "mean", "none"):
        raise ValueError(f"{reduction} is not a valid value for reduction")
    num_codebooks = codebook_indexes.shape[-1]
    assert list(predictor.shape[:-1]) == list(codebook_indexes.shape[:-1])  # prediction.py:40
    assert linear2_weight.shape[0] == num_codebooks and num_codebooks > 1
    if not predictor.is_cuda:
        raise RuntimeError("quantization_b200 has no CPU path: predictor must be a CUDA tensor")
    out_dtype = predictor.dtype
    pred = predictor.reshape(-1, predictor.shape[-1]).to(torch.float32).contiguous()
    codes = codebook_indexes.reshape(-1, num_codebooks)
    if codes.dtype not in (torch.uint8, torch.int32, torch.int64):
        codes = codes.to(torch.int64)  # prediction.py:39
    codes = codes.contiguous()
    if pred.shape[0] == 0:
        z = (pred.sum() + linear2_bias.sum() * 0.0)
        return z if reduction != "none" else pred.new_zeros(0)
    loss = _JointCodebookLossFn.apply(pred, codes, linear1_weight, linear1_bias, codebook_embedding_weight,
                                      linear2_weight, linear2b_weight, linear2_bias, ignore_index, reduction, checkpoint)
    return loss if out_dtype == torch.float32 else loss.to(out_dtype)


class JointCodebookLoss(nn.Module):
    """Drop-in for the reference module (prediction.py:86-197); see the module docstring for what runs where.

    Args (identical to the reference, :118-125):
        predictor_channels: number of features of the predictor.
        num_codebooks: number of codebooks predicted (> 1), normally the Quantizer's.
        hidden_channels: hidden dimension of the one-hidden-layer network (a multiple of 4 here).
        codebook_size: entries per codebook (<= 1024 here).
        reduction: 'sum' (default), 'mean' or 'none'.
        ignore_index: value of codebook_indexes that marks padding.
        checkpoint: recompute the forward in backward instead of storing activations.
    """

    def __init__(self, predictor_channels: int, num_codebooks: int, hidden_channels: int = 512, codebook_size: int = 256,
                 reduction: str = "sum", ignore_index: int = -100, checkpoint: bool = True):
        super().__init__()
        assert num_codebooks > 1  # prediction.py:128
        assert hidden_channels % 4 == 0, "hidden_channels must be a multiple of 4 (128-bit accesses)"
        self.num_codebooks = num_codebooks
        self.codebook_size = codebook_size
        self.hidden_channels = hidden_channels
        self.ignore_index = ignore_index
        self.reduction = reduction
        self.checkpoint = checkpoint
        # same construction order and initial distributions as the reference (:136-152), so equal seeds give equal
        # parameters
        self.linear1 = nn.Linear(predictor_channels, hidden_channels)
        self.codebook_embedding = nn.Embedding(
            (num_codebooks - 1) * codebook_size, hidden_channels,
            _weight=torch.randn((num_codebooks - 1) * codebook_size, hidden_channels) * (hidden_channels ** -0.5))
        self.linear2_weight = nn.Parameter(torch.randn(num_codebooks, codebook_size, hidden_channels)
                                           * (hidden_channels ** -0.5))
        self.linear2b_weight = nn.Parameter(torch.randn(num_codebooks, codebook_size, predictor_channels)
                                            * (predictor_channels ** -0.5))
        self.linear2_bias = nn.Parameter(torch.zeros(num_codebooks, codebook_size))

    def forward(self, predictor: Tensor, codebook_indexes: Tensor) -> Tensor:
        """predictor (*, predictor_channels), codebook_indexes (*, num_codebooks) -> loss (prediction.py:155-197)."""
        return joint_codebook_loss(predictor, codebook_indexes, self.linear1.weight, self.linear1.bias,
                                   self.codebook_embedding.weight, self.linear2_weight, self.linear2b_weight,
                                   self.linear2_bias, self.ignore_index, self.reduction, self.checkpoint)
