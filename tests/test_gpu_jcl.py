"""GPU parity of JointCodebookLoss (quantization_b200/prediction.py + csrc/jcl.cu, C ABI mcq_jcl_*) against
(i) the fixtures the reference itself produced (tests/golden/golden_jcl.npz), (ii) the numpy oracle
(oracle/jcl_oracle.py) -- the hidden stage bit for bit, losses and gradients within the fp32 tolerance written below --
and (iii) at a realistic size, a plain PyTorch fp32 evaluation of the same function on the same device."""
import numpy as np
import pytest
import torch

import helpers
from oracle import jcl_oracle as jo
from quantization_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
_TORCH = {"int64": torch.int64, "int32": torch.int32, "uint8": torch.uint8}


def _module(m, par, checkpoint):
    from quantization_b200 import JointCodebookLoss
    mod = JointCodebookLoss(m["P"], m["N"], hidden_channels=m["H"], codebook_size=m["K"], reduction=m["reduction"],
                            checkpoint=checkpoint)
    mod.load_state_dict({k: torch.from_numpy(v) for k, v in par.items()})  # the reference's own state_dict keys
    return mod.to(DEV)


@pytest.mark.parametrize("checkpoint", [False, True])
@pytest.mark.parametrize("name", helpers.jcl_case_names())
def test_jcl_matches_reference_golden(name, checkpoint):
    g, meta = helpers.jcl_golden()
    m, pred, codes, par, grads = helpers.jcl_case(g, meta, name)
    mod = _module(m, par, checkpoint)
    shape = m["shape"]
    x = torch.from_numpy(pred).reshape(*shape, m["P"]).to(DEV).requires_grad_(True)
    c = torch.from_numpy(codes).reshape(*shape, m["N"]).to(_TORCH[m["codes_dtype"]]).to(DEV)
    loss = mod(x, c)
    ref = torch.from_numpy(g[name + "/loss"])
    assert tuple(loss.shape) == tuple(ref.shape)
    assert torch.allclose(loss.cpu(), ref, rtol=1e-5, atol=1e-5), (loss, ref)
    up = torch.from_numpy(g[name + "/upstream"]).to(DEV)
    (loss * up).sum().backward()
    got = dict(mod.named_parameters())
    for pn, ref_g in grads.items():
        gg = got[pn].grad.cpu().numpy()
        assert np.abs(gg - ref_g).max() <= 1e-4 * (np.abs(ref_g).max() + 1e-30), pn  # fp32 products, different order
    ref_gx = g[name + "/g_pred"]
    assert np.abs(x.grad.cpu().numpy() - ref_gx).max() <= 1e-4 * (np.abs(ref_gx).max() + 1e-30)


@pytest.mark.parametrize("N,K,H,B,dt", [(8, 256, 512, 1000, torch.uint8), (4, 16, 36, 257, torch.int64),
                                        (40, 4, 8, 33, torch.int32), (2, 1000, 128, 50, torch.int64)])
def test_jcl_hidden_stage_bit_exact(N, K, H, B, dt):
    """mcq_jcl_hidden_forward against the fp32 restatement of prediction.py:41-68: identical bits."""
    gen = torch.Generator().manual_seed(5)
    hidden = torch.randn(B, H, generator=gen)
    emb = torch.randn((N - 1) * K, H, generator=gen) * H ** -0.5
    codes = torch.randint(0, K, (B, N), generator=gen)
    if dt != torch.uint8:
        codes[::7] = -100
    L = _lib.lib()
    act = torch.empty(N, B, H, device=DEV)
    hd, ed, cd = hidden.to(DEV), emb.to(DEV), codes.to(dt).to(DEV)
    _lib.check(L.mcq_jcl_hidden_forward(hd.data_ptr(), cd.data_ptr(), _lib.idx_dtype_code(cd), B, N, K, H, ed.data_ptr(),
                                        jo.embedding_scale(H, N), act.data_ptr(), _lib.stream_ptr(DEV)), "hidden")
    ref = jo.hidden_stage(hidden.numpy(), codes.numpy(), emb.numpy(), K)
    assert np.array_equal(act.cpu().numpy(), ref)


def test_jcl_cross_entropy_stage():
    """mcq_jcl_cross_entropy against an fp64 evaluation: row losses, sums, count, in-place gradient; reproducible."""
    B, N, K = 3000, 8, 256
    gen = torch.Generator().manual_seed(6)
    logits = 3 * torch.randn(B, N, K, generator=gen)
    bias = torch.randn(N, K, generator=gen)
    codes = torch.randint(0, K, (B, N), generator=gen)
    codes[::5] = -100
    L = _lib.lib()
    outs = []
    for _ in range(2):
        ld, bd, cd = logits.to(DEV), bias.to(DEV), codes.to(DEV)
        row = torch.empty(B, N, device=DEV)
        sums = torch.empty(2, device=DEV)
        part = torch.empty(L.mcq_jcl_partials(), device=DEV)
        _lib.check(L.mcq_jcl_cross_entropy(ld.data_ptr(), bd.data_ptr(), cd.data_ptr(), _lib.idx_dtype_code(cd), B, N, K,
                                           -100, 1, row.data_ptr(), sums.data_ptr(), part.data_ptr(),
                                           _lib.stream_ptr(DEV)), "ce")
        outs.append((row.cpu(), sums.cpu(), ld.cpu()))
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))  # fixed-order sums: bitwise reproducible
    row, sums, dl = outs[0]
    z = (logits + bias).double()
    logp = z.log_softmax(2)
    valid = codes != -100
    ref_row = -(logp.gather(2, codes.clamp(min=0).unsqueeze(2)).squeeze(2)) * valid
    assert torch.allclose(row.double(), ref_row, rtol=1e-5, atol=1e-5)
    assert abs(sums[0].item() - ref_row.sum().item()) <= 1e-5 * ref_row.sum().item()
    assert sums[1].item() == valid.sum().item()
    onehot = torch.zeros(B, N, K, dtype=torch.float64).scatter_(2, codes.clamp(min=0).unsqueeze(2), 1.0)
    ref_dl = (logp.exp() - onehot) * valid.unsqueeze(2)
    assert (dl.double() - ref_dl).abs().max().item() <= 2e-6


def _torch_reference_loss(x, codes, mod, mask=None, return_pre=False):
    """prediction.py:9-82 restated with PyTorch ops (fp32, the device's library kernels).  With `mask` (B, N, H) the ReLU
    is applied as a product by that 0/1 mask (see the test below)."""
    N, K, H = mod.linear2_weight.shape
    c = codes.to(torch.int64)
    first = c[:, :-1].clamp(min=0) + torch.arange(0, (N - 1) * K, K, device=x.device)
    e = torch.nn.functional.embedding(first, mod.codebook_embedding.weight) * jo.embedding_scale(H, N)
    pre = torch.cumsum(torch.cat((mod.linear1(x).unsqueeze(1), e), dim=1), dim=1)
    if return_pre:
        return pre
    a = torch.relu(pre) if mask is None else pre * mask
    lg = torch.matmul(a.transpose(0, 1), mod.linear2_weight.transpose(1, 2)).transpose(0, 1)
    lg = lg + torch.matmul(x, mod.linear2b_weight.transpose(1, 2)).transpose(0, 1) + mod.linear2_bias
    return torch.nn.functional.cross_entropy(lg.reshape(-1, K), c.reshape(-1), ignore_index=mod.ignore_index,
                                             reduction=mod.reduction)


@pytest.mark.parametrize("checkpoint", [False, True])
def test_jcl_realistic_size_against_torch(checkpoint):
    """8 codebooks of 256 predicted from 512 channels, 8,192 frames, codes straight from Quantizer.encode (uint8); every
    dense product on tcgen05 (mcq_gemm_nt / mcq_gemm_tn).

    The function has 33 M ReLUs; two fp32 evaluations of linear1 differ in the last bits, so a handful of units whose
    pre-activation is within rounding of zero switch side, and each switch moves a few gradient entries by a finite
    amount.  That is adjudicated separately (few, and only at |pre-activation| < 1e-4); the gradients are then compared
    with the PyTorch evaluation using OUR ReLU pattern, which makes the comparison smooth."""
    from quantization_b200 import JointCodebookLoss, prediction, synth
    torch.manual_seed(3)
    P, N, K, H, B = 512, 8, 256, 512, 8192
    q = helpers.make_quantizer(256, N, K, synth.synth_params(256, N, K, 1), DEV)
    codes = q.encode(synth.synth_x(B, 256, 2).to(DEV), refine_indexes_iters=1)
    assert codes.dtype == torch.uint8
    mod = JointCodebookLoss(P, N, hidden_channels=H, codebook_size=K, checkpoint=checkpoint).to(DEV)
    x = torch.randn(B, P, device=DEV, requires_grad=True)
    loss = mod(x, codes)
    loss.backward()
    got = {n: p.grad.clone() for n, p in mod.named_parameters()}
    gx = x.grad.clone()
    mod.zero_grad()
    x.grad = None
    with torch.no_grad():
        _, _, act, _ = prediction._stages_forward(x.detach(), codes, mod.linear1.weight, mod.linear1.bias,
                                                  mod.codebook_embedding.weight, mod.linear2_weight, mod.linear2b_weight,
                                                  mod.linear2_bias, -100, False)
        ours_on = act.transpose(0, 1) > 0  # (B, N, H)
        pre = _torch_reference_loss(x, codes, mod, return_pre=True)
        switched = ours_on != (pre > 0)
        assert switched.sum().item() <= 1e-5 * switched.numel()
        if switched.any():
            assert pre[switched].abs().max().item() <= 1e-4
    ref = _torch_reference_loss(x, codes, mod, mask=ours_on.float())
    ref.backward()
    assert torch.allclose(loss, ref, rtol=1e-5)
    for n, p in mod.named_parameters():
        assert (got[n] - p.grad).abs().max().item() <= 1e-4 * p.grad.abs().max().item(), n
    assert (gx - x.grad).abs().max().item() <= 1e-4 * x.grad.abs().max().item()


def test_jcl_edge_cases():
    from quantization_b200 import JointCodebookLoss
    mod = JointCodebookLoss(16, 4, hidden_channels=32, codebook_size=16, reduction="sum").to(DEV)
    # no frames
    z = mod(torch.zeros(0, 16, device=DEV), torch.zeros(0, 4, dtype=torch.int64, device=DEV))
    assert z.item() == 0.0
    # every frame padded: loss 0, zero gradients
    x = torch.randn(5, 16, device=DEV, requires_grad=True)
    loss = mod(x, torch.full((5, 4), -100, device=DEV))
    loss.backward()
    assert loss.item() == 0.0 and x.grad.abs().max().item() == 0.0
    # shape mismatch and CPU tensors are rejected loudly
    with pytest.raises(AssertionError):
        mod(torch.randn(5, 16, device=DEV), torch.zeros(4, 4, dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError):
        mod(torch.randn(5, 16), torch.zeros(5, 4, dtype=torch.int64))
    with pytest.raises(_lib.McqError):
        L = _lib.lib()
        _lib.check(L.mcq_jcl_hidden_forward(0, 0, 1, 4, 4, 16, 30, 0, 1.0, 0, 0), "hidden")  # H not a multiple of 4
