"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/mcq.h declares, host-side
argument validation, and the Python mirror of the reference interface."""
import os
import re

import numpy as np
import pytest
import torch

import oracle
from quantization_b200 import Quantizer, QuantizerTrainer, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mcq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mcq.h but not exported by libmcq.so"
    assert sorted(_lib.EXPORTS) == names
    assert L.mcq_version() >= 1


def test_packed_cols_matches_reference_packing():
    L = _lib.lib()
    for K, N in ((256, 8), (16, 8), (16, 2), (4, 16), (2, 32), (16, 1), (64, 4), (256, 1)):
        idx = np.zeros((3, N), np.int64)
        assert L.mcq_packed_cols(N, K) == oracle.pack(idx, K).shape[1]


def test_sizes_and_shape_validation():
    L = _lib.lib()
    assert L.mcq_prepared_bytes(8, 256, 512) > 2048 * 2048 * 4
    assert L.mcq_workspace_bytes(1 << 20, 512, 8, 256) >= L.mcq_workspace_bytes(1024, 512, 8, 256) > 0
    assert L.mcq_prepared_bytes(3, 256, 512) == 0  # not a power of two (quantization.py:33-36)
    assert L.mcq_prepared_bytes(2, 4, 512) == 0    # K < 16 with N > 1: the reference raises too
    assert b"UnboundLocalError" in L.mcq_last_error()
    # null pointers / bad arguments are rejected before any CUDA call
    assert L.mcq_encode(None, 0, 16, 512, 8, 256, None, 5, None, 0, None, 0, None) == -1
    assert L.mcq_encode(None, 0, 0, 512, 8, 256, None, 5, None, 0, None, 0, None) == 0  # empty batch is a no-op
    assert L.mcq_decode(None, 0, 4, 3, 8, 256, 512, None, None, 0, None) == -1  # 3 columns do not divide 8
    # host-buffer encode, re-entrant form: staging size = two frame buffers + two code buffers + one encode workspace
    # of the largest chunk (75,776 frames), monotone in the batch up to that chunk; bad arguments rejected up front
    big = L.mcq_encode_host_ws_bytes(1 << 20, 512, 8, 256, _lib.F32, _lib.U8)
    small = L.mcq_encode_host_ws_bytes(1000, 512, 8, 256, _lib.F32, _lib.U8)
    assert big == L.mcq_encode_host_ws_bytes(1 << 22, 512, 8, 256, _lib.F32, _lib.U8) > small > 0
    assert big >= 2 * 75776 * 512 * 4 + L.mcq_workspace_bytes(75776, 512, 8, 256)
    assert L.mcq_encode_host_ws_bytes(1000, 512, 3, 256, _lib.F32, _lib.U8) == 0
    assert L.mcq_encode_host_ws(None, 0, 16, 512, 8, 256, None, 5, None, 0, None, 0, None) == -1
    assert L.mcq_encode_host_ws(None, 0, 0, 512, 8, 256, None, 5, None, 0, None, 0, None) == 0


def test_quantizer_surface_and_init_rng_order():
    torch.manual_seed(123)
    q = Quantizer(dim=32, codebook_size=16, num_codebooks=4)
    torch.manual_seed(123)
    lin = torch.nn.Linear(32, 64)
    assert torch.equal(q.to_logits.weight, lin.weight) and torch.equal(q.to_logits.bias, lin.bias)
    assert torch.equal(q.centers.reshape(64, 32), lin.weight)  # centers start as a clone of the classifier weight
    assert list(q.state_dict().keys()) == ["centers", "logits_scale", "centers_scale", "id_buf", "to_logits.weight",
                                           "to_logits.bias"]
    assert len(q.get_id()) == 8 and q.scale_speed == 10.0
    assert q.show_init_invocation() == "quantization.Quantizer(dim=32, codebook_size=16, num_codebooks=4)"
    q2 = Quantizer(32, 16, 4)
    q2.load_state_dict(q.state_dict())
    assert q2.get_id() == q.get_id()
    with pytest.raises(AssertionError):
        Quantizer(32, 12, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q.encode(torch.zeros(2, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        q.decode(torch.zeros(2, 4, dtype=torch.int64))
    assert tuple(q.get_centers().shape) == (4, 16, 32) and tuple(q.get_data_mean().shape) == (32,)
    c = q.compute_codebook_correlations()
    assert tuple(c.shape) == (4, 4) and torch.allclose(c.diag(), torch.ones(4), atol=1e-5)


def test_separate_indexes_matches_oracle_unpack():
    q = Quantizer(8, 16, 8)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 16, size=(11, 8), dtype=np.int64)
    packed = oracle.pack(idx, 16)
    out = q._maybe_separate_indexes(torch.from_numpy(packed))
    assert np.array_equal(out.numpy(), idx)


def test_product_quantizer_matches_loop_definition():
    torch.manual_seed(0)
    q = Quantizer(dim=6, codebook_size=4, num_codebooks=4)
    with torch.no_grad():
        q.logits_scale.fill_(0.03)
        q.centers_scale.fill_(-0.02)
    p = q.get_product_quantizer()
    assert (p.codebook_size, p.num_codebooks, p.dim) == (16, 2, 6)
    assert float(p.logits_scale) == float(q.logits_scale) and float(p.centers_scale) == float(q.centers_scale)
    K = 4
    for c_out in range(2):  # the reference's triple loop (quantization.py:98-111), restated
        for k1 in range(K):
            for k2 in range(K):
                r1, r2, ro = K * (2 * c_out) + k1, K * (2 * c_out + 1) + k2, 16 * c_out + k1 * K + k2
                assert torch.equal(p.to_logits.weight[ro], q.to_logits.weight[r1] + q.to_logits.weight[r2])
                assert torch.equal(p.to_logits.bias[ro], q.to_logits.bias[r1] + q.to_logits.bias[r2])
                assert torch.equal(p.centers[c_out, k1 * K + k2], q.centers[2 * c_out, k1] + q.centers[2 * c_out + 1, k2])


def test_trainer_construction_matches_reference_settings():
    tr = QuantizerTrainer(dim=16, bytes_per_frame=2, device=torch.device("cpu"), phase_one_iters=40,
                          phase_two_iters=80, lr=0.01)
    assert (tr.quantizer.codebook_size, tr.quantizer.num_codebooks) == (16, 4)
    g = tr.optim.param_groups[0]
    assert g["lr"] == 0.01 and g["betas"] == (0.9, 0.98) and g["eps"] == 1e-9 and g["weight_decay"] == 1e-6
    assert tr.scheduler.step_size == 10 and tr.scheduler.gamma == 0.5 and tr.two_iter_prob == 0.5
    assert not tr.done()
    with pytest.raises(AssertionError):
        QuantizerTrainer(dim=16, bytes_per_frame=3, device=torch.device("cpu"))


def test_joint_codebook_loss_mirror_has_the_reference_parameters():
    """quantization_b200.JointCodebookLoss keeps the reference module's parameter names, shapes and construction
    (prediction.py:128-152): the state_dict of the reference-generated fixture loads unchanged.  (No compute: there is no
    CPU path; a CPU tensor is rejected loudly.)"""
    import numpy as np
    import pytest
    import torch
    import helpers
    from quantization_b200 import JointCodebookLoss
    g, meta = helpers.jcl_golden()
    for name in helpers.jcl_case_names():
        m, pred, codes, par, _ = helpers.jcl_case(g, meta, name)
        mod = JointCodebookLoss(m["P"], m["N"], hidden_channels=m["H"], codebook_size=m["K"], reduction=m["reduction"])
        sd = mod.state_dict()
        assert set(sd.keys()) == set(par.keys())
        for k, v in par.items():
            assert tuple(sd[k].shape) == v.shape, k
        mod.load_state_dict({k: torch.from_numpy(v) for k, v in par.items()})
        assert np.array_equal(mod.linear2_weight.detach().numpy(), par["linear2_weight"])
    with pytest.raises(RuntimeError):
        mod(torch.from_numpy(pred), torch.from_numpy(codes))
    with pytest.raises(AssertionError):
        JointCodebookLoss(8, 1)  # num_codebooks must be > 1 (prediction.py:128)


def test_read_hdf5_data_with_a_stand_in_h5py(monkeypatch):
    """read_hdf5_data (reference quantization.py:744-820) against its definition, with a minimal stand-in for h5py (the
    package is not in this image): datasets of shape (*, dim) are flattened in key order, cast to float16, shuffled by
    one np.random.shuffle call, and split into (train, valid) with valid = the first min(5 %, 10000) rows."""
    import sys
    import types
    import numpy as np
    import torch
    from quantization_b200 import read_hdf5_data

    rng = np.random.default_rng(0)
    sets = {"dataset_0": rng.standard_normal((7, 3, 4)).astype(np.float32),
            "dataset_1": rng.standard_normal((50, 4)).astype(np.float16),
            "dataset_2": rng.standard_normal((2, 5, 8, 4)).astype(np.float64)}

    class FakeFile(dict):
        def __init__(self, name, mode):
            assert mode == "r"
            super().__init__(sets)

    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=FakeFile))
    np.random.seed(123)
    train, valid = read_hdf5_data("whatever.h5")
    flat = np.concatenate([np.ascontiguousarray(v).reshape(-1, 4).astype(np.float16) for v in sets.values()])
    np.random.seed(123)
    np.random.shuffle(flat)
    n_valid = int(0.05 * len(flat))
    assert train.dtype == torch.float16 and valid.dtype == torch.float16
    assert np.array_equal(valid.numpy(), flat[:n_valid]) and np.array_equal(train.numpy(), flat[n_valid:])
    assert len(flat) == 7 * 3 + 50 + 2 * 5 * 8 and n_valid == 7
    # the 10000-row cap
    sets.clear()
    sets["big"] = np.zeros((250000, 2), dtype=np.float16)
    train, valid = read_hdf5_data("big.h5")
    assert valid.shape == (10000, 2) and train.shape == (240000, 2)
