"""quantization_b200 -- B200-native (sm_100a) implementation of the encode / decode / refine hot path of
danpovey/quantization, behind the reference's `Quantizer` / `QuantizerTrainer` API.  See DESIGN.md."""
from .quantizer import Quantizer  # noqa: F401
from .trainer import QuantizerTrainer  # noqa: F401
from .prediction import JointCodebookLoss  # noqa: F401
from .data import read_hdf5_data  # noqa: F401


def checkpoint(function, *args):
    """The reference package also exports `checkpoint` (quantization/__init__.py:5, checkpoint.py:43-44): run
    `function(*args)` without storing its intermediates and recompute them in backward.  Not on the hot path
    (JointCodebookLoss(checkpoint=True) recomputes inside its own autograd function); provided so that
    `from quantization import *` users find the same names: PyTorch's own implementation does the job."""
    import torch.utils.checkpoint as _cp
    return _cp.checkpoint(function, *args, use_reentrant=False)


__all__ = ["Quantizer", "QuantizerTrainer", "JointCodebookLoss", "read_hdf5_data", "checkpoint"]
