"""quantization_b200 -- B200-native (sm_100a) implementation of the encode / decode / refine hot path of
danpovey/quantization, behind the reference's `Quantizer` / `QuantizerTrainer` API.  See DESIGN.md."""
from .quantizer import Quantizer  # noqa: F401
from .trainer import QuantizerTrainer  # noqa: F401
from .prediction import JointCodebookLoss  # noqa: F401
from .data import read_hdf5_data  # noqa: F401

__all__ = ["Quantizer", "QuantizerTrainer", "JointCodebookLoss", "read_hdf5_data"]
