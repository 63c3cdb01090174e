"""`read_hdf5_data`: the reference's training-data reader (quantization/quantization.py:744-820), host side only.

Reads every dataset of an HDF5 archive (each of shape (*, dim), any float type) into one (tot_frames, dim) float16
array, shuffles the rows with ONE `np.random.shuffle` call (so a seeded numpy RNG gives the reference's order), and
returns `(train, valid)` CPU tensors where `valid` is the first min(5 % of the frames, 10000) shuffled rows and `train`
the rest.  `h5py` is imported when the function is called; it is not part of this image, and the function raises
ImportError without it (nothing else in the package needs it).

One deliberate difference: the reference computes the validation size as a float (`0.05 * tot_frames`, :811-814) and
slices with it, which raises TypeError whenever 5 % of the frames is <= 10000; here the size is truncated to an int, which
is what the reference does implicitly in the only case where it works (the 10000 cap).
"""
import logging
from typing import Tuple

import numpy as np
import torch


def read_hdf5_data(filename: str) -> Tuple[torch.Tensor, torch.Tensor]:
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - depends on the environment
        raise ImportError("read_hdf5_data needs the h5py package (the reference imports it at module level, "
                          "quantization.py:2)") from e
    logging.info(f"Opening file {filename}")
    hf = h5py.File(filename, "r")
    keys = list(hf.keys())
    dim = -1
    tot_frames = 0
    for key in keys:  # first pass: sizes (reference :775-783)
        shape = list(hf[key].shape)
        if dim == -1:
            dim = shape[-1]
        else:
            assert dim == shape[-1], "Dataset must have consistent dimension (last element of shape"
        tot_frames += int(np.prod(shape[:-1], dtype=np.int64))
    logging.info(f"read_data: tot_frames = {tot_frames}")

    ans = np.empty((tot_frames, max(dim, 0)), dtype=np.float16)
    cur_pos = 0
    for key in keys:  # second pass: contents, in key order (reference :787-794)
        array = np.ascontiguousarray(hf[key][:]).reshape(-1, dim)
        ans[cur_pos:cur_pos + array.shape[0], :] = array
        cur_pos += array.shape[0]
    assert cur_pos == tot_frames

    np.random.shuffle(ans)  # rows, in place: the reference's single RNG call (:797)
    ans_torch = torch.from_numpy(ans)

    valid_frames = min(int(0.05 * tot_frames), 10000)
    train_frames = tot_frames - valid_frames
    logging.info(f"read_data: train_frames={train_frames}, valid_frames={valid_frames}")
    return ans_torch[valid_frames:tot_frames], ans_torch[:valid_frames]
