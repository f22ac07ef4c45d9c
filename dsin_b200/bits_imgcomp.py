"""bitcost_to_bpp (/root/reference/src/bits_imgcomp.py:4-20): sum(bits) / (N*H*W)."""
import numpy as np


def num_pixels_in_input_batch(input_batch):
    assert int(input_batch.shape[1]) == 3, "Expected N3HW, got {}".format(tuple(input_batch.shape))
    return int(np.prod(input_batch.shape)) // 3


def bitcost_to_bpp(bit_cost, input_batch):
    assert bit_cost.dim() == input_batch.dim() == 4, "Expected NChw and N3HW"
    sums = getattr(bit_cost, "_dsin_sum", None)
    if sums is None:  # a plain tensor: reduce on device in fp64
        sums = bit_cost.double().sum().reshape(1)
    num_bits = float(sums.sum().item())
    return np.float32(num_bits / float(num_pixels_in_input_batch(input_batch)))
