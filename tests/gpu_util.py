"""Helpers for the -m gpu parity tests: device buffers (torch is only the allocator) + C-ABI calls."""
import ctypes as C

import numpy as np
import torch

from ipavsr_b200 import _lib


def lib():
    return _lib.load()


def dev(a, dtype=None):
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a).astype(dtype))
    return torch.from_numpy(a).cuda()


def zeros(shape, dtype=torch.float32):
    return torch.zeros(shape, dtype=dtype, device='cuda')


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    _lib.call(name, *args)


F16_LO_SCALE = 1.0      # x 2^e = hi + lo / F16_LO_SCALE (csrc/common.cuh)


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def interleave_gates(W, H):
    """Lasagne-stacked [.. , 4H] = [i|f|c|o] blocks -> device layout column 4u+g."""
    W = np.asarray(W)
    lead = W.shape[:-1]
    return np.ascontiguousarray(W.reshape(lead + (4, H)).swapaxes(-1, -2).reshape(lead + (4 * H,)))


def deinterleave_gates(W, H):
    W = np.asarray(W)
    lead = W.shape[:-1]
    return np.ascontiguousarray(W.reshape(lead + (H, 4)).swapaxes(-1, -2).reshape(lead + (4 * H,)))
