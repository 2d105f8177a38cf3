"""`utils/preprocessing.py` of the reference, hot-path subset (SURVEY §8a rows a10–a14), executed by the sm_100a
kernels of csrc/preprocess.cu.  Same names, argument meaning and return values; arrays go host -> device -> host per
call (use `ipavsr_b200.utils.device_pre` objects to keep a dataset resident in HBM instead).

No CPU fallback: these raise without a CUDA device.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(a, dtype=np.float32):
    if not torch.cuda.is_available():
        raise RuntimeError('ipavsr_b200.utils.preprocessing needs a CUDA device (there is no CPU path)')
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def _offsets(seqlens):
    lens = np.asarray(seqlens, dtype=np.int64)
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), lens


def normalize_input(input, centralize=True, quantize=False):
    """`utils/preprocessing.py:218-242` (sample-wise z-normalisation; `quantize` is not on the hot path)."""
    if quantize or not centralize:
        raise ValueError('only centralize=True, quantize=False is implemented on the device')
    x = _dev(input)
    frames, D = x.shape
    y = torch.empty_like(x)
    _lib.call('ipavsr_norm_samplewise', x.data_ptr(), D, y.data_ptr(), D, frames, D, _st())
    out = y.cpu().numpy()
    if isinstance(input, np.ndarray) and input.dtype == np.float32:
        input[...] = out                     # the reference normalises in place and returns the same array
        return input
    return out


def featurewise_normalize_sequence(input):
    """`utils/preprocessing.py:245-257`: returns (normalised, feature_means, feature_std)."""
    x = _dev(input)
    frames, F = x.shape
    mean = torch.empty(F, dtype=torch.float32, device='cuda')
    std = torch.empty(F, dtype=torch.float32, device='cuda')
    scratch = torch.empty(3 * F, dtype=torch.float64, device='cuda')
    _lib.call('ipavsr_norm_featurewise_stats', x.data_ptr(), F, mean.data_ptr(), std.data_ptr(), scratch.data_ptr(),
              frames, F, _st())
    y = torch.empty_like(x)
    _lib.call('ipavsr_norm_featurewise_apply', x.data_ptr(), F, mean.data_ptr(), std.data_ptr(), y.data_ptr(), F,
              frames, F, _st())
    return y.cpu().numpy(), mean.cpu().numpy(), std.cpu().numpy()


def featurewise_apply(input, mean, std):
    """(X - mean) / std with given statistics (`runners/2stream_dct.py:104-106`)."""
    x = _dev(input)
    frames, F = x.shape
    y = torch.empty_like(x)
    m, s = _dev(mean), _dev(std)
    _lib.call('ipavsr_norm_featurewise_apply', x.data_ptr(), F, m.data_ptr(), s.data_ptr(), y.data_ptr(), F, frames, F,
              _st())
    return y.cpu().numpy()


def _per_utterance(name, input, seqlens):
    offs, lens = _offsets(seqlens)
    x = _dev(input)
    frames, D = x.shape
    if offs[-1] != frames:
        raise ValueError('sequence lengths sum to %d but the input has %d frames' % (offs[-1], frames))
    y = torch.zeros_like(x)
    d_offs = torch.from_numpy(offs).cuda()
    U = len(lens)
    for u0 in range(0, U, 65535):
        n = min(65535, U - u0)
        _lib.call(name, x.data_ptr(), D, y.data_ptr(), D, d_offs.data_ptr() + 8 * u0, n, D, _st())
    return y.cpu().numpy()


def sequencewise_mean_image_subtraction(input, seqlens, axis=0):
    """`utils/preprocessing.py:260-277`."""
    if axis != 0:
        raise ValueError('only axis=0 is implemented')
    return _per_utterance('ipavsr_seq_mean_sub', input, seqlens)


def compute_diff_images(X, vidlenvec):
    """`utils/preprocessing.py:506-517`.  Every utterance needs at least 2 frames (the reference raises IndexError)."""
    if np.min(np.asarray(vidlenvec)) < 2:
        raise IndexError('compute_diff_images needs at least 2 frames per utterance')
    return _per_utterance('ipavsr_diff_image', X, vidlenvec)


def concat_first_second_deltas(X, vidlenvec, w=9):
    """`utils/preprocessing.py:465-489` (+ `deltas` :17-51, including its left-pad-with-column-1 quirk): float64
    (frames, 3F) = [x, d1, d2]."""
    offs, lens = _offsets(vidlenvec)
    x = _dev(X)
    frames, F = x.shape
    y = torch.zeros(frames, 3 * F, dtype=torch.float64, device='cuda')
    d_offs = torch.from_numpy(offs).cuda()
    U = len(lens)
    for u0 in range(0, U, 65535):
        n = min(65535, U - u0)
        _lib.call('ipavsr_deltas_fir', x.data_ptr(), F, y.data_ptr(), 3 * F, d_offs.data_ptr() + 8 * u0, n, F, int(w),
                  int(lens.max()), _st())
    return y.cpu().numpy()


def deltas(x, w=9):
    """`utils/preprocessing.py:17-51` on one (features x time) matrix."""
    x = np.asarray(x)
    out = concat_first_second_deltas(np.ascontiguousarray(x.T, dtype=np.float32), [x.shape[1]], w)
    F = x.shape[0]
    return np.ascontiguousarray(out[:, F:2 * F].T)
