"""`utils/preprocessing.py` of the reference, hot-path subset (SURVEY §8a rows a10–a14), executed by the sm_100a
kernels of csrc/preprocess.cu.  Same names, argument meaning and return values.  A NumPy array goes host -> device -> host
per call and comes back as a NumPy array (the reference's contract); a CUDA `torch` tensor stays in HBM — the result is a
CUDA tensor and nothing is copied, so a dataset can be normalised / differenced / projected once, kept resident
(`utils.datagen.DeviceDataset`) and fed to the compiled functions without ever crossing the host link.

No CPU fallback: these raise without a CUDA device.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _is_dev(a):
    return isinstance(a, torch.Tensor) and a.is_cuda


def _dev(a, dtype=np.float32):
    if not torch.cuda.is_available():
        raise RuntimeError('ipavsr_b200.utils.preprocessing needs a CUDA device (there is no CPU path)')
    if isinstance(a, torch.Tensor):
        t = a.to(device='cuda', dtype=torch.from_numpy(np.zeros(0, dtype)).dtype)
        return t.contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def _out(t, like):
    """Result in the form of the input: CUDA tensor for a CUDA tensor (no copy), NumPy array otherwise."""
    return t if _is_dev(like) else t.cpu().numpy()


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
    if _is_dev(input):
        if input.dtype == torch.float32 and input.is_contiguous():
            input.copy_(y)                   # in place, like the reference
            return input
        return y
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
    return _out(y, input), _out(mean, input), _out(std, input)


def featurewise_apply(input, mean, std):
    """(X - mean) / std with given statistics (`runners/2stream_dct.py:104-106`)."""
    x = _dev(input)
    frames, F = x.shape
    y = torch.empty_like(x)
    m, s = _dev(mean), _dev(std)
    _lib.call('ipavsr_norm_featurewise_apply', x.data_ptr(), F, m.data_ptr(), s.data_ptr(), y.data_ptr(), F, frames, F,
              _st())
    return _out(y, input)


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
    return _out(y, input)


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
    return _out(y, X)


def deltas(x, w=9):
    """`utils/preprocessing.py:17-51` on one (features x time) matrix."""
    x = np.asarray(x)
    out = concat_first_second_deltas(np.ascontiguousarray(x.T, dtype=np.float32), [x.shape[1]], w)
    F = x.shape[0]
    return np.ascontiguousarray(out[:, F:2 * F].T)


# ---- zigzag / compute_dct_features / reorder_data / force_align on the device (csrc/features.cu) ----

def zigzag_order(rows, cols):
    """Row-major positions of the `zigzag` traversal (`utils/preprocessing.py:280-338`), from the C-ABI's host-side
    index walk.  IndexError where the reference's walk leaves the array."""
    order = np.empty(int(rows) * int(cols), dtype=np.int32)
    lib = _lib.load()
    if lib.ipavsr_zigzag_indices(int(rows), int(cols), order.ctypes.data_as(C.c_void_p)) != 0:
        raise IndexError(lib.ipavsr_last_error().decode())
    return order


def zigzag(X):
    """`utils/preprocessing.py:280-338` (index work on the host; the device path of compute_dct_features never
    materialises the traversal, it projects on the selected basis vectors only)."""
    X = np.asarray(X)
    return X.reshape(-1)[zigzag_order(*X.shape)]


def fill_zigzag(shape):
    """`utils/preprocessing.py:341-403`: the array whose zigzag traversal is 1..rows*cols."""
    rows, cols = shape
    out = np.zeros(rows * cols, dtype=np.float64)
    out[zigzag_order(rows, cols)] = np.arange(1, rows * cols + 1)
    return out.reshape(rows, cols)


def _dct_project(x, cols):
    """x (frames, D) on the device -> (frames, len(cols)) device tensor of the chosen DCT-II (ortho) coefficients."""
    frames, D = x.shape
    K = len(cols)
    d_cols = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int32)).cuda()
    ldb = (K + 3) // 4 * 4                                   # 16-byte rows: the cp.async path of the kernel
    basis = torch.empty(D, ldb, dtype=torch.float32, device='cuda')
    _lib.call('ipavsr_dct_basis', basis.data_ptr(), ldb, d_cols.data_ptr(), D, K, _st())
    out = torch.empty(frames, K, dtype=torch.float32, device='cuda')
    step = 65535 * 128
    for f0 in range(0, frames, step):
        n = min(step, frames - f0)
        _lib.call('ipavsr_dct_project', x.data_ptr() + 4 * D * f0, D, basis.data_ptr(), ldb, out.data_ptr() + 4 * K * f0,
                  K, n, D, K, _st())
    return out


def compute_dct_features(X, image_shape, no_coeff=30, method='zigzag'):
    """`utils/preprocessing.py:417-462`.  The reference's DCT is scipy's 1-D type-2 orthonormal DCT over the flattened
    image (:427); only the kept columns are computed here (zigzag positions 1..no_coeff), or — for the selection
    methods — every AC column, whose std / energy then ranks them (argsort on the host over D-1 numbers)."""
    if method not in ('zigzag', 'variance', 'rel_variance', 'energy'):
        raise NotImplementedError("method not implemented, use only 'zigzag', 'variance', 'rel_variance")
    x = _dev(X)
    frames, D = x.shape
    if method == 'zigzag':
        if int(image_shape[0]) * int(image_shape[1]) != D:
            raise ValueError('cannot reshape array of size %d into shape %r' % (D, tuple(image_shape)))
        cols = zigzag_order(*image_shape)[1:no_coeff + 1]
        return _out(_dct_project(x, cols), X)
    ac = _dct_project(x, np.arange(1, D))                     # X_dct[:, 1:]
    F = D - 1
    if method == 'energy':
        score = torch.empty(F, dtype=torch.float64, device='cuda')
        _lib.call('ipavsr_col_abs_sum', ac.data_ptr(), F, score.data_ptr(), frames, F, _st())
    else:                                                     # the std of mean-removed columns is the std of the columns
        mean = torch.empty(F, dtype=torch.float32, device='cuda')
        score = torch.empty(F, dtype=torch.float32, device='cuda')
        scratch = torch.empty(3 * F, dtype=torch.float64, device='cuda')
        _lib.call('ipavsr_norm_featurewise_stats', ac.data_ptr(), F, mean.data_ptr(), score.data_ptr(),
                  scratch.data_ptr(), frames, F, _st())
    idxs = np.argsort(score.cpu().numpy())[::-1][:no_coeff]
    d_idx = torch.from_numpy(np.ascontiguousarray(idxs, dtype=np.int32)).cuda()
    K = len(idxs)
    out = torch.empty(frames, K, dtype=torch.float32, device='cuda')
    _lib.call('ipavsr_gather_cols', ac.data_ptr(), F, d_idx.data_ptr(), out.data_ptr(), K, frames, K, _st())
    return _out(out, X)


def reorder_data(X, shape, orig_order='f', desired_order='c'):
    """`utils/preprocessing.py:492-503`: per-frame (d1, d2) images from Fortran to C packing or back."""
    d1, d2 = int(shape[0]), int(shape[1])
    orig_order, desired_order = orig_order.lower(), desired_order.lower()
    if orig_order not in 'fc' or desired_order not in 'fc':
        raise ValueError("order must be 'f' or 'c'")
    x = _dev(X.reshape(-1, d1 * d2) if _is_dev(X) else np.asarray(X).reshape(-1, d1 * d2))
    if orig_order == desired_order:
        return _out(x, X)
    y = torch.empty_like(x)
    _lib.call('ipavsr_reorder', x.data_ptr(), d1 * d2, y.data_ptr(), d1 * d2, x.shape[0], d1, d2,
              1 if desired_order == 'c' else 0, _st())
    return _out(y, X)


def _align_plan(lens_in, lens_out, fill_rel):
    """Offsets of the device gather and the same gather as a host row index (for the per-frame target vectors)."""
    lens_in = np.asarray(lens_in, dtype=np.int64)
    lens_out = np.asarray(lens_out, dtype=np.int64)
    in_off = np.concatenate([[0], np.cumsum(lens_in)]).astype(np.int64)
    out_off = np.concatenate([[0], np.cumsum(lens_out)]).astype(np.int64)
    fill = in_off[:-1] + np.asarray(fill_rel, dtype=np.int64)
    u = np.repeat(np.arange(len(lens_in)), lens_out)
    j = np.arange(out_off[-1]) - out_off[u]
    return in_off, out_off, fill, u, j


def _align_gather(x, in_off, out_off, fill):
    xd = _dev(x)
    rows_in, D = xd.shape
    if in_off[-1] != rows_in:
        raise ValueError('sequence lengths sum to %d but the stream has %d frames' % (in_off[-1], rows_in))
    out_rows = int(out_off[-1])
    y = torch.empty(out_rows, D, dtype=torch.float32, device='cuda')
    U = len(in_off) - 1
    if U > 0 and out_rows > 0:
        d_in, d_out, d_fill = (torch.from_numpy(a).cuda() for a in (in_off, out_off, fill))    # alive until the copy back
        _lib.call('ipavsr_align_fill', xd.data_ptr(), D, y.data_ptr(), D, d_in.data_ptr(), d_out.data_ptr(),
                  d_fill.data_ptr(), U, D, out_rows, _st())
    return y.cpu().numpy().astype(np.asarray(x).dtype, copy=False)


def force_align(x1, x2, mode='fill'):
    """`utils/preprocessing.py:607-660` (mode 'fill'; 'discard' is a TODO in the reference and returns empty streams
    there).  Stream 2's fill frame is `x2[x2_curr_idx + l1 - 1]` (:652) — indexed with stream 1's length, i.e. a frame
    of a LATER utterance whenever stream 1 is the longer one — reproduced, IndexError included; its fill target is the
    utterance's own last target (:653).  The length vectors are updated in place and returned, like the reference."""
    x1, t1, lens1 = x1
    x2, t2, lens2 = x2
    if mode != 'fill':
        raise NotImplementedError("only mode='fill' exists in the reference")
    l1 = np.asarray(lens1, dtype=np.int64).copy()
    l2 = np.asarray(lens2, dtype=np.int64).copy()
    lmax = np.maximum(l1, l2)
    in1, out1, fill1, u1, j1 = _align_plan(l1, lmax, l1 - 1)
    in2, out2, fill2, u2, j2 = _align_plan(l2, lmax, l1 - 1)                 # :652: l1, not l2
    if len(l2) and np.any((lmax > l2) & (fill2 >= in2[-1])):
        raise IndexError('index %d is out of bounds for axis 0 with size %d' % (fill2[(lmax > l2)].max(), in2[-1]))
    n1 = _align_gather(x1, in1, out1, fill1)
    n2 = _align_gather(x2, in2, out2, fill2)
    t1, t2 = np.asarray(t1), np.asarray(t2)
    nt1 = t1[in1[u1] + np.minimum(j1, l1[u1] - 1)]
    nt2 = t2[in2[u2] + np.minimum(j2, l2[u2] - 1)]                          # :653: the utterance's own last target
    for i in range(len(l1)):
        lens1[i] = lmax[i]
        lens2[i] = lmax[i]
    return (n1, nt1, lens1), (n2, nt2, lens2)


def multistream_force_align(orig_streams, mode='fill'):
    """`utils/preprocessing.py:672-712`: every stream's utterance is extended to the longest stream's length with copies
    of its own last frame and target; the length vectors are updated in place."""
    if mode != 'fill':
        raise NotImplementedError("only mode='fill' exists in the reference")
    lens = [np.asarray(s[2], dtype=np.int64).copy() for s in orig_streams]
    lmax = np.max(np.stack(lens), axis=0)
    res = []
    for (x, t, lvec), l in zip(orig_streams, lens):
        in_off, out_off, fill, u, j = _align_plan(l, lmax, l - 1)
        nx = _align_gather(x, in_off, out_off, fill)
        nt = np.asarray(t)[in_off[u] + np.minimum(j, l[u] - 1)]
        for i in range(len(l)):
            lvec[i] = lmax[i]
        res.append((nx, nt, lvec))
    return res
