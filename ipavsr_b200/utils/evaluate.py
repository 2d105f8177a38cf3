"""On-device mirror of the runners' evaluation helpers.

`evaluate_model` / `evaluate_model2` (+ the 2/3/4-stream argument orders, which the reference re-defines under the same
name in each runner) keep the reference's signatures and return values: classification rate and the (C, C) confusion
matrix of the per-utterance majority vote over the frame-level argmax, or of the sequence-level argmax.  The probabilities stay in HBM: the compiled function is
called with `device_output=True` and `ipavsr_vote_eval` (csrc/evaluate.cu) does the argmax / vote / confusion counting;
only the C*C counts and one integer come back.  There is no CPU path: the calls fail without the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_u8(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.uint8).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(np.uint8))).to(device)


def vote_predictions(output, mask=None, y=None):
    """output: (N, T, C) or (N, C) probabilities (device tensor or host array); mask (N, T) or None; y (N,) or None.
    Returns (pred int32 (N,) device tensor, confusion (C, C) int numpy or None, number correct or None)."""
    if not isinstance(output, torch.Tensor):
        output = torch.from_numpy(np.ascontiguousarray(np.asarray(output, dtype=np.float32))).cuda()
    output = output.to(torch.float32).contiguous()
    if output.dim() == 2:
        output = output.unsqueeze(1)
        mask = None
    N, T, Cn = output.shape
    dev = output.device
    m = _dev_u8(mask, dev) if mask is not None else None
    if m is not None and tuple(m.shape) != (N, T):
        raise ValueError('mask must be (N, T)')
    yt = _dev_u8(y, dev) if y is not None else None
    if yt is not None and yt.numel() != N:
        raise ValueError('one target per utterance')
    pred = torch.empty(N, dtype=torch.int32, device=dev)
    conf = torch.zeros(Cn, Cn, dtype=torch.int32, device=dev) if yt is not None else None
    corr = torch.zeros(1, dtype=torch.int32, device=dev) if yt is not None else None
    _lib.call('ipavsr_vote_eval', output.data_ptr(), Cn, m.data_ptr() if m is not None else None,
              yt.data_ptr() if yt is not None else None, N, T, Cn, pred.data_ptr(),
              conf.data_ptr() if conf is not None else None, corr.data_ptr() if corr is not None else None, _st())
    if yt is None:
        return pred, None, None
    return pred, conf.cpu().numpy().astype('int'), int(corr.item())


def _rate(ok, y_val):
    return ok / float(len(y_val))


def evaluate_model(X_val, y_val, mask_val, window, eval_fn):
    """Sequence-level outputs (N, C) (`runners/1stream_noencoder.py:42-64`): argmax, classification rate, confusion."""
    output = eval_fn(X_val, mask_val, window, device_output=True)
    _, conf, ok = vote_predictions(output, None, y_val)
    return _rate(ok, y_val), conf


def evaluate_model2(X_val, y_val, mask_val, window_size, eval_fn):
    """Frame-level vote, one stream (`runners/1stream.py:48-81`, `1stream_variable_lr.py:49-81`)."""
    output = eval_fn(X_val, mask_val, window_size, device_output=True)
    _, conf, ok = vote_predictions(output, mask_val, y_val)
    return _rate(ok, y_val), conf


def evaluate_model2_2stream(X_val, y_val, mask_val, X_diff_val, window_size, eval_fn):
    """`runners/2stream_dct.py:48-81`, `2stream.py:48` (their evaluate_model2)."""
    output = eval_fn(X_val, mask_val, X_diff_val, window_size, device_output=True)
    _, conf, ok = vote_predictions(output, mask_val, y_val)
    return _rate(ok, y_val), conf


def evaluate_model2_3stream(X_s1_val, X_s2_val, X_s3_val, y_val, mask_val, window_size, eval_fn):
    """`runners/3stream.py:48-83` (its evaluate_model2)."""
    output = eval_fn(X_s1_val, X_s2_val, X_s3_val, mask_val, window_size, device_output=True)
    _, conf, ok = vote_predictions(output, mask_val, y_val)
    return _rate(ok, y_val), conf


def evaluate_model2_4stream(X_s1_val, X_s2_val, X_s3_val, X_s4_val, y_val, mask_val, window_size, eval_fn):
    """`runners/4stream.py:52-88` (its evaluate_model2)."""
    output = eval_fn(X_s1_val, X_s2_val, X_s3_val, X_s4_val, mask_val, window_size, device_output=True)
    _, conf, ok = vote_predictions(output, mask_val, y_val)
    return _rate(ok, y_val), conf
