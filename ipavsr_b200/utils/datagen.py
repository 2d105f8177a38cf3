"""Device-side mirror of the reference's in-memory batch builders (`utils/datagen.py`), SURVEY §8f rank 1.

The reference pads every utterance of every stream on the host each step (`np.concatenate` per utterance,
`utils/datagen.py:136`) and then uploads the padded batch.  Here the packed dataset lives in HBM (`DeviceDataset`) and a
batch is ONE gather kernel per stream (`ipavsr_batch_gather`, csrc/batch.cu) whose outputs feed the compiled functions
directly (they accept device tensors, so no host<->device copy is left in the step).  Same names, argument meaning,
shuffling (NumPy's global RNG: `np.random.seed` reproduces the reference's batch order) and return values; the arrays
are `torch` tensors on the device instead of NumPy arrays.  There is no CPU path: the calls fail without the CUDA library.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def compute_integral_len(lengths):
    """`utils/datagen.py:211-216` (returns a Python list like the reference)."""
    integral_lens = [0]
    for i in range(1, len(lengths)):
        integral_lens.append(integral_lens[i - 1] + int(lengths[i - 1]))
    return integral_lens


class DeviceDataset(object):
    """A packed variable-length dataset resident in HBM: `data` (total_frames, F) float32, optional per-frame labels
    `y` (uint8), per-utterance lengths `seqlens`; `integral_lens` as `compute_integral_len` gives them."""

    def __init__(self, data, seqlens, y=None, integral_lens=None, device=None):
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if isinstance(data, torch.Tensor):
            self.data = data.to(device=dev, dtype=torch.float32).contiguous()
        else:
            self.data = torch.from_numpy(np.ascontiguousarray(np.asarray(data), dtype=np.float32)).to(dev)
        assert self.data.dim() == 2, 'data must be (total_frames, F)'
        lens = np.asarray(seqlens, dtype=np.int64).reshape(-1)
        integ = np.asarray(compute_integral_len(lens) if integral_lens is None else integral_lens, dtype=np.int64)
        assert len(integ) == len(lens)
        assert len(lens) == 0 or int((integ + lens).max()) <= self.data.shape[0], 'utterances exceed the data matrix'
        self.seqlens_host, self.integral_host = lens, integ
        self.seqlens = torch.from_numpy(lens.astype(np.int32)).to(dev)
        self.integral = torch.from_numpy(integ).to(dev)
        self.y = None
        if y is not None:
            self.y = torch.from_numpy(np.ascontiguousarray(np.asarray(y).reshape(-1).astype(np.uint8))).to(dev)
            assert self.y.numel() == self.data.shape[0], 'one label per frame (the reference reads y[start])'
        self.max_timesteps = int(lens.max()) if len(lens) else 0

    @property
    def feature_len(self):
        return int(self.data.shape[1])

    def gather(self, idxs, max_timesteps=None, with_mask=False, with_labels=False):
        """(X_batch (N,T,F) float32[, mask (N,T) uint8][, y_batch (N,) uint8]) for the utterances `idxs`."""
        T = self.max_timesteps if max_timesteps is None else int(max_timesteps)
        idx_host = np.asarray(idxs, dtype=np.int64).reshape(-1)
        n = len(idx_host)
        if n and (idx_host.min() < 0 or idx_host.max() >= len(self.seqlens_host)):
            raise IndexError('utterance index out of range')
        if n and int(self.seqlens_host[idx_host].max()) > T:
            raise ValueError('max_timesteps is shorter than an utterance of the batch')
        dev = self.data.device
        idx = torch.from_numpy(idx_host.astype(np.int32)).to(dev, non_blocking=True)
        F = self.feature_len
        X = torch.empty(n, T, F, dtype=torch.float32, device=dev)
        mask = torch.empty(n, T, dtype=torch.uint8, device=dev) if with_mask else None
        yb = None
        if with_labels:
            if self.y is None:
                raise ValueError('the dataset holds no labels')
            yb = torch.empty(n, dtype=torch.uint8, device=dev)
        _lib.call('ipavsr_batch_gather', self.data.data_ptr(), F, self.integral.data_ptr(), self.seqlens.data_ptr(),
                  idx.data_ptr(), self.y.data_ptr() if yb is not None else None, X.data_ptr(), F,
                  mask.data_ptr() if mask is not None else None, yb.data_ptr() if yb is not None else None, n, T, F, _st())
        out = (X,)
        if with_mask:
            # the lengths are known on the host: the engine's packed execution (engine._PackPlan) reads them from the mask
            # tensor instead of synchronising with the device to recover them
            mask._ipavsr_lens = self.seqlens_host[idx_host].copy()
            out += (mask,)
        if with_labels:
            out += (yb,)
        return out if len(out) > 1 else X


def gen_seq_batch_from_idx(data, idxs, seqlens, integral_lens, max_timesteps):
    """`utils/datagen.py:219-229`.  `data`: a `DeviceDataset`, a CUDA tensor or a host array (uploaded once per call —
    keep a `DeviceDataset` across steps to avoid that)."""
    ds = data if isinstance(data, DeviceDataset) else DeviceDataset(data, seqlens, integral_lens=integral_lens)
    return ds.gather(idxs, max_timesteps)


def gen_lstm_batch_random(X, y, seqlen, batchsize=30, shuffle=True):
    """`utils/datagen.py:92-153`: infinite generator of (X_batch, y_batch, mask, batch_video_idxs).  `X` may be a
    `DeviceDataset` (then `y`/`seqlen` may be None) or the packed host arrays, which are uploaded once."""
    ds = X if isinstance(X, DeviceDataset) else DeviceDataset(X, seqlen, y=y)
    no_videos = len(ds.seqlens_host)
    max_timesteps = ds.max_timesteps
    start_video = 0
    randomized = np.random.permutation(no_videos) if shuffle else np.arange(no_videos)
    while True:
        end_video = start_video + batchsize
        reset = end_video >= no_videos            # all videos iterated: this batch takes the remainder
        batch_video_idxs = randomized[start_video:] if reset else randomized[start_video:end_video]
        X_batch, mask, y_batch = ds.gather(batch_video_idxs, max_timesteps, with_mask=True, with_labels=True)
        if reset:
            randomized = np.random.permutation(no_videos) if shuffle else np.arange(no_videos)
            start_video = 0
        else:
            start_video = end_video
        yield X_batch, y_batch, mask, batch_video_idxs
