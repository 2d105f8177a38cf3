"""CPU ORACLE — test infrastructure only.  NumPy restatement of the six `utils/preprocessing.py` functions on
the hot path (SURVEY §8a rows a10–a14).  PINNED: `tests/golden/make_golden.py` imports the reference's own
`utils/preprocessing.py` in the build container and stores its outputs in `tests/golden/preprocessing.npz`;
`tests/test_oracle_golden.py` checks these restatements against those vectors.
"""
import numpy as np


def normalize_input(x, centralize=True, quantize=False):
    """`utils/preprocessing.py:218-242`: per-frame (x-mean)/std with population std; optional min-max."""
    out = np.array(x, copy=True)
    for i in range(out.shape[0]):
        item = out[i]
        if centralize:
            item = item - item.mean()
            item = item / np.std(item)
        if quantize:
            mn, mx = np.min(item), np.max(item)
            item = (item - mn) / (mx - mn)
        out[i] = item
    return out


def featurewise_normalize_sequence(x):
    """`utils/preprocessing.py:245-257`: returns (normalised, feature_means, feature_std)."""
    mean = np.mean(x, axis=0)
    x = x - mean
    std = np.std(x, axis=0)
    return x / std, mean, std


def sequencewise_mean_image_subtraction(x, seqlens):
    """`utils/preprocessing.py:260-277`: remove each utterance's mean frame.  All arithmetic stays in the input dtype
    (float32), as under the reference's NumPy 1.x value-based casting (`float32_array / integer_scalar`)."""
    out = np.zeros(x.shape, x.dtype)
    start = 0
    for l in seqlens:
        l = int(l)
        seq = x[start:start + l]
        out[start:start + l] = seq - np.sum(seq, 0, x.dtype) / x.dtype.type(l)
        start += l
    return out


def deltas(x, w=9):
    """`utils/preprocessing.py:17-51`: feacalc-style FIR slope on rows of x (features x time).
    d[t] = sum_{j=-h..h} j * xx[t+j] with right padding = last column and left padding = column index **1**
    (the reference's `x[:, 1]`, :43 — a mis-port of `dbn/deltas.m:20`)."""
    x = np.asarray(x)
    rows, cols = x.shape
    h = w // 2
    xx = np.concatenate([np.repeat(x[:, 1:2], h, 1), x, np.repeat(x[:, -1:], h, 1)], axis=1).astype(np.float64)
    d = np.zeros((rows, cols), np.float64)
    for j in range(-h, h + 1):
        d += j * xx[:, h + j: h + j + cols]
    return d


def concat_first_second_deltas(X, vidlenvec, w=9):
    """`utils/preprocessing.py:465-489`: per utterance [x, deltas(x), deltas(deltas(x))]; float64 output."""
    F = X.shape[1]
    Y = np.zeros((X.shape[0], 3 * F))
    start = 0
    for l in vidlenvec:
        seq = X[start:start + l]
        d1 = deltas(seq.T, w)
        d2 = deltas(d1, w)
        Y[start:start + l] = np.concatenate([seq, d1.T, d2.T], axis=1)
        start += l
    return Y


def compute_diff_images(X, vidlenvec):
    """`utils/preprocessing.py:506-517`: frame differences; frame 0 duplicates the first difference."""
    out = np.zeros(X.shape, X.dtype)
    start = 0
    for l in vidlenvec:
        l = int(l)
        seq = X[start:start + l]
        d = np.diff(seq, 1, 0)
        out[start] = d[0]
        out[start + 1:start + l] = d
        start += l
    return out
