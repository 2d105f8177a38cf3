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


# ---- SURVEY §8f rank 3.  PINNED: `tests/golden/make_features_golden.py` runs the reference's own zigzag,
# compute_dct_features, reorder_data, force_align and multistream_force_align and stores their outputs in
# `tests/golden/features.npz`; `tests/test_oracle_golden.py` checks these restatements against them. ----

def zigzag_order(rows, cols):
    """Row-major positions visited by `zigzag` (`utils/preprocessing.py:280-338`), in visiting order.  Closed form of the
    reference's state machine: anti-diagonals s = r + c, even s walked up-right (r falling), odd s down-left (r rising).
    The reference's walk steps off degenerate arrays (one row: the odd column 1 moves diagonally down, :305-308; one
    column: the even row 2 moves diagonally up-right, :322-325) and raises IndexError on the next read; so does this."""
    if (rows == 1 and cols >= 3) or (cols == 1 and rows >= 4):
        raise IndexError('zigzag walks off a %d x %d array' % (rows, cols))
    order = []
    for s in range(rows + cols - 1):
        rs = range(max(0, s - cols + 1), min(rows - 1, s) + 1)
        for r in (reversed(rs) if s % 2 == 0 else rs):
            order.append(r * cols + (s - r))
    return np.asarray(order, dtype=np.int64)


def zigzag(X):
    """`utils/preprocessing.py:280-338`."""
    rows, cols = X.shape
    return X.reshape(-1)[zigzag_order(rows, cols)]


def dct_ortho(X):
    """scipy.fftpack.dct(X, type=2, norm='ortho') over the last axis (`utils/preprocessing.py:427`), as the float64
    matrix product y[k] = s_k sum_n x[n] cos(pi k (2n+1) / (2N)), s_0 = sqrt(1/N), s_k = sqrt(2/N)."""
    N = X.shape[-1]
    n = np.arange(N, dtype=np.float64)
    B = np.cos(np.pi * np.outer(2 * n + 1, n) / (2 * N)) * np.sqrt(2.0 / N)
    B[:, 0] = np.sqrt(1.0 / N)
    return np.asarray(X, dtype=np.float64) @ B


def compute_dct_features(X, image_shape, no_coeff=30, method='zigzag'):
    """`utils/preprocessing.py:417-462` (float64 arithmetic; the reference's is the float32 FFT of scipy)."""
    X_dct = dct_ortho(X)
    if method == 'zigzag':
        order = zigzag_order(*image_shape)
        return X_dct[:, order[1:no_coeff + 1]]
    X_dct = X_dct[:, 1:]
    if method in ('rel_variance', 'variance'):
        score = np.std(X_dct - np.mean(X_dct, 0), 0) if method == 'rel_variance' else np.std(X_dct, 0)
    elif method == 'energy':
        score = np.sum(np.abs(X_dct), 0)
    else:
        raise NotImplementedError("method not implemented, use only 'zigzag', 'variance', 'rel_variance")
    idxs = np.argsort(score)[::-1][:no_coeff]
    return X_dct[:, idxs]


def reorder_data(X, shape, orig_order='f', desired_order='c'):
    """`utils/preprocessing.py:492-503`."""
    d1, d2 = shape
    return X.reshape((-1, d1, d2), order=orig_order).reshape((-1, d1 * d2), order=desired_order)


def force_align(x1, x2, mode='fill'):
    """`utils/preprocessing.py:607-660`, mode 'fill' (the only one the reference implements).  Keeps the reference's
    indexing of the fill frame of stream 2, which is relative to stream 1's length (:652): `x2[x2_curr_idx + l1 - 1]`."""
    x1, t1, lens1 = x1
    x2, t2, lens2 = x2
    n1, nt1, n2, nt2 = [], [], [], []
    i1 = i2 = 0
    for i, l1 in enumerate(lens1):
        l1 = int(l1)
        l2 = int(lens2[i])
        diff = l1 - l2
        if mode == 'fill':
            n1.extend(x1[i1:i1 + l1]); nt1.extend(t1[i1:i1 + l1])
            n2.extend(x2[i2:i2 + l2]); nt2.extend(t2[i2:i2 + l2])
            if diff < 0:
                n1.extend([x1[i1 + l1 - 1]] * -diff); nt1.extend([t1[i1 + l1 - 1]] * -diff)
                lens1[i] = l1 - diff
            else:
                if diff > 0:
                    n2.extend([x2[i2 + l1 - 1]] * diff); nt2.extend([t2[i2 + l2 - 1]] * diff)
                lens2[i] = l2 + diff
            i1 += l1
            i2 += l2
    return (np.array(n1), np.array(nt1), lens1), (np.array(n2), np.array(nt2), lens2)


def multistream_force_align(orig_streams, mode='fill'):
    """`utils/preprocessing.py:672-712`: every stream's utterance is extended to the longest stream's length by repeating
    its own last frame / target.  The length vectors are updated in place like the reference does."""
    inputs = [s[0] for s in orig_streams]
    targets = [s[1] for s in orig_streams]
    lens = [s[2] for s in orig_streams]
    new = [([], [], l) for l in lens]
    cur = [0] * len(orig_streams)
    for i in range(len(lens[0])):
        ls = [int(l[i]) for l in lens]
        longest = ls[int(np.argmax(ls))]
        for j in range(len(orig_streams)):
            l = ls[j]
            new[j][0].extend(inputs[j][cur[j]:cur[j] + l])
            new[j][1].extend(targets[j][cur[j]:cur[j] + l])
            new[j][0].extend([inputs[j][cur[j] + l - 1]] * (longest - l))
            new[j][1].extend([targets[j][cur[j] + l - 1]] * (longest - l))
            new[j][2][i] = longest
            cur[j] += l
    return [(np.array(a), np.array(b), c) for a, b, c in new]
