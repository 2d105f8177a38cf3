"""CPU ORACLE — test infrastructure only.  NumPy restatement of the reference's in-memory batch builders
(`utils/datagen.py`: `compute_integral_len` :211-216, `gen_seq_batch_from_idx` :219-229, `gen_lstm_batch_random`
:92-153), SURVEY §8f rank 1.  PINNED: `tests/golden/make_datagen_golden.py` imports the reference's own
`utils/datagen.py` in the build container and stores its batches in `tests/golden/datagen.npz`;
`tests/test_oracle_golden.py` checks these restatements against those vectors.
"""
import numpy as np


def compute_integral_len(lengths):
    """`utils/datagen.py:211-216`: exclusive prefix sum of the utterance lengths (a Python list in the reference)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    out = np.zeros(len(lengths), dtype=np.int64)
    if len(lengths) > 1:
        out[1:] = np.cumsum(lengths[:-1])
    return out


def seq_batch_from_idx(data, idxs, seqlens, integral_lens, max_timesteps):
    """`utils/datagen.py:219-229`: rows of utterance idxs[i] followed by zero rows up to max_timesteps."""
    idxs = np.asarray(idxs, dtype=np.int64)
    out = np.zeros((len(idxs), int(max_timesteps), data.shape[-1]), dtype=data.dtype)
    for i, u in enumerate(idxs):
        n, s = int(seqlens[u]), int(integral_lens[u])
        out[i, :n] = data[s:s + n]
    return out


def lstm_batch(X, y, seqlen, idxs, max_timesteps=None):
    """One batch of `gen_lstm_batch_random` (`utils/datagen.py:127-141`) for the utterances `idxs`:
    (X_batch float (N,T,F), y_batch uint8 (N,) = label of the first frame, mask uint8 (N,T))."""
    seqlen = np.asarray(seqlen, dtype=np.int64)
    T = int(np.max(seqlen)) if max_timesteps is None else int(max_timesteps)
    integral = compute_integral_len(seqlen)
    xb = seq_batch_from_idx(X, idxs, seqlen, integral, T)
    idxs = np.asarray(idxs, dtype=np.int64)
    yb = np.asarray(y)[integral[idxs]].astype('uint8')
    mask = (np.arange(T)[None, :] < seqlen[idxs][:, None]).astype('uint8')
    return xb, yb, mask


def batch_schedule(n_utts, batchsize, shuffle=True, permutation=np.random.permutation):
    """Infinite generator of the index lists `gen_lstm_batch_random` draws (`utils/datagen.py:115-151`): one permutation
    per epoch (drawn up front, and again right after the batch that reaches the end), consecutive slices of `batchsize`,
    the batch whose end reaches or passes the last utterance takes the remainder and closes the epoch."""
    order = permutation(n_utts) if shuffle else np.arange(n_utts)
    start = 0
    while True:
        end = start + batchsize
        if end >= n_utts:
            batch = order[start:]
            order = permutation(n_utts) if shuffle else np.arange(n_utts)
            start = 0
        else:
            batch = order[start:end]
            start = end
        yield np.asarray(batch, dtype=np.int64)
