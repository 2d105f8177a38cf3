"""CPU ORACLE — test infrastructure only.  NumPy restatement of the runners' evaluation helpers:
frame-level majority vote (`runners/2stream_dct.py:48-81` evaluate_model2, same body in every runner) and sequence-level
argmax (`runners/1stream_noencoder.py:42-64` evaluate_model).  PINNED: `tests/golden/make_evaluate_golden.py` lifts the
reference's own two functions out of the runner files in the build container and stores their results in
`tests/golden/evaluate.npz`; `tests/test_oracle_golden.py` checks these restatements against those vectors.
"""
import numpy as np


def vote_predictions(output, mask):
    """Per utterance: argmax over classes of the first sum(mask) frames, votes per class, argmax of the votes
    (`runners/2stream_dct.py:62-72`); np.argmax semantics (first maximum) at both levels."""
    output = np.asarray(output)
    N, T, C = output.shape
    lens = np.asarray(mask).sum(axis=-1).astype(np.int64) if mask is not None else np.full(N, T, np.int64)
    frame = output.argmax(axis=-1)
    pred = np.zeros(N, dtype=np.int64)
    for i in range(N):
        pred[i] = np.bincount(frame[i, :lens[i]], minlength=C).argmax()
    return pred


def confusion(pred, y, C):
    """`runners/2stream_dct.py:77-79`: confusion_matrix[target, prediction] += 1."""
    m = np.zeros((C, C), dtype='int')
    np.add.at(m, (np.asarray(y, dtype=np.int64), np.asarray(pred, dtype=np.int64)), 1)
    return m


def evaluate_vote(output, y, mask):
    """(classification rate, confusion matrix) of the frame-level vote."""
    pred = vote_predictions(output, mask)
    return float((pred == np.asarray(y)).sum()) / float(len(pred)), confusion(pred, y, output.shape[-1])


def evaluate_sequence(output, y):
    """(classification rate, confusion matrix) of sequence-level outputs (N, C) (`runners/1stream_noencoder.py:52-64`)."""
    pred = np.asarray(output).argmax(axis=1)
    return float((pred == np.asarray(y)).sum()) / float(len(pred)), confusion(pred, y, output.shape[1])
