"""Generates tests/golden/evaluate.npz by running the REFERENCE's own evaluation helpers on seeded network outputs.
Run in the build container only:

    python tests/golden/make_evaluate_golden.py

The helpers live in runner scripts that import Theano/Lasagne/matplotlib at module level, so the two function
definitions are lifted out of the files at run time (ast, from /root/reference, read-only) and executed with NumPy alone:
`evaluate_model2` of runners/2stream_dct.py:48-81 (frame-level majority vote, two input streams) and `evaluate_model` of
runners/1stream_noencoder.py:42-64 (sequence-level argmax).  Nothing under /root/reference is modified or copied.
"""
import ast
import os

import numpy as np


def lift(path, name):
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {'np': np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
            return ns[name]
    raise KeyError(name)


def main():
    vote = lift('/root/reference/runners/2stream_dct.py', 'evaluate_model2')
    seq = lift('/root/reference/runners/1stream_noencoder.py', 'evaluate_model')
    rng = np.random.default_rng(20261017)
    out = {}
    N, T, C = 23, 17, 7
    lens = rng.integers(1, T + 1, size=N)
    lens[0], lens[1] = T, 1
    mask = (np.arange(T)[None, :] < lens[:, None]).astype('uint8')
    probs = rng.random((N, T, C)).astype('float32')
    probs /= probs.sum(-1, keepdims=True)
    # ties: utterance 2 votes 2:2 between classes 5 and 1 (lowest index must win), utterance 3 has frames whose two
    # largest probabilities are equal (first maximum must win)
    lens[2] = 4
    mask[2] = (np.arange(T) < 4)
    probs[2, :4] = 0.01
    probs[2, 0, 5] = probs[2, 1, 1] = probs[2, 2, 5] = probs[2, 3, 1] = 0.9
    probs[3, :, :] = 0.05
    probs[3, :, 4] = probs[3, :, 2] = 0.4
    y = rng.integers(0, C, size=N).astype('uint8')
    rate, conf = vote(np.zeros((N, T, 3), 'float32'), y, mask, None, 9, lambda a, m, b, w: probs)
    out.update(vote_probs=probs, vote_mask=mask, vote_y=y, vote_rate=np.float64(rate), vote_conf=conf)
    so = rng.random((31, 10)).astype('float32')
    so[5, 3] = so[5, 8] = 2.0                    # tie: the first maximum wins
    ys = rng.integers(0, 10, size=31).astype('uint8')
    rate, conf = seq(np.zeros((31, 5, 3), 'float32'), ys, None, 9, lambda a, m, w: so)
    out.update(seq_probs=so, seq_y=ys, seq_rate=np.float64(rate), seq_conf=conf)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'evaluate.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    main()
