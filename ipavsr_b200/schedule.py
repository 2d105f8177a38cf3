"""Issue-order analysis of a layer graph (host logic of the device engine; no CUDA needed, so it is unit-tested on CPU).

The reference compiles one Theano function per network and leaves the order of independent sub-graphs to Theano
(`runners/3stream.py:268-279`); here the order is explicit:

* `branch_assignment`: a layer that depends on exactly one non-mask input (a stream's encoder, its DeltaLayer, its LSTM —
  `modelzoo/adenet_3stream.py:145-238`) belongs to that input's branch and runs on the branch's CUDA stream; everything
  behind the fusion is the trunk.
* `lstm_sibling_groups`: consecutive LSTM layers of the walk that read the same input (the two directions of the
  aggregate BLSTM, `modelzoo/adenet_v2.py:79-92`) form a group: forward, all their input projections are computed before
  the first recurrence starts; backward, they are visited in the order their recurrences were launched.
"""
from . import layers as L


def _inputs(l):
    return [i for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)]) if i is not None]


def branch_assignment(layers, mask_layers):
    """layers: topological order (layers.get_all_layers).  Returns (branch_of, n_branches, trunk_fed):
    branch_of[layer] = index of the layer's branch (numbered by the position of its input layer in `layers`) or None for
    the trunk; trunk_fed = branch layers whose output (also) feeds a trunk layer — their gradient arrives from the trunk."""
    deps = {}
    for l in layers:
        if isinstance(l, L.InputLayer):
            deps[l] = frozenset() if l in mask_layers else frozenset([l])
        else:
            d = frozenset()
            for i in _inputs(l):
                d = d | deps.get(i, frozenset())
            deps[l] = d
    roots = sorted({next(iter(d)) for d in deps.values() if len(d) == 1}, key=layers.index)
    branch_of = {l: (roots.index(next(iter(deps[l]))) if len(deps[l]) == 1 else None) for l in layers}
    trunk_fed = set()
    for l in layers:
        if branch_of[l] is None:
            for i in _inputs(l):
                if branch_of.get(i) is not None:
                    trunk_fed.add(i)
    return branch_of, len(roots), trunk_fed


def lstm_sibling_groups(layers, branch_of, enabled=True):
    """Returns (siblings, backward_order): siblings[lstm] = the other LSTMs of its group; backward_order = reversed(layers)
    with every sibling group put back into forward order (still a valid reverse topological order: siblings do not feed
    each other)."""
    siblings = {}
    order = list(reversed(layers))
    n = len(layers)
    i = 0 if enabled else n
    while i < n:
        l = layers[i]
        j = i
        if isinstance(l, L.LSTMLayer):
            while (j + 1 < n and isinstance(layers[j + 1], L.LSTMLayer) and
                   layers[j + 1].input_layers[0] is l.input_layers[0] and branch_of[layers[j + 1]] == branch_of[l] and
                   not any(layers[j + 1] is x or x in _inputs(layers[j + 1]) for x in layers[i:j + 1])):
                j += 1
            if j > i:
                grp = tuple(layers[i:j + 1])
                for g in grp:
                    siblings[g] = tuple(x for x in grp if x is not g)
                order[n - 1 - j:n - i] = list(grp)          # forward order inside the group
        i = j + 1
    return siblings, order
