"""Loss expressions — mirrors `custom/objectives.py:4-39` and Lasagne's `categorical_crossentropy`."""
from ..function import LossExpr


def temporal_softmax_loss(x, y, mask):
    """Masked per-frame cross-entropy that re-applies softmax to the (already softmaxed) network output and
    divides by the mask sum (`custom/objectives.py:27-37`).  x: `get_output(net)` of a frame-level net (N,T,V);
    y: (N,T) int targets; mask: the network's (N,T) uint8 mask variable."""
    return LossExpr('temporal_softmax', x, y, mask)


def categorical_crossentropy(predictions, targets):
    """Element-wise `-log p[n, y_n]`; wrap in `function.tensor.mean(...)` as the runners do
    (`avletters/trimodal.py:327`)."""
    return LossExpr('categorical_crossentropy_elemwise', predictions, targets, None)


def squared_error(a, b):
    """Element-wise `(a - b)**2` (lasagne.objectives.squared_error, the `objective_loss_function` of the auto-encoder
    fine-tuning nets, `avletters/trimodal.py:81`); wrap in `function.tensor.mean(...)`; an L2 penalty is added with
    `+ coefficient * ipavsr_b200.regularization.regularize_network_params(net, l2)`."""
    return LossExpr('squared_error_elemwise', a, b, None)
