"""Update rules — mirrors `custom/updates.py:10-99` (generate_lr_map, adam_vlr) and the Lasagne rules the runners
use (`lasagne.updates.adam/adadelta/sgd/momentum/nesterov_momentum`).  They return a spec; the
arithmetic is the fused multi-tensor kernel `ipavsr_optim_step`."""
from ..function import UpdateSpec


def generate_lr_map(params, lr_config, default):
    """`custom/updates.py:10-32`: learning rate per parameter, keyed by the layer-name prefix of `param.name`."""
    lr_map = {}
    for param in params:
        layer_name = param.name[:param.name.rfind('.')]
        lr_map[param] = lr_config[layer_name] if layer_name in lr_config else default
    return lr_map


def adam_vlr(loss_or_grads, params, lr_map, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """`custom/updates.py:35-99`."""
    return UpdateSpec('adam', loss_or_grads, params, lr_map=lr_map, beta1=beta1, beta2=beta2, epsilon=epsilon)


def adam(loss_or_grads, params, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    return UpdateSpec('adam', loss_or_grads, params, learning_rate=learning_rate, beta1=beta1, beta2=beta2,
                      epsilon=epsilon)


def adadelta(loss_or_grads, params, learning_rate=1.0, rho=0.95, epsilon=1e-6):
    return UpdateSpec('adadelta', loss_or_grads, params, learning_rate=learning_rate, rho=rho, epsilon=epsilon)


def sgd(loss_or_grads, params, learning_rate):
    return UpdateSpec('sgd', loss_or_grads, params, learning_rate=learning_rate)


def momentum(loss_or_grads, params, learning_rate, momentum=0.9):
    return UpdateSpec('momentum', loss_or_grads, params, learning_rate=learning_rate, momentum=momentum)


def nesterov_momentum(loss_or_grads, params, learning_rate, momentum=0.9):
    return UpdateSpec('nesterov', loss_or_grads, params, learning_rate=learning_rate, momentum=momentum)
