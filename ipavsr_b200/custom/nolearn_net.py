"""The slice of nolearn's `NeuralNet` the reference uses to fine-tune the DBNF auto-encoder before the encoder layers are
lifted into the stream networks (`avletters/trimodal.py:41-89,254-264`, `oulu/bimodal.py:34-82`, `oulu/unimodal.py:60-110`):

    dbn = NeuralNet(layers=[(InputLayer, {...}), (DenseLayer, {...}), ...], max_epochs=30,
                    objective_loss_function=squared_error, update=nesterov_momentum, regression=True,
                    update_learning_rate=0.001, update_momentum=0.05, objective_l2=0.005)
    dbn.initialize(); dbn.fit(X, X); recon = dbn.predict(X); dbn.get_all_layers()[1..4].W / .b

nolearn is a third-party dependency that is absent from the reference tree and unpinned there (README installs
`nolearn@master`, "0.7.git"); its published behaviour is restated: TrainSplit(eval_size=0.2) takes the FIRST fold of an
unshuffled 5-fold split as validation data for regression targets, BatchIterator(128) walks the rows in order (last
batch partial), the objective is `mean(loss_function(output, target)) + objective_l2 * sum of squares of the
regularizable parameters`, the epoch's train / valid losses are batch-size weighted means, `predict` runs the
deterministic forward in batches.  The dense stack runs on the engine's tcgen05 GEMMs (the encoder path);
the objective on `ipavsr_squared_error` + `ipavsr_l2_penalty`; the update on `ipavsr_optim_step`.
"""
import numpy as np

from .. import layers as L
from ..function import function, tensor as T
from ..regularization import regularize_network_params, l2
from . import objectives, updates as U


def train_split(n, eval_size=0.2):
    """Row indices (train, valid) of nolearn's TrainSplit for regression: the first fold of KFold(n, round(1/eval_size))."""
    if not eval_size:
        return np.arange(n), np.arange(0)
    folds = int(round(1.0 / eval_size))
    first = n // folds + (1 if n % folds else 0)
    return np.arange(first, n), np.arange(0, first)


class NeuralNet(object):
    def __init__(self, layers, max_epochs=100, objective_loss_function=None, update=None, regression=False, verbose=0,
                 objective_l2=0.0, batch_size=128, eval_size=0.2, **kwargs):
        if not regression:
            raise NotImplementedError('the reference only fine-tunes regression nets (auto-encoders) with NeuralNet')
        if objective_loss_function not in (None, objectives.squared_error):
            raise NotImplementedError('objective_loss_function must be squared_error')
        self.layer_specs, self.max_epochs, self.verbose = layers, int(max_epochs), verbose
        self.update = update or U.nesterov_momentum
        self.update_kwargs = {k[len('update_'):]: v for k, v in kwargs.items() if k.startswith('update_')}
        other = [k for k in kwargs if not k.startswith('update_')]
        if other:
            raise TypeError('unsupported NeuralNet arguments: %r' % (other,))
        self.objective_l2, self.batch_size, self.eval_size = float(objective_l2), int(batch_size), eval_size
        self.train_history_ = []
        self.layers_ = None

    # -- nolearn API ------------------------------------------------------------------------------------
    def initialize(self):
        if self.layers_ is not None:
            return
        layer, by_name, self._reshape = None, {}, None
        for cls, kw in self.layer_specs:
            kw = dict(kw)
            if cls is L.InputLayer:
                shape = tuple(kw.pop('shape'))
                if len(shape) != 2:
                    raise ValueError('NeuralNet input must be (None, D)')
                self.input_var = T.tensor3('X')
                # the engine's inputs are (N, T, D) sequences: rows are fed as N one-frame utterances
                self._input = L.InputLayer((None, None, shape[1]), self.input_var, name=kw.get('name'))
                layer = self._reshape = L.ReshapeLayer(self._input, (-1, shape[1]))
                by_name[kw.get('name')] = self._input
                continue
            if cls is not L.DenseLayer:
                raise NotImplementedError('NeuralNet supports InputLayer and DenseLayer specs')
            layer = L.DenseLayer(layer, **kw)
            by_name[kw.get('name')] = layer
        self._out, self.layers_ = layer, by_name
        self.target_var = T.matrix('y')
        params = L.get_all_params(layer, trainable=True)

        def objective(deterministic):
            loss = T.mean(objectives.squared_error(L.get_output(layer, deterministic=deterministic), self.target_var))
            if self.objective_l2:
                loss = loss + self.objective_l2 * regularize_network_params(layer, l2)
            return loss

        train_loss = objective(False)
        self.train_iter_ = function([self.input_var, self.target_var], train_loss,
                                    updates=self.update(train_loss, params, **self.update_kwargs))
        self.eval_iter_ = function([self.input_var, self.target_var], objective(True))
        self.predict_iter_ = function([self.input_var], L.get_output(layer, deterministic=True))

    def get_all_layers(self):
        """[input, dense1, ...] like nolearn (the builders read `[1..4].W/.b`, `modelzoo/deltanet.py:63-73`)."""
        self.initialize()
        return [l for l in L.get_all_layers(self._out) if l is not self._reshape]

    def get_all_params_values(self):
        self.initialize()
        return {name: [p.get_value() for p in l.get_params()] for name, l in self.layers_.items() if name is not None}

    def _batches(self, n):
        return [slice(i, min(i + self.batch_size, n)) for i in range(0, n, self.batch_size)]

    @staticmethod
    def _rows(X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        return X.reshape(X.shape[0], 1, -1)

    def fit(self, X, y, epochs=None):
        """nolearn's train loop.  The training matrix is uploaded ONCE and stays in HBM (the AVLetters / OuluVS frame matrices
        are 0.1-1 GB); every batch is a row-slice view of it, so a step moves no input bytes over the host link — only the
        loss comes back, like the reference's `train_iter_` return value."""
        import torch
        self.initialize()
        same = y is X
        X = np.ascontiguousarray(X, dtype=np.float32)
        dev = self.train_iter_.engine.device
        Xd = torch.from_numpy(X).to(dev)
        yd = Xd if same else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).to(dev)
        tr, va = train_split(len(X), self.eval_size)           # validation = the first len(va) rows, training = the rest
        Xt, yt, Xv, yv = Xd[len(va):], yd[len(va):], Xd[:len(va)], yd[:len(va)]
        for _ in range(epochs or self.max_epochs):
            tl, tn, vl, vn = [], [], [], []
            for s in self._batches(len(Xt)):
                tl.append(float(self.train_iter_(Xt[s].unsqueeze(1), yt[s])))
                tn.append(s.stop - s.start)
            for s in self._batches(len(Xv)):
                vl.append(float(self.eval_iter_(Xv[s].unsqueeze(1), yv[s])))
                vn.append(s.stop - s.start)
            info = {'epoch': len(self.train_history_) + 1, 'train_loss': float(np.average(tl, weights=tn)),
                    'valid_loss': float(np.average(vl, weights=vn)) if vl else float('nan')}
            self.train_history_.append(info)
            if self.verbose:
                print('%(epoch)6d  %(train_loss)12.5f  %(valid_loss)12.5f' % info)
        return self

    def predict_proba(self, X):
        """nolearn's `predict_proba`: the rows are independent, so the matrix goes up ONCE, runs in large row chunks and comes
        back in one copy (nolearn's own 128-row batches would each cost an upload, a launch-bound forward and a download).
        A CUDA tensor stays on the device: the result is a CUDA tensor."""
        import torch
        self.initialize()
        dev = self.predict_iter_.engine.device
        on_dev = isinstance(X, torch.Tensor) and X.is_cuda
        Xd = X.to(torch.float32) if on_dev else torch.from_numpy(np.ascontiguousarray(np.asarray(X), dtype=np.float32)).to(dev)
        Xd = Xd.reshape(Xd.shape[0], -1).contiguous()
        chunk = 16384
        outs = [self.predict_iter_(Xd[i:i + chunk].unsqueeze(1), device_output=True) for i in range(0, len(Xd), chunk)]
        res = torch.cat(outs, dim=0) if len(outs) != 1 else outs[0]
        return res if on_dev else res.cpu().numpy()

    predict = predict_proba
