"""`.ini` options that reach the hot path (SURVEY §5 "Config / flags", §8b): both schema generations of the reference.

modern:  [stream1..4] shape, nonlinearities, input_dimensions, model ... / [lstm_classifier] fusiontype, weight_init,
         use_peepholes, windowsize, output_classes, lstm_size, use_blstm ... / [training] learning_rate ...
         (`runners/2stream_dct.py:151-199`, `runners/4stream.py:159-224`)
legacy:  [models] fusiontype / [training] lstm_units, output_units, learning_rate, decay_rate, decay_start,
         do_finetune, save_finetune, load_finetune(_diff)   (`avletters/trimodal.py:198-211`, `oulu/unimodal.py:224-234`)
Same key names and defaults; values are returned as a plain dict the builders' kwargs are filled from.
"""
try:
    import configparser
except ImportError:                      # pragma: no cover
    import ConfigParser as configparser

from . import init


def weight_init_fn(name):
    """`runners/2stream_dct.py:187-195`: glorot | norm | uniform | ortho."""
    table = {'glorot': init.GlorotUniform(), 'norm': init.Normal(0.1), 'uniform': init.Uniform(),
             'ortho': init.Orthogonal()}
    if name not in table:
        raise ValueError('unknown weight_init %r' % (name,))
    return table[name]


def _get(cp, sec, key, default=None, conv=str):
    if cp.has_section(sec) and cp.has_option(sec, key):
        if conv is bool:
            return cp.getboolean(sec, key)
        return conv(cp.get(sec, key))
    return default


def read(path_or_file):
    cp = configparser.ConfigParser()
    if hasattr(path_or_file, 'read'):
        cp.read_file(path_or_file)
    else:
        if not cp.read(path_or_file):
            raise IOError('cannot read config %r' % (path_or_file,))
    return cp


def model_options(cp):
    """Options that select / size the network, from either schema generation."""
    o = {}
    o['fusiontype'] = _get(cp, 'lstm_classifier', 'fusiontype', _get(cp, 'models', 'fusiontype', 'sum'))
    o['lstm_size'] = _get(cp, 'lstm_classifier', 'lstm_size', _get(cp, 'training', 'lstm_units', 250, int), int)
    o['output_classes'] = _get(cp, 'lstm_classifier', 'output_classes', _get(cp, 'training', 'output_units', 26, int), int)
    o['use_peepholes'] = _get(cp, 'lstm_classifier', 'use_peepholes', False, bool)
    o['use_blstm'] = _get(cp, 'lstm_classifier', 'use_blstm', True, bool)
    o['use_blstm_substream'] = _get(cp, 'lstm_classifier', 'use_blstm_substream', False, bool)
    o['use_dropout'] = _get(cp, 'lstm_classifier', 'use_dropout', False, bool)
    o['windowsize'] = _get(cp, 'lstm_classifier', 'windowsize', 9, int)
    o['weight_init'] = _get(cp, 'lstm_classifier', 'weight_init', 'glorot')
    o['matlab_target_offset'] = _get(cp, 'lstm_classifier', 'matlab_target_offset', True, bool)
    streams = []
    for k in range(1, 5):
        sec = 'stream%d' % k
        if not cp.has_section(sec):
            continue
        streams.append({
            'data': _get(cp, sec, 'data'), 'model': _get(cp, sec, 'model'),
            'shape': _get(cp, sec, 'shape', '2000,1000,500,50'),
            'nonlinearities': _get(cp, sec, 'nonlinearities', 'rectify,rectify,rectify,linear'),
            'input_dimensions': _get(cp, sec, 'input_dimensions', None, int),
            'imagesize': _get(cp, sec, 'imagesize'), 'has_encoder': _get(cp, sec, 'has_encoder', True, bool),
            'reorderdata': _get(cp, sec, 'reorderdata', False, bool), 'diffimage': _get(cp, sec, 'diffimage', False, bool),
            'meanremove': _get(cp, sec, 'meanremove', False, bool),
            'samplewisenormalize': _get(cp, sec, 'samplewisenormalize', False, bool),
            'featurewisenormalize': _get(cp, sec, 'featurewisenormalize', False, bool),
            'lstm_model': _get(cp, sec, 'lstm_model')})
    o['streams'] = streams
    return o


def training_options(cp):
    o = {}
    o['learning_rate'] = _get(cp, 'training', 'learning_rate', 1e-3, float)
    o['decay_rate'] = _get(cp, 'training', 'decay_rate', 1.0, float)
    o['decay_start'] = _get(cp, 'training', 'decay_start', 0, int)
    o['num_epoch'] = _get(cp, 'training', 'num_epoch', 25, int)
    o['epochsize'] = _get(cp, 'training', 'epochsize', 20, int)
    o['batchsize'] = _get(cp, 'training', 'batchsize', 26, int)
    o['validation_window'] = _get(cp, 'training', 'validation_window', 4, int)
    o['update_rule'] = _get(cp, 'training', 'update_rule', 'adam')
    o['momentum'] = _get(cp, 'training', 'momentum', 0.9, float)
    for k in ('do_finetune', 'save_finetune', 'load_finetune', 'load_finetune_diff'):
        o[k] = _get(cp, 'training', k, False, bool)
    return o
