"""the remaining modelzoo variants (adenet_v1_1, v2_1..v2_4, v2_nodelta, v4-v6, adenet_2stream incl. pretrained
sub-stream LSTMs, 3stream_dct / _dropout / create_pretrained_model, avnet, lstm_classifier_majority_vote) through the same
engine, against the NumPy oracle: probabilities, argmax, loss and every parameter gradient.  Run with -m gpu on the B200."""
import pytest

import model_util as MU
from test_gpu_models import _check_forward_backward

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', ['fp32', 'f16x3'])
@pytest.mark.parametrize('name', MU.VARIANTS)
def test_builder_variants_forward_backward_parity(name, mode):
    fus = {'adenet_v5': ['sum', 'adasum'], 'adenet_v2_3': ['adasum'], 'adenet_v4': ['sum'], 'adenet_v1_1': ['sum'],
           'adenet_v2_2': ['concat', 'adasum']}.get(name, ['concat'])
    _check_forward_backward(name, fus, mode)
