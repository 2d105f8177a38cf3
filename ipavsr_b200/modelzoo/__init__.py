"""Model builders with the reference's `modelzoo.<net>.create_model(...)` signatures."""
from . import (pretrained_encoder, deltanet, deltanet_majority_vote, deltanet_v1, lstm_classifier_baseline,
               adenet_v1, adenet_v2, adenet_v3, adenet_3stream, adenet_4stream)
