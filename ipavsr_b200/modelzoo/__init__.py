"""Model builders with the reference's `modelzoo.<net>.create_model(...)` signatures."""
from . import (pretrained_encoder, autoencoder, deltanet, deltanet_majority_vote, deltanet_v1, lstm_classifier_baseline,
               lstm_classifier_majority_vote, adenet_v1, adenet_v1_1, adenet_v2, adenet_v2_1, adenet_v2_2, adenet_v2_3,
               adenet_v2_4, adenet_v2_nodelta, adenet_v3, adenet_v4, adenet_v5, adenet_v6, adenet_2stream, adenet_3stream,
               adenet_3stream_dct, adenet_3stream_dropout, adenet_4stream, avnet)
