"""Host mirrors of the reference's `utils` package for the hot path: `preprocessing` (normalisation, deltas, DCT features,
reorder, force-align — executed by csrc/preprocess.cu and csrc/features.cu), `signal` (append_delta_coeff), `datagen`
(device-side batch assembly), `evaluate` (on-device vote / confusion matrix) and `io` (pickle / .mat layouts)."""
