"""Input streams the device computes from another stream of the same call.

In the reference the diff-image stream and the DCT stream are functions of the raw mouth-ROI stream, computed once on the
host before training: `presplit_dataprocessing` applies `compute_diff_images(data_matrix, vidlens)` when a stream's ini
section says `diffimage = true` (`runners/3stream.py:85-99`, `utils/preprocessing.py:506-517`), and the DCT features are
`compute_dct_features(X, (30, 40), 30, method='zigzag')` followed by `concat_first_second_deltas(dct_feats, vidlens)`
(`avletters/preprocess_images.py:20-21`, `utils/preprocessing.py:417-462, 465-489`), stored as `dctFeatures` and read back
`.astype('float32')` (`avletters/bimodal.py:351`).  A runner then uploads all three padded streams every step.

Passing one of these objects in place of a stream's array makes the engine compute that stream on the device from the
frames of the source stream it has already staged (kernels `ipavsr_diff_image`, `ipavsr_dct_project`,
`ipavsr_deltas_fir_f32`): only the raw stream crosses the host link.

    train(raw, DiffImages(raw), DctFeatures(raw, (30, 40), 30), targets, mask, window)

`source` is the array object passed for the other stream in the same call (matched by identity).  The utterance lengths
come from the mask, which must be a prefix mask (the reference's generators only produce those, `utils/datagen.py:131,141`).
"""


class Derived(object):
    def __init__(self, source):
        self.source = source


class DiffImages(Derived):
    """`compute_diff_images(source, lens)` (`utils/preprocessing.py:506-517`): frame differences per utterance, the first
    frame duplicating the first difference."""


class DctFeatures(Derived):
    """`compute_dct_features(source, image_shape, no_coeff, 'zigzag')` (`utils/preprocessing.py:417-462`), by default
    followed by `concat_first_second_deltas(., lens)` with the reference's window of 9 (`:465-489`, `deltas` `:17-51`)
    and the float32 rounding of the runners: F = 3 * no_coeff features (no_coeff without deltas)."""

    def __init__(self, source, image_shape, no_coeff=30, deltas=True, window=9):
        Derived.__init__(self, source)
        self.image_shape = (int(image_shape[0]), int(image_shape[1]))
        self.no_coeff, self.deltas, self.window = int(no_coeff), bool(deltas), int(window)

    @property
    def width(self):
        return self.no_coeff * (3 if self.deltas else 1)
