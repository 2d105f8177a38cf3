"""Debug aid: ipavsr_b200.utils.preprocessing.featurewise_normalize_sequence against the golden vectors."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ipavsr_b200.utils import preprocessing as P
G = np.load(os.path.join(ROOT, 'tests', 'golden', 'preprocessing.npz'))
X = G['X']
for i in range(3):
    n, m, s = P.featurewise_normalize_sequence(X)
    print(i, 'n err', np.abs(n - G['featurewise_norm']).max(), 'm err', np.abs(m - G['featurewise_mean']).max(),
          's err', np.abs(s - G['featurewise_std']).max(), m[:4], s[:4])
print(P.normalize_input(X.copy())[:1, :4])
n, m, s = P.featurewise_normalize_sequence(X)
print('after normalize_input: m err', np.abs(m - G['featurewise_mean']).max())
