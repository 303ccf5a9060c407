"""Histogram helper of the reference's KS tests (tests/helpers/plotting.py:12-15)."""

import numpy as np


def make_norm_histo(array, range_, weights=None):
    histo = np.histogram(array, 100, range=range_, weights=weights)[0]
    return histo / np.sum(histo)
