"""Dataset access for the CLI.  The reference reads `freyfaces.pkl` / `mnist.pkl.gz` from the
working directory (VAEB.py:541-556); both are absent from the reference checkout
(.MISSING_LARGE_BLOBS), so synthetic data of the same shapes is offered behind
`--synthetic`."""
from __future__ import annotations

import gzip
import os
import pickle

import numpy as np


def synthetic_mnist(n, seed=15485863, D=784):
    """x in [0,1]^{n x 784}, ~19% non-zero pixels, non-binarised like mnist.pkl.gz."""
    rng = np.random.RandomState(seed)
    x = rng.uniform(size=(n, D)).astype(np.float32)
    x *= (rng.uniform(size=(n, D)) < 0.19)
    return x


def synthetic_frey(n=1965, seed=15485863, D=560):
    rng = np.random.RandomState(seed)
    return np.clip(0.5 + 0.2 * rng.normal(size=(n, D)), 0, 1).astype(np.float32)


def load_frey(path="freyfaces.pkl", synthetic=False):
    """VAEB.py:544-549: first 1500 rows train, the rest validation."""
    if os.path.exists(path):
        with open(path, "rb") as f:
            data = np.asarray(pickle.load(f, encoding="latin1"), dtype=np.float32)
    elif synthetic:
        data = synthetic_frey()
    else:
        raise IOError("%s not found (pass --synthetic for Frey-shaped synthetic data)" % path)
    return data[:1500], data[1500:]


def load_mnist(path="mnist.pkl.gz", synthetic=False, n_train=50000, n_valid=10000):
    """VAEB.py:553-555: 50000/10000/10000 split; only x_train and x_valid are used."""
    if os.path.exists(path):
        with gzip.open(path, "rb") as f:
            (x_train, _), (x_valid, _), _ = pickle.load(f, encoding="latin1")
        return np.asarray(x_train, np.float32), np.asarray(x_valid, np.float32)
    if synthetic:
        x = synthetic_mnist(n_train + n_valid)
        return x[:n_train], x[n_train:]
    raise IOError("%s not found (pass --synthetic for MNIST-shaped synthetic data)" % path)
