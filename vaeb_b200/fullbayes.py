"""`VAEBfullbayes.py` of the reference: class `VAE` (VAEBfullbayes.py:13-201) -- a plain VAE
whose objective is the MEAN bound (`T.mean(KL + logpXgivenZ)`, :142) without the weight prior
and whose Adagrad step carries an extra `- lr*1e-6*p**2` (:183-184) -- plus its script body
(:203-244).  Same kernels, different scalars (VAEB_VARIANT_FULLBAYES)."""
from __future__ import annotations

import time

import numpy as np

from .data import load_frey, load_mnist
from .model import VAEB


class VAE(VAEB):
    def __init__(self, x_train, continuous=False, hidden_units=500, latent_size=10,
                 batch_size=100, L=1, learning_rate=0.01, **ext):
        # L is stored but the graph draws a single eps (VAEBfullbayes.py:129-133)
        VAEB.__init__(self, x_train, continuous, hidden_units, latent_size, batch_size, 1, learning_rate,
                      False, False, variant="fullbayes", **ext)
        self.L = L


def main(n_epochs=2000, continuous=True, n_latent=10, synthetic=False, **ext):
    """The script body of VAEBfullbayes.py:203-244: Frey Face (200 hidden units) or MNIST (500), a fixed seed of
    10, one printed pair of lines per epoch; `validate` already returns a mean here (:243)."""
    np.random.seed(10)
    print("loading data")
    loader, hidden = (load_frey, 200) if continuous else (load_mnist, 500)
    x_train, x_valid = loader(synthetic=synthetic)
    print("creating the model")
    model = VAE(x_train, continuous, hidden, n_latent, **ext)
    print("learning")
    batch_order = np.arange(model.N // model.batch_size)
    LB = LBvalidation = float("nan")
    for epoch in range(1, n_epochs + 1):
        t_epoch = time.time()
        np.random.shuffle(batch_order)
        LB = float(np.sum(model.update_many(batch_order), dtype=np.float64)) / len(batch_order)
        print("Epoch %s : [Lower bound: %s, time: %s]" % (epoch, LB, time.time() - t_epoch))
        LBvalidation = float(model.validate(x_valid))
        print("          [Lower bound on validation set: %s]" % LBvalidation)
    return model, LB, LBvalidation
