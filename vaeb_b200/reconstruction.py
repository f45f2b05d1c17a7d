"""Reconstruction / MSE evaluation of trained models (reconstruction.py:9-61 of the reference): load a `.mdl`,
reconstruct the test rows from the posterior mean (n = 0) or from an average over n = 20 posterior samples, report
the mean squared reconstruction error and dump a few original / reconstruction image pairs.

The reference calls `model.reconstruct` once per test row for the Gaussian decoder (reconstruction.py:14-18); here
the whole test set is ONE `vaeb_reconstruct` call -- the same estimator, one launch sequence
instead of N_test Theano compilations."""
from __future__ import annotations

import numpy as np

from . import manifold

size_continuous_latent_space = [2, 10, 20]
model_file = 'reconstruction_res/{0}_{1}.mdl'
log_file = 'reconstruction_res/MSE.res'


def MSE(model, x_test, num_samples, eps=None):
    """mean_i || reconstruct(x_i) - x_i ||^2  (reconstruction.py:9-18)."""
    x = np.asarray(x_test, np.float32).reshape(-1, model.input_size)
    rec = model.reconstruct(x, num_samples, eps=eps)
    d = np.asarray(rec, np.float64) - x
    return float(np.mean(np.sum(d * d, axis=1)))


def reconstruction_test(x_test, model, file_prefix, continuous, log=log_file, n_examples=8, image_ext="jpg"):
    """reconstruction.py:20-42: for n in (0, 20): a few original / reconstruction image pairs, then one line
    `data_type,latent_size,sample_type,MSE` appended to the log.  Returns {n: mse}."""
    out = {}
    kind = 'continuous' if continuous else 'discrete'
    for num_samples in (0, 20):
        print('num_samples :\n{0}'.format(num_samples))
        k = min(n_examples, len(x_test))
        samples = model.reconstruct(np.asarray(x_test[:k], np.float32), num_samples)
        for i in range(k):
            stem = file_prefix + '_image_{0}_{1}_'.format(num_samples, i)
            manifold.save_image(x_test[i], stem + 'original.' + image_ext)
            manifold.save_image(samples[i], stem + 'sample.' + image_ext)
        mse = MSE(model, x_test, num_samples)
        out[num_samples] = mse
        if log:
            with open(log, 'a') as f:
                f.write('{0},{1},{2},{3}\n'.format(kind, model.n_latent, 'mean' if num_samples == 0 else 'sample', mse))
    return out


def main(data_types=('discrete', 'continuous'), sizes=tuple(size_continuous_latent_space), synthetic=False, **ext):
    """reconstruction.py:45-58: every shipped model file, both decoders."""
    from .model import VAEB
    from .data import load_frey, load_mnist
    with open(log_file, 'w') as f:
        f.write('data_type,latent_size,sample_type,MSE\n')
    results = {}
    for data_type in data_types:
        for s in sizes:
            data = None
            if synthetic:
                data = load_frey(synthetic=True) if data_type == 'continuous' else load_mnist(synthetic=True)
            model, data = VAEB.load(model_file.format(data_type, s), data=data, **ext)
            x_test = data[1] if model.continuous else (data[2] if len(data) > 2 else data[1])
            if isinstance(x_test, (tuple, list)):          # the reference's MNIST triple carries (x, y) pairs
                x_test = x_test[0]
            results[(data_type, s)] = reconstruction_test(
                x_test, model, "reconstruction_res/{0}_{1}_".format(data_type, s), model.continuous)
            model.close()
    return results


if __name__ == '__main__':
    main()
