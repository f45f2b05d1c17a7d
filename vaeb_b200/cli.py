"""The args/flags-dictionary CLI and the training driver of the reference (VAEB.py:25-38,
471-612), Python 3, over the B200 model class.  Same argument names (including the
`full_varational` spelling), defaults, printed blocks, `.trc` layout and `.mdl` layout."""
from __future__ import annotations

import sys
import time

import numpy as np

from . import io
from .data import load_frey, load_mnist
from .model import VAEB

# name -> (default, type): the dictionary-driven CLI of VAEB.py:22-38; a new option is one more entry here
command_line_args = {'seed': (15485863, int),
                     'n_latent': (10, int),
                     'n_epochs': (2000, int),
                     'batch_size': (100, int),
                     'L': (1, int),
                     'hidden_unit': (-1, int),
                     'learning_rate': (0.01, float),
                     'trace_file': ('', str),
                     'save_file': ('', str),
                     'load_file': ('', str),
                     'vb_param_file': ('', str),
                     # --- extensions (defaults keep the reference behaviour) ---
                     'device': (0, int),
                     'precision': ('fp32', str),      # fp32 | bf16
                     'eps_mode': ('philox', str),     # philox | theano
                     'manifold_file': ('', str)}      # PGM of the learned 2-d manifold (freyFace.py:346-369)
command_line_flags = ['continuous', 'generic_estimator', 'full_varational',
                      # --- extensions ---
                      'synthetic',        # synthetic data of the dataset's shape if the files are absent
                      'sample_weights']   # full-VB with sampled weights (VAEB.py:127-129 live)


def _pop_option(tokens, name, n_values):
    """Removes the first `--name` from `tokens` together with its `n_values` values; returns the values (a list)
    or None when the option is absent.  The reference's parser (VAEB.py:471-489) behaves the same way: first
    occurrence wins, later duplicates stay behind and end up in the 'unused' report."""
    key = "--" + name
    try:
        at = tokens.index(key)
    except ValueError:
        return None
    values = tokens[at + 1:at + 1 + n_values]
    if len(values) < n_values:                       # `--key` as the last token: the reference raises IndexError here
        raise IndexError("missing value for %s" % key)
    tokens[at:at + 1 + n_values] = []
    return values


def get_arg(arg, args, default, type_):
    """`--arg value` cast with the dictionary's type, else the default (VAEB.py:471-480)."""
    found = _pop_option(args, arg, 1)
    return default if found is None else type_(found[0])


def get_flag(flag, args):
    """True iff `--flag` was given (VAEB.py:483-489)."""
    return _pop_option(args, flag, 0) is not None


def parse_args(argv=None):
    """VAEB.py:491-504: every key of the two dictionaries ends up in the result; what is left over is reported."""
    tokens = list(sys.argv[1:] if argv is None else argv)
    parsed = {name: get_arg(name, tokens, default, cast) for name, (default, cast) in command_line_args.items()}
    parsed.update((name, get_flag(name, tokens)) for name in command_line_flags)
    if tokens:
        print('Have unused args: {0}'.format(tokens))
    return parsed


def print_args(args):
    """VAEB.py:507-512."""
    rule = '-' * 38
    body = ['\t{0}: {1}'.format(key, args[key]) for key in args]
    print('\n'.join(['Parameters used:', rule] + body + [rule]))


def _default_hidden(continuous, hidden_unit):
    # VAEB.py:542-543,551-552: 200 units for Frey Face, 500 for MNIST unless --hidden_unit says otherwise
    return hidden_unit if hidden_unit >= 0 else (200 if continuous else 500)


def train_model(args, data=None):
    """The training driver of VAEB.py:524-598: seed numpy's global RNG (it only drives the shuffle), load the
    dataset of the chosen mode, build the model (on top of --vb_param_file for full VB), then per epoch: shuffle
    the minibatch order, update on every minibatch, validate, write the trace line (twice, as the reference does)
    and print the two progress lines; finally save."""
    np.random.seed(args['seed'])
    continuous = args['continuous']
    synthetic = args.get('synthetic', False)
    ext = dict(device=args.get('device', 0), precision=args.get('precision', 'fp32'),
               eps_mode=args.get('eps_mode', 'philox'))
    trace_file, save_file = args['trace_file'], args['save_file']

    print("loading data")
    if data is None:
        data = load_frey(synthetic=synthetic) if continuous else load_mnist(synthetic=synthetic)
    x_train, x_valid = data[0], data[1]

    print("creating the model")
    params = None
    if args['full_varational']:                        # VAEB.py:559-561: the MAP solution seeds the variational means
        seed_model, _ = VAEB.load(args['vb_param_file'], data=data, **ext)
        params = seed_model.get_params()
        seed_model.close()
    model = VAEB(x_train, continuous, _default_hidden(continuous, args['hidden_unit']), args['n_latent'],
                 args['batch_size'], args['L'], args['learning_rate'], args['generic_estimator'],
                 args['full_varational'], params, sample_weights=args.get('sample_weights', False), **ext)

    print("learning")
    tracing = len(trace_file) > 0
    if tracing:
        io.trace_header(trace_file)
    batch_order = np.arange(model.N // model.batch_size)           # the remainder rows are never visited (:571)
    for epoch in range(args['n_epochs']):
        t_epoch = time.time()
        np.random.shuffle(batch_order)
        # the reference accumulates `model.update(batch)` over the order (VAEB.py:577-579); update_many performs
        # the same updates in the same order without a host round trip per minibatch
        LB = float(np.sum(model.update_many(batch_order), dtype=np.float64)) / len(batch_order)
        LBvalidation = float(model.validate(x_valid)) / x_valid.shape[0]
        seen = model.N * (epoch + 1)
        if tracing:
            io.trace_line(trace_file, seen, LB, LBvalidation)
        print("Epoch %s : [Lower bound: %s, time: %s]" % (epoch, io.py2_float(LB), time.time() - t_epoch))
        print("          [Lower bound on validation set: %s]" % io.py2_float(LBvalidation))
        if tracing:                                                # every line appears twice (VAEB.py:591-593)
            io.trace_line(trace_file, seen, LB, LBvalidation)

    if len(save_file) > 0:
        model.save(save_file)

    # freyFace.py:346-369 renders the learned manifold after training: the decoder on a 10 x 10 grid of Gaussian
    # quantiles, tiled into one image (VAEBImage.multipleImages)
    if len(args.get('manifold_file', '')) > 0:
        from . import manifold
        _, tiled = manifold.render(model)
        manifold.save_pgm(tiled, args['manifold_file'])

    return model, data


def main(argv=None):
    args = parse_args(argv)
    print_args(args)
    if len(args['load_file']) == 0:
        model, data = train_model(args)
    else:
        data = None
        if args.get('synthetic', False):
            data = load_frey(synthetic=True) if args['continuous'] else load_mnist(synthetic=True)
        model, data = VAEB.load(args['load_file'], data=data, device=args['device'], precision=args['precision'],
                                eps_mode=args['eps_mode'])
    return model, data
