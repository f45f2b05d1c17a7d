"""The args/flags-dictionary CLI and the training driver of the reference (VAEB.py:25-38,
471-612), Python 3, over the B200 model class.  Same argument names (including the
`full_varational` spelling), defaults, printed blocks, `.trc` layout and `.mdl` layout."""
from __future__ import annotations

import copy
import sys
import time

import numpy as np

from . import io
from .data import load_frey, load_mnist
from .model import VAEB

#   to add another command line argument, add its name as a key and a tuple of its default
#   value and type (VAEB.py:22-36)
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


def get_arg(arg, args, default, type_):
    arg = '--' + arg
    if arg in args:
        index = args.index(arg)
        value = args[args.index(arg) + 1]
        del args[index]     # remove arg-name
        del args[index]     # remove value
        return type_(value)
    else:
        return default


def get_flag(flag, args):
    flag = '--' + flag
    have_flag = flag in args
    if have_flag:
        args.remove(flag)
    return have_flag


def parse_args(argv=None):
    args = copy.deepcopy(sys.argv[1:] if argv is None else list(argv))
    arg_dict = {}
    for (arg_name, arg_args) in command_line_args.items():
        (arg_default_val, arg_type) = arg_args
        arg_dict[arg_name] = get_arg(arg_name, args, arg_default_val, arg_type)
    for flag_name in command_line_flags:
        arg_dict[flag_name] = get_flag(flag_name, args)
    if len(args) > 0:
        print('Have unused args: {0}'.format(args))
    return arg_dict


def print_args(args):
    print('Parameters used:')
    print('--------------------------------------')
    for (k, v) in args.items():
        print('\t{0}: {1}'.format(k, v))
    print('--------------------------------------')


def train_model(args, data=None):
    """VAEB.py:524-598."""
    np.random.seed(args['seed'])
    n_latent = args['n_latent']
    n_epochs = args['n_epochs']
    continuous = args['continuous']
    batch_size = args['batch_size']
    L = args['L']
    hidden_unit = args['hidden_unit']
    learning_rate = args['learning_rate']
    trace_file = args['trace_file']
    generic_estimator = args['generic_estimator']
    full_varational = args['full_varational']
    save_file = args['save_file']
    vb_param_file = args['vb_param_file']
    ext = dict(device=args.get('device', 0), precision=args.get('precision', 'fp32'),
               eps_mode=args.get('eps_mode', 'philox'))

    print("loading data")
    if continuous:
        if hidden_unit < 0:
            hidden_unit = 200
        if data is None:
            data = load_frey(synthetic=args.get('synthetic', False))
    else:
        if hidden_unit < 0:
            hidden_unit = 500
        if data is None:
            data = load_mnist(synthetic=args.get('synthetic', False))
    x_train, x_valid = data[0], data[1]

    print("creating the model")
    if full_varational:
        model, tmp = VAEB.load(vb_param_file, data=data, **ext)
        params = model.get_params()
        model.close()
    else:
        params = None

    model = VAEB(x_train, continuous, hidden_unit, n_latent, batch_size, L, learning_rate, generic_estimator,
                 full_varational, params, sample_weights=args.get('sample_weights', False), **ext)

    print("learning")
    if len(trace_file) > 0:
        io.trace_header(trace_file)
    batch_order = np.arange(int(model.N / model.batch_size))  # ordering of the batches
    for epoch in range(n_epochs):
        start = time.time()
        np.random.shuffle(batch_order)
        # the reference loops `LB += model.update(batch)` (VAEB.py:577-579); update_many runs
        # the same updates in the same order without a host round-trip per minibatch
        LB = float(np.sum(model.update_many(batch_order), dtype=np.float64))
        LB /= len(batch_order)
        LBvalidation = float(model.validate(x_valid)) / x_valid.shape[0]
        if len(trace_file) > 0:
            io.trace_line(trace_file, model.N * (epoch + 1), LB, LBvalidation)

        print("Epoch %s : [Lower bound: %s, time: %s]" % (epoch, io.py2_float(LB), time.time() - start))
        print("          [Lower bound on validation set: %s]" % io.py2_float(LBvalidation))

        if len(trace_file) > 0:   # the reference writes every line twice (VAEB.py:591-593)
            io.trace_line(trace_file, model.N * (epoch + 1), LB, LBvalidation)

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
