#!/usr/bin/env python
"""Drop-in for the reference's `VAEB.py` entry point: same CLI (VAEB.py:25-38,471-612), same
module-level names (`VAEB`, `train_model`, `parse_args`, ...), B200 kernels underneath."""
from vaeb_b200.cli import (command_line_args, command_line_flags, get_arg, get_flag, main,  # noqa: F401
                           parse_args, print_args, train_model)
from vaeb_b200.model import VAEB  # noqa: F401

if __name__ == '__main__':
    main()
