"""vaeb_b200 -- B200-native AEVB training/evaluation hot path behind the reference's
`VAEB.py` surface.  `from vaeb_b200 import VAEB` needs the CUDA library
(`python -m vaeb_b200.build`); there is no CPU fallback."""
from .model import VAEB, SharedParam, comm_unique_id, mlp_forward, pinned_empty  # noqa: F401
from .cli import train_model, parse_args, print_args, command_line_args, command_line_flags  # noqa: F401

__all__ = ["VAEB", "SharedParam", "train_model", "parse_args", "print_args", "command_line_args",
           "command_line_flags", "mlp_forward", "comm_unique_id", "pinned_empty"]
