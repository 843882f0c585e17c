"""B200-native VAR hot path (triplet training step + batched reward query) of
PeixinC/VoiceControlledRobot-VAR.  The package directory name contains a hyphen; import it
through the `var_b200` shim at the repository root or `importlib.import_module`."""
from . import _lib  # noqa: F401  (raises ImportError when libvar_b200.so is not built)
from .engine import ITHOR, KUKA, VarEngine, multistep_lr  # noqa: F401

__all__ = ["VarEngine", "KUKA", "ITHOR", "multistep_lr"]
