"""rpgp -- B200-native matrix-free K.V for randomly-projected additive GPs (host side of librpgp.so)."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
