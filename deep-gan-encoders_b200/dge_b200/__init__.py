"""dge_b200 -- runtime binding of the B200-native GAN-inversion hot path (see include/dge_b200.h)."""
from . import _lib  # noqa: F401
from ._lib import DgeError  # noqa: F401
