"""`model/stylegan1/lreq.py` is an identical copy of `model/utils/lreq.py` in the reference; re-export."""
from model.utils.lreq import *  # noqa: F401,F403
from model.utils.lreq import Bool, Conv2d, ConvTranspose2d, Linear, use_implicit_lreq  # noqa: F401
