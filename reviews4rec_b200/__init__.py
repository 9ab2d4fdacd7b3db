"""reviews4rec_b200 -- B200-native rating-prediction training hot path of noveens/reviews4rec.

Python host layer (this package) over hand-written sm_100a CUDA kernels behind a C ABI
(``include/r4r_b200.h`` -> ``libr4r_b200.so``).  The module classes mirror the reference's
``pytorch_models/`` constructors, ``forward(data)`` signatures and ``state_dict`` layout so they drop
into the reference's own ``main.train()`` loop.  There is no CPU fallback: every op raises if the
library is missing or a tensor is not on a CUDA device.
"""
from . import _lib  # noqa: F401  (loads the shared library eagerly; raises if it is absent)
from .loss import MSELoss  # noqa: F401
from .pytorch_models.DeepCoNN import DeepCoNN  # noqa: F401
from .pytorch_models.MF import MF  # noqa: F401
from .pytorch_models.NARRE import NARRE  # noqa: F401
from .pytorch_models.NeuMF import GMF, MLP, NeuMF  # noqa: F401
from .pytorch_models.TransNet import TransNet  # noqa: F401

__all__ = ["DeepCoNN", "MF", "NARRE", "TransNet", "GMF", "MLP", "NeuMF", "MSELoss"]
