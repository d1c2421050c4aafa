"""aesmc_b200 -- B200-native implementation of the SMC core of tuananhle7/aesmc.

Same module layout and call signatures as the reference package (`losses`, `inference`,
`statistics`, `train`, plus `state` and `math`), so user-supplied torch.nn models are drop-in; the
per-time-step hot path runs in hand-written sm_100a CUDA kernels behind a C ABI
(include/aesmc_b200.h, aesmc_b200/libaesmc_b200.so).  There is no CPU fallback.

``install_as_aesmc()`` registers this package under the name ``aesmc`` so that code written against
the reference (`import aesmc; aesmc.state.BatchShapeMode...`) runs unchanged.
"""
import sys

__version__ = "0.1.0"

from . import state  # noqa: E402
from . import math  # noqa: E402
from . import losses  # noqa: E402
from . import inference  # noqa: E402
from . import statistics  # noqa: E402
from . import train  # noqa: E402
from . import distributed  # noqa: E402
from . import fused  # noqa: E402
from ._ops import get_resampling_mode, set_resampling_mode  # noqa: E402,F401


def install_as_aesmc(force=False):
    """Alias this package (and its submodules) as ``aesmc`` in sys.modules."""
    if "aesmc" in sys.modules and sys.modules["aesmc"] is not sys.modules[__name__] and not force:
        raise RuntimeError("a different `aesmc` package is already imported; pass force=True to replace it")
    me = sys.modules[__name__]
    sys.modules["aesmc"] = me
    for sub in ("state", "math", "losses", "inference", "statistics", "train", "distributed"):
        sys.modules["aesmc." + sub] = getattr(me, sub)
    return me
