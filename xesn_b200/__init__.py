"""xesn_b200 -- B200-native implementation of the xesn ESN hot path.

Same public names as the reference package for this path (`xesn/__init__.py:13-22`):
`ESN`, `LazyESN`, `RandomMatrix`, `SparseRandomMatrix`; `ESNEnsemble` batches the candidates of the
reference's macro training (cost.py:190-203).  The numerical layer is the
sm_100a CUDA library `libxesn_b200.so` (C ABI in include/xesn_b200.h), loaded on the
first numerical call; there is no CPU fallback.
"""
from .matrix import RandomMatrix, SparseRandomMatrix
from .esn import ESN
from .lazyesn import LazyESN
from .ensemble import ESNEnsemble
from .cost import CostFunction
from .io import from_zarr, from_npz, to_npz
from .psd import psd
from ._lib import XesnError, LIB_PATH, launch_count

__all__ = ["ESN", "LazyESN", "ESNEnsemble", "CostFunction", "RandomMatrix", "SparseRandomMatrix", "from_zarr", "from_npz", "to_npz",
           "psd", "XesnError", "LIB_PATH", "launch_count"]
__version__ = "0.1.0"
