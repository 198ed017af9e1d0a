"""golemflavor_b200 -- B200-native hot path of GolemFlavor.

Batched log-posterior evaluation (mixing angles -> PMNS -> effective Hamiltonian ->
3x3 Hermitian eigendecomposition -> averaged transition matrix -> measured flavor
ratio -> Gaussian flavor-ratio likelihood + prior) behind the reference's Python
function signatures (``golemflavor/fr.py``, ``golemflavor/llh.py``,
``golemflavor/mcmc.py``).  All arithmetic runs in hand-written sm_100a CUDA kernels
reached through the C ABI of ``include/golemflavor_b200.h``; there is no CPU
implementation -- without the built library or without a CUDA device every compute
call raises.
"""

__version__ = '0.1.0'

from . import enums, param  # noqa: F401  (host-side containers, no device needed)
