"""phlash_b200: B200-native (sm_100a) drop-in for the PSMC coalescent-HMM log-likelihood /
gradient hot path of jthlab/phlash (reference: src/phlash/gpu.py, kernel.py, hmm.py).

Only what that path needs lives here: ``csrc/`` (CUDA kernels + the C ABI declared in
``include/phlash_b200.h``) and the host-side mirror of the reference's kernel interface
(``gpu.PSMCKernel``, ``kernel.get_kernel``, ``params.PSMCParams``).  There is no CPU fallback:
without the compiled CUDA library (or without a GPU) every evaluation raises.
"""

from phlash_b200.params import PSMCParams  # noqa: F401

__all__ = ["PSMCParams"]
