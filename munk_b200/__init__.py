"""munk_b200 — MUNK's dense linear-algebra hot path on NVIDIA B200 (sm_100a).

Drop-in for the reference `munk` package (munk/munk/__init__.py:1-11): same exported names.
The CUDA library (munk_b200/lib/libmunk_b200.so, C ABI in include/munk_b200.h) must be built
(`python -m munk_b200.build`); importing the package without it fails — there is no CPU path.
"""
import sys as _sys

# `python -m munk_b200.build` in a clean checkout imports this file before the library exists:
# only then is the API (which loads the library, and fails without it) left out.
if "munk_b200.build" not in getattr(_sys, "orig_argv", ()):
    from . import io, util                                                  # noqa: F401
    from .munk import regularized_laplacian, rkhs_factor                    # noqa: F401  kernels and factor
    from .munk import embed_matrices, embed_networks, save_embeddings       # noqa: F401  embeddings
    from .munk import scores, separate_scores, release_device_cache         # noqa: F401  scores

__version__ = '0.1'
