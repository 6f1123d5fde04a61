"""surface-remesher_b200 — B200-native discrete-CVT Lloyd engine behind Surface-Remesher's
`gCVT` / `discretization_d` entry points.

Host-side mirror of the reference interface (source/gcvt.h, source/discretization.h) over the
C ABI of libsrm.so (include/srm.h).  The CUDA library is the product; this module only marshals
numpy / torch buffers.  There is no CPU fallback: if libsrm.so is missing or no CUDA device is
present, every compute call raises.
"""
from .api import (  # noqa: F401
    MARKER,
    Context,
    SrmError,
    centroidalVoronoi,
    discretization_d,
    gCVT,
    generateMask,
    lib,
    lib_path,
    locate,
    putConstrains,
    randomPoints,
    recover,
    row_bands,
    row_bands_balanced,
    rebalance_bands,
)
from .batch import BatchLloyd, shard_meshes  # noqa: F401

__all__ = [
    "MARKER", "Context", "SrmError", "centroidalVoronoi", "discretization_d", "gCVT", "generateMask",
    "lib", "lib_path", "locate", "putConstrains", "randomPoints", "recover", "row_bands", "row_bands_balanced", "rebalance_bands",
    "BatchLloyd", "shard_meshes",
]
