"""stringwars_b200 — B200-native batched edit-distance engine behind the StringWars `similarities` surface.

The package is a thin host layer over one native CUDA library (sm_100a) reached through a C ABI
(include/stringwars_b200.h). Importing the package does not load the library; the first engine call does,
and fails loudly if it is missing or no GPU is present — there is no CPU fallback.
"""
from ._lib import Error, LIB_PATH  # noqa: F401
from .szs import (  # noqa: F401
    ALGO_AUTO, ALGO_BITPARALLEL, ALGO_WAVEFRONT, GAP_BIO, GAP_STRINGZILLA, TRANSPORT_COPY_ENGINE, TRANSPORT_NCCL,
    DeviceScope, DeviceTape, LevenshteinDistances, LevenshteinDistancesUtf8, NeedlemanWunschScores, SmithWatermanScores,
    device_count, launch_count, multiprocessor_count, unary_class_costs,
)
from ._lib import RESULT_16, RESULT_32, RESULT_64, TapeView  # noqa: F401
from .tape import BytesTape, ManagedBuffer, PinnedBuffer, unified_mat  # noqa: F401

__version__ = "0.2.0"
