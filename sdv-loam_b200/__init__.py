"""sdv-loam_b200 — B200-native (sm_100a CUDA) implementation of SDV-LOAM's tracking / optimisation hot path.

The product is the C-ABI shared library `libsdv_b200.so` (include/sdv_b200.h); this package is the thin Python host
mirror of the reference's call surface (CoarseTracker / EnergyFunctional) used by tests and bench.  It never falls back to
a CPU implementation: importing `api` fails loudly if the CUDA library is missing.

The directory name contains a hyphen, so import it through the root-level shim:  `import sdv_loam_b200`.
"""
from .build import build_library, library_path  # noqa: F401
