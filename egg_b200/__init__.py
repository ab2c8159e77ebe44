"""egg_b200 — B200-native photometry engine for egg-gencat.

The package is a thin host mirror over the C-ABI CUDA library ``libeggphot.so``
(``include/eggphot.h``).  Nothing here computes fluxes on the CPU.
"""
from .photometry import EggphotError, Photometry, load_library  # noqa: F401
