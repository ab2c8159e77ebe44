"""Galaxy sharding across GPUs: galaxies are independent, so each rank takes a
contiguous range of the catalog and no data-path collective is needed; only the small bookkeeping
(counts, a checksum of checksums) is reduced."""
from __future__ import annotations

import numpy as np

GALAXY_COLUMNS = ("z", "d", "m_disk", "m_bulge", "m2lcor", "opt_sed_disk", "opt_sed_bulge", "ir_sed", "lir", "fpah",
                  "mdust", "igm_abs", "vdisp", "line_lum_disk", "line_lum_bulge")


def shard_range(ngal: int, rank: int, world: int) -> "tuple[int, int]":
    """[first, last) of rank's contiguous share; shares differ by at most one galaxy."""
    base, rem = divmod(ngal, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def shard_galaxies(gal: dict, rank: int, world: int) -> dict:
    n = len(gal["z"])
    a, b = shard_range(n, rank, world)
    return {k: (v[a:b] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in gal.items()}


def flux_checksum(flux: np.ndarray) -> np.ndarray:
    """Order-independent fingerprint of a flux block: (count, sum, sum of squares) in float64."""
    f = np.asarray(flux, np.float64)
    return np.array([f.size, f.sum(), (f * f).sum()])
