"""Bin-row stripe partition of one frame across ranks.

The path shards by bin rows (256-px stripes): backdrop only propagates left-to-right inside a tile row, tiles, PTCL
and segments are per tile, so horizontal stripes need NO data-path collective. Each rank renders
`bin_rows = stripe_for(rank, world, height)`; concatenating the stripes in rank order is the full frame.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

BIN_PX = 256  # 16 tiles x 16 px


def n_bin_rows(height: int) -> int:
    return (height + BIN_PX - 1) // BIN_PX


def stripe_for(rank: int, world: int, height: int) -> Tuple[int, int]:
    """Contiguous, as-even-as-possible bin-row range [b0, b1) of `rank`. Ranks beyond the number of bin rows
    get an empty range (b0 == b1)."""
    n = n_bin_rows(height)
    base, rem = divmod(n, world)
    b0 = rank * base + min(rank, rem)
    b1 = b0 + base + (1 if rank < rem else 0)
    return b0, b1


def stripe_pixel_rows(bin_rows: Tuple[int, int], height: int) -> Tuple[int, int]:
    return min(bin_rows[0] * BIN_PX, height), min(bin_rows[1] * BIN_PX, height)


def assemble(stripes: List[np.ndarray]) -> np.ndarray:
    """Concatenate per-rank stripes (rank order) into the frame; empty stripes are skipped."""
    parts = [s for s in stripes if s.shape[0] > 0]
    return np.concatenate(parts, axis=0)
