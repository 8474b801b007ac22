"""`RenderConfig` / `ConfigUniform` / MSAA mask LUTs, restated.

* ConfigUniform (22 x u32)        -- vello_encoding/src/config.rs:120-154, shader/shared/config.wgsl:5-42
* RenderConfig::new               -- vello_encoding/src/config.rs:168-196
* WorkgroupCounts (dispatch sizes)-- vello_encoding/src/config.rs:228-273
* make_mask_lut / make_mask_lut_16-- vello_encoding/src/mask.rs:10-98

Unlike the reference (config.rs:398-408: hand-picked `1 << 21` constants) the bump-allocated
buffer capacities are NOT part of this config: the CUDA renderer sizes its arenas itself
(count passes + grow-and-retry) and writes the capacities it actually used into the uniform.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .encoding import Layout, Color, BLACK

TILE_WIDTH = 16
TILE_HEIGHT = 16
N_TILE_X = 16
N_TILE_Y = 16

AA_AREA, AA_MSAA8, AA_MSAA16 = 0, 1, 2


@dataclass
class RenderParams:
    """`vello::RenderParams` (vello/src/lib.rs:357-369)."""

    base_color: Color = BLACK
    width: int = 0
    height: int = 0
    antialiasing_method: int = AA_AREA


def tiles_for(width: int, height: int):
    wt = (width + TILE_WIDTH - 1) // TILE_WIDTH
    ht = (height + TILE_HEIGHT - 1) // TILE_HEIGHT
    return wt, ht


def _one_mask(slope: float, translation: float, is_pos: bool, pattern, n: int) -> int:
    if is_pos:
        translation = 1.0 - translation
    result = 0
    inv = 1.0 / n
    for i, item in enumerate(pattern):
        y = (i + 0.5) * inv
        x = (item + 0.5) * inv
        if not is_pos:
            y = 1.0 - y
        if (x - (1.0 - translation)) * (1.0 - slope) - (y - translation) * slope >= 0.0:
            result |= 1 << i
    return result


_PATTERN_8 = [0, 5, 3, 7, 1, 4, 6, 2]
_PATTERN_16 = [1, 8, 4, 11, 15, 7, 3, 12, 0, 9, 5, 13, 2, 10, 6, 14]


def make_mask_lut() -> np.ndarray:
    """32x32 u8 half-plane masks, returned as 256 u32 words (mask.rs:36-48)."""
    W = H = 32
    out = np.zeros(W * H, dtype=np.uint8)
    for i in range(W * H):
        u, v = i % W, i // W
        is_pos = v >= H // 2
        y = ((v % (H // 2)) + 0.5) * (1.0 / (H // 2))
        x = (u + 0.5) * (1.0 / W)
        out[i] = _one_mask(y, x, is_pos, _PATTERN_8, 8)
    return out.view(np.uint32).copy()


def make_mask_lut_16() -> np.ndarray:
    """64x64 u16 half-plane masks, returned as 2048 u32 words (mask.rs:83-98)."""
    W = H = 64
    out = np.zeros(W * H, dtype=np.uint16)
    for i in range(W * H):
        u, v = i % W, i // W
        is_pos = v >= H // 2
        y = ((v % (H // 2)) + 0.5) * (1.0 / (H // 2))
        x = (u + 0.5) * (1.0 / W)
        out[i] = _one_mask(y, x, is_pos, _PATTERN_16, 16)
    return out.view(np.uint32).copy()
