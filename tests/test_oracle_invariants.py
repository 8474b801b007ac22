"""Geometric invariants that tie the oracle's INTERMEDIATE buffers to independent definitions (the pixel-level ground truth is
tests/test_oracle_analytic.py). The GPU parity tests compare the CUDA stages with these buffers byte for byte, so what is
checked here is what the stage outputs mean:

* `path_bboxes` (flatten): the integer bounds of the path's own line soup;
* tile `backdrop` (path_count + backdrop_dyn): the winding number of the path at the tile's top-left corner;
* `segments` (path_tiling): the lines cut at tile boundaries -- the total segment length equals the length of the lines inside
  the frame, and every segment lies inside its 16 x 16 tile;
* `lines` are watertight (the reference's debug/validate.rs check)."""
import math

import numpy as np
import pytest

from vello_b200.config import AA_AREA
from vello_b200.encoding import BLACK, FILL_EVEN_ODD, FILL_NON_ZERO, Color, Scene, Stroke, resolve
from vello_b200.shapes import Affine, BezPath, Circle

from . import parity

WHITE = Color.from_rgba8(255, 255, 255)


def _scene(seed, hang_over=False):
    rng = np.random.default_rng(seed)
    s = Scene()
    lo, hi = (-30, 190) if hang_over else (6, 154)
    for k in range(5):
        pts = rng.uniform(lo, hi, (int(rng.integers(4, 10)), 2))  # self-overlapping: winding numbers beyond 0 / 1
        p = BezPath()
        p.move_to(*pts[0])
        for q in pts[1:]:
            p.line_to(*q)
        p.close_path()
        s.fill(FILL_NON_ZERO if k % 2 == 0 else FILL_EVEN_ODD, Affine.IDENTITY, Color.from_rgba8(40 * k + 30, 200 - 30 * k, 90, 255), None, p)
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, WHITE, None, Circle(80.0, 70.0, 41.5))
    s.stroke(Stroke(7.0), Affine.IDENTITY, WHITE, None, BezPath([("M", 20.0, 30.0), ("C", 60.0, 150.0, 110.0, -20.0, 150.0, 120.0)]))
    return s


def _run(oracle, scene, w, h):
    oracle.render(resolve(scene.encoding), w, h, BLACK.premul_rgba8_u32(), AA_AREA)
    bump = oracle.buffer("bump")[0]
    lines = oracle.buffer("lines")[: int(bump["lines"])]
    return bump, lines


def _winding(lines, px, py):
    """Crossings of the horizontal ray from (px, py) towards -x, signed as path_count.wgsl:96 does (a line going down, y
    growing, counts -1)."""
    p0, p1 = lines["p0"].astype(np.float64), lines["p1"].astype(np.float64)
    y0, y1 = p0[:, 1], p1[:, 1]
    crosses = (y0 <= py) != (y1 <= py)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (py - y0) / (y1 - y0)
    x = p0[:, 0] + t * (p1[:, 0] - p0[:, 0])
    hit = crosses & (x < px)
    return int(np.where(y1[hit] > y0[hit], -1, 1).sum())


@pytest.mark.parametrize("seed,hang", [(1, False), (2, False), (3, True)])
def test_path_bboxes_backdrops_and_segments(oracle, seed, hang):
    W, H = 160, 144
    bump, lines = _run(oracle, _scene(seed, hang), W, H)
    assert len(parity.unpaired_endpoints(lines)) == 0  # watertight
    paths = oracle.buffer("paths")
    tiles = oracle.buffer("tiles")
    bboxes = oracle.buffer("path_bboxes")
    n_paths = int(lines["path_ix"].max()) + 1
    nonzero_backdrops = 0
    for pi in range(n_paths):
        mine = lines[lines["path_ix"] == pi]
        assert len(mine) > 0
        # ---- flatten's bounding box: integer bounds of the path's own lines
        xs = np.concatenate([mine["p0"][:, 0], mine["p1"][:, 0]])
        ys = np.concatenate([mine["p0"][:, 1], mine["p1"][:, 1]])
        bb = bboxes[pi]
        assert (int(bb["x0"]), int(bb["y0"]), int(bb["x1"]), int(bb["y1"])) == (math.floor(xs.min()), math.floor(ys.min()), math.ceil(xs.max()), math.ceil(ys.max()))
        # ---- backdrop: winding number at the tile's top-left corner (just inside, away from vertices)
        bx0, by0, bx1, by1 = (int(v) for v in paths[pi]["bbox"])
        base, stride = int(paths[pi]["tiles"]), bx1 - bx0
        for ty in range(by0, by1):
            for tx in range(bx0, bx1):
                b = int(tiles[base + (ty - by0) * stride + (tx - bx0)]["backdrop"])
                w = _winding(mine, 16 * tx + 1e-4, 16 * ty + 1e-4)
                assert b == w, (pi, tx, ty, b, w)
                nonzero_backdrops += b != 0
    assert nonzero_backdrops > 20  # the scenes do exercise it

    # ---- path_tiling: every segment inside its tile's 16 x 16 box (tile-relative coordinates), total length conserved
    segs = oracle.buffer("segments")[: int(bump["segments"])]
    for c in ("p0", "p1"):
        assert segs[c].min() >= -1e-3 and segs[c].max() <= 16.0 + 1e-3
    seg_len = float(np.hypot(*(segs["p1"] - segs["p0"]).astype(np.float64).T).sum())
    # length of the lines inside the frame (tile-aligned), by Liang-Barsky clipping in float64
    fw, fh = 16 * ((W + 15) // 16), 16 * ((H + 15) // 16)
    total = 0.0
    for ln in lines:
        (x0, y0), (x1, y1) = ln["p0"].astype(np.float64), ln["p1"].astype(np.float64)
        t0, t1 = 0.0, 1.0
        dx, dy = x1 - x0, y1 - y0
        ok = True
        for p, q in ((-dx, x0), (dx, fw - x0), (-dy, y0), (dy, fh - y0)):
            if p == 0:
                if q < 0:
                    ok = False
            else:
                r = q / p
                if p < 0:
                    t0 = max(t0, r)
                else:
                    t1 = min(t1, r)
        if ok and t1 > t0:
            total += (t1 - t0) * math.hypot(dx, dy)
    # horizontal lines lying exactly on a tile-row boundary carry no winding and are dropped (path_count.wgsl:77): none here
    assert abs(seg_len - total) / total < 2e-4, (seg_len, total)
