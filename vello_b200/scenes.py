"""Seeded scene generators for tests and benchmarks (all synthetic or restated recipes).

* golden recipes   : vello_tests/tests/smoke_snapshots.rs:17-52, regression.rs:33-210,
                     property.rs:21-197, known_issues.rs:21-52
* tiger            : examples/assets/Ghostscript_Tiger.svg via the draw list pico_svg produces
                     (examples/scenes/src/pico_svg.rs:134-195, svg.rs `render_svg_rec`), read from
                     tests/golden/tiger_paths.json.gz (see tests/golden/make_golden.py)
* paris_like       : BASELINE.md config C3/C4 -- paris-30k.svg is not in the reference, so a seeded
                     stand-in with the same order of magnitude (30k paths, ~1.5M segments, ~12 MB)
* beziers_clips    : BASELINE.md config C5
* robust_paths, funky_paths, fill_types, stroke_styles, many_clips, deep_blend : recipes restated
  from examples/scenes/src/test_scenes.rs (:1610-1691, :293-333, :699-770, :335-560, :1278-1304,
  :1241-1276)
* blend_grid (:1213-1239 + render_blend_square :1398-1436), tricky_strokes (:513-697), gradient_extend (:978-1043, the
  text labels need fonts and are left out), two_point_radial (:1045-1211), many_draw_objects (:1928-1948),
  conflation_artifacts (:1444-1531), image_extend_modes (:2168-2213), longpathdash (:779-819): the reference's recipes,
  statement for statement; compose_grid is ours (the reference has no scene that walks the Compose operators)
"""
from __future__ import annotations

import gzip
import json
import math
import os
from typing import List, Tuple

import numpy as np

from .encoding import (
    ALPHA_PREMULTIPLIED, ALPHA_STRAIGHT, BLACK, BLUE, COMPOSE_CLEAR, COMPOSE_SRC_OVER, COMPOSE_PLUS, COMPOSE_XOR,
    COMPOSE_COPY, COMPOSE_DEST, COMPOSE_DEST_OVER, COMPOSE_SRC_IN, COMPOSE_DEST_IN, COMPOSE_SRC_OUT, COMPOSE_DEST_OUT,
    COMPOSE_SRC_ATOP, COMPOSE_DEST_ATOP, COMPOSE_PLUS_LIGHTER,
    Color, EXTEND_PAD, EXTEND_REFLECT, EXTEND_REPEAT, Encoding, FILL_EVEN_ODD, FILL_NON_ZERO, FORMAT_BGRA8,
    FORMAT_RGBA8, Gradient, Image, LIME, MIX_MULTIPLY, MIX_NORMAL, MIX_SCREEN, MIX_HUE, MIX_DIFFERENCE, QUALITY_HIGH,
    QUALITY_LOW, QUALITY_MEDIUM, RED, STYLE_CAP_BUTT, STYLE_CAP_ROUND, STYLE_CAP_SQUARE, STYLE_JOIN_BEVEL,
    STYLE_JOIN_MITER, STYLE_JOIN_ROUND, Scene, Stroke, TAG_LINE_TO_F32, TAG_PATH, TAG_QUAD_TO_F32, TAG_CUBIC_TO_F32,
    TAG_SUBPATH_END_BIT, TRANSPARENT, WHITE, style_from_stroke, DRAWTAG_COLOR,
)
from .shapes import Affine, BezPath, Circle, Ellipse, Line, Rect, RoundedRect

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIGER_FIXTURE = os.path.join(_REPO, "tests", "golden", "tiger_paths.json.gz")


# ---------------------------------------------------------------------------------------------
# golden recipes
# ---------------------------------------------------------------------------------------------
def filled_square() -> Tuple[Scene, int, int]:
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, BLUE, None, Rect.from_center_size((10.0, 10.0), (6.0, 6.0)))
    return s, 20, 20


def filled_circle() -> Tuple[Scene, int, int]:
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, BLUE, None, Circle(10.0, 10.0, 7.0))
    return s, 20, 20


def simple_square() -> Tuple[Scene, int, int]:
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, RED, None, Rect.from_center_size((100.0, 100.0), (50.0, 50.0)))
    return s, 150, 150


def gradient_color_alpha(premul: bool) -> Tuple[Scene, int, int]:
    s = Scene()
    g = Gradient.linear((0.0, 0.0), (100.0, 0.0),
                        [(0.0, Color.from_rgba8(255, 255, 0, 0)), (1.0, Color.from_rgba8(0, 0, 255, 255))],
                        premul_interp=premul)
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, g, None, Rect(0.0, 0.0, 100.0, 50.0))
    return s, 100, 50


def image_roundtrip(img_rgba: np.ndarray, extend: int) -> Tuple[Scene, int, int]:
    s = Scene()
    s.draw_image(Image(img_rgba, quality=QUALITY_LOW, x_extend=extend, y_extend=extend), Affine.IDENTITY)
    return s, img_rgba.shape[1], img_rgba.shape[0]


def layer_size() -> Tuple[Scene, int, int]:
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8(0, 255, 0), None, Rect.from_origin_size((0.0, 0.0), (60.0, 60.0)))
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8(255, 0, 0), None, Rect.from_origin_size((20.0, 20.0), (20.0, 20.0)))
    s.push_layer(FILL_NON_ZERO, MIX_NORMAL, COMPOSE_CLEAR, 1.0, Affine.IDENTITY, Rect.from_origin_size((20.0, 20.0), (20.0, 20.0)))
    s.pop_layer()
    return s, 60, 60


# ---------------------------------------------------------------------------------------------
# tiger
# ---------------------------------------------------------------------------------------------
def _parse_hex(c: str) -> Color:
    c = c.strip()
    if c.startswith("#"):
        h = c[1:]
        if len(h) == 3:
            r, g, b = (int(ch * 2, 16) for ch in h)
        else:
            r, g, b = int(h[0:2], 16), int(h[2:4], 16), int(h[4:6], 16)
        return Color.from_rgba8(r, g, b)
    return Color(1.0, 0.0, 1.0, 0.5)  # pico_svg falls back to translucent fuchsia


def _opacity(c: Color, it: dict, key: str) -> Color:
    if key in it:
        v = it[key]
        a = float(v[:-1]) * 0.01 if v.endswith("%") else float(v)
        return c.with_alpha(float(np.float32(min(max(a, 0.0), 1.0))))
    return c


def tiger(width: int, height: int, fixture: str = TIGER_FIXTURE) -> Scene:
    """Ghostscript tiger scaled to fit (scale = min(w, h) / 200, vello_tests/src/lib.rs:293-299)."""
    with gzip.open(fixture, "rt") as f:
        doc = json.load(f)
    scale = min(width, height) / 200.0
    t = Affine.scale(scale)
    s = Scene()
    for it in doc["items"]:
        path = BezPath.from_svg(it["d"])
        if it.get("fill") is not None:
            col = _opacity(_opacity(_parse_hex(it["fill"]), it, "fill-opacity"), it, "opacity")
            s.fill(FILL_NON_ZERO, t, col, None, path)
        if it.get("stroke") not in (None, "none"):
            w = float(it.get("stroke-width", 1.0))
            col = _opacity(_opacity(_parse_hex(it["stroke"]), it, "stroke-opacity"), it, "opacity")
            s.stroke(Stroke(w), t, col, None, path)
    return s


# ---------------------------------------------------------------------------------------------
# bulk (numpy) appenders: identical streams to PathEncoder for well-formed polylines
# (every consecutive point pair differs by more than 1e-12), without the per-segment Python cost
# ---------------------------------------------------------------------------------------------
def _bulk_fill_polygon(e: Encoding, pts: np.ndarray):
    """pts: (n, 2) f32, n >= 3, first != last. M p0 L p1.. Z PATH (path.rs:508-676)."""
    n = pts.shape[0]
    e.path_data.extend(pts.reshape(-1).tolist())
    e.path_data.extend(pts[0].tolist())
    e.path_tags.extend([TAG_LINE_TO_F32] * (n - 1))
    e.path_tags.append(TAG_LINE_TO_F32 | TAG_SUBPATH_END_BIT)
    e.path_tags.append(TAG_PATH)
    e.n_path_segments += n
    e.n_paths += 1


def _bulk_stroke_polyline(e: Encoding, pts: np.ndarray):
    """Open polyline, n >= 2 points: lines, then the quad-to cap marker (path.rs:711-730)."""
    n = pts.shape[0]
    f = np.float32
    p0, p1 = pts[0], pts[1]
    third = f(1.0) / f(3.0)
    tan = (p0[0] + third * (p1[0] - p0[0]), p0[1] + third * (p1[1] - p0[1]))
    e.path_data.extend(pts.reshape(-1).tolist())
    e.path_data.extend([float(p0[0]), float(p0[1]), float(tan[0]), float(tan[1])])
    e.path_tags.extend([TAG_LINE_TO_F32] * (n - 1))
    e.path_tags.append(TAG_QUAD_TO_F32 | TAG_SUBPATH_END_BIT)
    e.path_tags.append(TAG_PATH)
    e.n_path_segments += n
    e.n_paths += 1


def _bulk_fill_cubics(e: Encoding, p0: np.ndarray, ctrl: np.ndarray):
    """Closed loop of cubics: p0 (2,), ctrl (n, 3, 2) with ctrl[-1, 2] == p0."""
    n = ctrl.shape[0]
    e.path_data.extend(p0.tolist())
    e.path_data.extend(ctrl.reshape(-1).tolist())
    e.path_tags.extend([TAG_CUBIC_TO_F32] * (n - 1))
    e.path_tags.append(TAG_CUBIC_TO_F32 | TAG_SUBPATH_END_BIT)
    e.path_tags.append(TAG_PATH)
    e.n_path_segments += n
    e.n_paths += 1


def _rand_color(rng: np.random.Generator, alpha_prob=0.2) -> Color:
    r, g, b = (int(v) for v in rng.integers(0, 256, 3))
    a = int(rng.integers(64, 230)) if rng.random() < alpha_prob else 255
    return Color.from_rgba8(r, g, b, a)


def paris_like(n_paths: int = 30000, size: int = 4096, seed: int = 30000) -> Scene:
    """Seeded stand-in for paris-30k: 70% filled polygons (8-60 segs, extent log-uniform 4-400 px at
    4096), 25% stroked polylines (20-200 segs, width 0.5-6), 5% closed cubic blobs; uniform over
    the canvas; 20% translucent colours. Designed at 4096 and scaled by `size / 4096`."""
    rng = np.random.default_rng(seed)
    s = Scene()
    e = s.encoding
    t = Affine.scale(size / 4096.0)
    f32 = np.float32
    for _ in range(n_paths):
        kind = rng.random()
        cx, cy = rng.uniform(0, 4096, 2)
        extent = math.exp(rng.uniform(math.log(4.0), math.log(400.0)))
        col = _rand_color(rng)
        if kind < 0.70:
            n = int(rng.integers(8, 61))
            ang = np.sort(rng.uniform(0, 2 * math.pi, n))
            rad = rng.uniform(0.35, 1.0, n) * (extent * 0.5)
            pts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1).astype(f32)
            e.encode_transform(t)
            e.encode_fill_style(FILL_NON_ZERO if rng.random() < 0.9 else FILL_EVEN_ODD)
            _bulk_fill_polygon(e, pts)
            e.encode_color(col)
        elif kind < 0.95:
            n = int(rng.integers(20, 201))
            step = extent / 16.0 + 1.0
            heading = rng.uniform(0, 2 * math.pi) + np.cumsum(rng.normal(0, 0.25, n))
            d = np.stack([np.cos(heading), np.sin(heading)], axis=1) * step
            pts = (np.array([cx, cy]) + np.cumsum(d, axis=0)).astype(f32)
            join = (STYLE_JOIN_ROUND, STYLE_JOIN_BEVEL, STYLE_JOIN_MITER)[int(rng.integers(0, 3))]
            cap = (STYLE_CAP_ROUND, STYLE_CAP_BUTT, STYLE_CAP_SQUARE)[int(rng.integers(0, 3))]
            st = Stroke(float(f32(rng.uniform(0.5, 6.0))), join=join, start_cap=cap, end_cap=cap)
            e.encode_transform(t)
            assert e.encode_stroke_style(st)
            _bulk_stroke_polyline(e, pts)
            e.encode_color(col)
        else:
            n = int(rng.integers(6, 13))
            ang = np.linspace(0, 2 * math.pi, n, endpoint=False) + rng.uniform(0, 1)
            rad = rng.uniform(0.5, 1.0, n) * (extent * 0.5)
            on = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
            tang = np.stack([-np.sin(ang), np.cos(ang)], axis=1) * (rad * (2 * math.pi / n) * 0.4)[:, None]
            ctrl = np.zeros((n, 3, 2))
            for i in range(n):
                j = (i + 1) % n
                ctrl[i, 0] = on[i] + tang[i]
                ctrl[i, 1] = on[j] - tang[j]
                ctrl[i, 2] = on[j]
            e.encode_transform(t)
            e.encode_fill_style(FILL_NON_ZERO)
            _bulk_fill_cubics(e, on[0].astype(f32), ctrl.astype(f32))
            e.encode_color(col)
    return s


def beziers_clips(n_paths: int = 100000, n_clips: int = 1000, size: int = 4096, seed: int = 100000,
                  max_depth: int = 8) -> Scene:
    """Config C5: cubic paths (4-12 cubics each) interleaved with `n_clips` push_clip_layer/pop pairs
    nested up to `max_depth`."""
    rng = np.random.default_rng(seed)
    s = Scene()
    e = s.encoding
    f32 = np.float32
    t = Affine.scale(size / 4096.0)
    clip_every = max(1, n_paths // max(1, n_clips))
    depth = 0
    pushed = 0
    for i in range(n_paths):
        if n_clips and i % clip_every == 0 and pushed < n_clips:
            if depth >= max_depth or (depth > 0 and rng.random() < 0.35):
                npop = int(rng.integers(1, depth + 1))
                for _ in range(npop):
                    s.pop_layer()
                depth -= npop
            cx, cy = rng.uniform(0, 4096, 2)
            r = rng.uniform(200, 1200)
            s.push_clip_layer(FILL_NON_ZERO, t, Circle(cx, cy, r) if rng.random() < 0.5 else Rect(cx - r, cy - r, cx + r, cy + r))
            depth += 1
            pushed += 1
        n = int(rng.integers(4, 13))
        cx, cy = rng.uniform(0, 4096, 2)
        extent = math.exp(rng.uniform(math.log(8.0), math.log(300.0)))
        ang = np.linspace(0, 2 * math.pi, n, endpoint=False) + rng.uniform(0, 1)
        rad = rng.uniform(0.4, 1.0, n) * (extent * 0.5)
        on = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
        tang = np.stack([-np.sin(ang), np.cos(ang)], axis=1) * (rad * (2 * math.pi / n) * rng.uniform(0.2, 0.7))[:, None]
        ctrl = np.zeros((n, 3, 2))
        for k in range(n):
            j = (k + 1) % n
            ctrl[k, 0] = on[k] + tang[k]
            ctrl[k, 1] = on[j] - tang[j]
            ctrl[k, 2] = on[j]
        e.encode_transform(t)
        e.encode_fill_style(FILL_NON_ZERO)
        _bulk_fill_cubics(e, on[0].astype(f32), ctrl.astype(f32))
        e.encode_color(_rand_color(rng, 0.3))
    while depth > 0:
        s.pop_layer()
        depth -= 1
    return s


# ---------------------------------------------------------------------------------------------
# restated test-scene recipes (small; exercise robustness and every draw-object kind)
# ---------------------------------------------------------------------------------------------
def robust_paths() -> Tuple[Scene, int, int]:
    """Tile-boundary-aligned rectangles/triangles and half-pixel slivers, both fill rules."""
    s = Scene()
    p = BezPath()
    for (x0, y0, x1, y1) in [(16, 16, 32, 32), (48, 16, 64, 32), (32, 48, 48, 64), (80, 16, 96.5, 32.5)]:
        p.move_to(x0, y0); p.line_to(x1, y0); p.line_to(x1, y1); p.line_to(x0, y1); p.close_path()
    p.move_to(8, 100); p.line_to(8.5, 100); p.line_to(8.5, 116); p.line_to(8, 116); p.close_path()
    p.move_to(16, 80); p.line_to(48, 80); p.line_to(16, 96); p.close_path()
    p.move_to(64, 64); p.line_to(96, 64); p.line_to(96, 96); p.close_path()
    # overlapping squares with opposite winding
    p.move_to(100, 20); p.line_to(130, 20); p.line_to(130, 50); p.line_to(100, 50); p.close_path()
    p.move_to(110, 30); p.line_to(110, 60); p.line_to(140, 60); p.line_to(140, 30); p.close_path()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8(255, 255, 0), None, p)
    s.fill(FILL_EVEN_ODD, Affine.translate(0.0, 130.0), Color.from_rgba8(0, 255, 255), None, p)
    # exact tile-grid aligned big rect and lines on vertical tile edges
    s.fill(FILL_NON_ZERO, Affine.translate(160.0, 0.0), Color.from_rgba8(200, 80, 200, 160), None, Rect(0, 0, 64, 64))
    s.fill(FILL_NON_ZERO, Affine.translate(160.0, 80.0), Color.from_rgba8(80, 200, 80), None, Rect(-20, -20, 48, 48))
    return s, 256, 272


def funky_paths() -> Tuple[Scene, int, int]:
    """Missing move-tos, empty paths, only-move-tos (PathEncoder state machine)."""
    s = Scene()
    missing = BezPath([("L", 100.0, 100.0), ("L", 100.0, 200.0), ("Z",), ("L", 0.0, 400.0), ("L", 100.0, 400.0)])
    only_moves = BezPath([("M", 0.0, 0.0), ("M", 100.0, 100.0)])
    empty = BezPath()
    s.fill(FILL_NON_ZERO, Affine.translate(100.0, 100.0), BLUE, None, missing)
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, BLUE, None, empty)
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, BLUE, None, only_moves)
    s.stroke(Stroke(8.0), Affine.translate(100.0, 100.0), Color.from_rgba8(0, 255, 255), None, missing)
    return s, 320, 560


def fill_types() -> Tuple[Scene, int, int]:
    s = Scene()
    star = BezPath()
    for i in range(5):
        a = -math.pi / 2 + i * 4 * math.pi / 5
        (star.move_to if i == 0 else star.line_to)(50 + 45 * math.cos(a), 50 + 45 * math.sin(a))
    star.close_path()
    arcs = BezPath.from_svg("M10 50 A40 40 0 1 1 90 50 A40 40 0 1 1 10 50 Z M30 50 A20 20 0 1 0 70 50 A20 20 0 1 0 30 50 Z")
    for i, (rule, shape) in enumerate([(FILL_NON_ZERO, star), (FILL_EVEN_ODD, star), (FILL_NON_ZERO, arcs), (FILL_EVEN_ODD, arcs)]):
        t = Affine.translate(10.0 + 110.0 * i, 10.0)
        s.fill(rule, t, Color.from_rgba8(230, 60, 60), None, shape)
        s.fill(rule, t * Affine.translate(0.0, 110.0) * Affine.rotate(0.06), Color.from_rgba8(40, 40, 230, 128), None, shape)
    return s, 460, 240


def stroke_styles(transform: Affine = Affine.IDENTITY) -> Tuple[Scene, int, int]:
    s = Scene()
    zig = BezPath([("M", 10.0, 10.0), ("L", 60.0, 40.0), ("L", 10.0, 70.0), ("L", 60.0, 100.0)])
    curve = BezPath([("M", 0.0, 0.0), ("C", 40.0, -30.0, 60.0, 60.0, 100.0, 20.0), ("Q", 130.0, -10.0, 150.0, 40.0)])
    closed = BezPath([("M", 10.0, 10.0), ("L", 90.0, 20.0), ("C", 120.0, 60.0, 40.0, 90.0, 20.0, 60.0), ("Z",)])
    caps = [STYLE_CAP_BUTT, STYLE_CAP_SQUARE, STYLE_CAP_ROUND]
    joins = [STYLE_JOIN_BEVEL, STYLE_JOIN_MITER, STYLE_JOIN_ROUND]
    y = 10.0
    for ci, cap in enumerate(caps):
        for ji, join in enumerate(joins):
            st = Stroke(10.0, join=join, start_cap=cap, end_cap=caps[(ci + 1) % 3], miter_limit=4.0)
            t = transform * Affine.translate(10.0 + 170.0 * ji, y)
            s.stroke(st, t, Color.from_rgba8(200, 200 - 60 * ji, 40 + 80 * ci), None, zig)
            s.stroke(Stroke(3.0, join=join, start_cap=cap, end_cap=cap), t * Affine.translate(70.0, 30.0), WHITE, None, curve)
        y += 120.0
    s.stroke(Stroke(6.0, join=STYLE_JOIN_MITER, miter_limit=1.5), transform * Affine.translate(20.0, y), LIME, None, closed)
    s.stroke(Stroke(6.0, join=STYLE_JOIN_ROUND), transform * Affine.translate(160.0, y), RED, None, closed)
    s.stroke(Stroke(0.5), transform * Affine.translate(300.0, y), WHITE, None, closed)
    s.stroke(Stroke(4.0), transform * Affine.translate(420.0, y + 20.0), WHITE, None, Line(0.0, 0.0, 60.0, 60.0))
    s.stroke(Stroke(5.0), transform * Affine.translate(500.0, y + 50.0), BLUE, None, Circle(0.0, 0.0, 30.0))
    return s, 560, 480


def many_clips(n: int = 40, depth: int = 12) -> Tuple[Scene, int, int]:
    rng = np.random.default_rng(7)
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8(30, 30, 60), None, Rect(0, 0, 400, 400))
    for i in range(n):
        d = int(rng.integers(1, depth + 1))
        for k in range(d):
            cx, cy = rng.uniform(50, 350, 2)
            r = rng.uniform(60, 220)
            shape = Circle(cx, cy, r) if (i + k) % 2 == 0 else RoundedRect(cx - r, cy - r * 0.7, cx + r, cy + r * 0.7, 20.0)
            s.push_clip_layer(FILL_NON_ZERO if k % 3 else FILL_EVEN_ODD, Affine.IDENTITY, shape)
        x, y = rng.uniform(0, 300, 2)
        s.fill(FILL_NON_ZERO, Affine.IDENTITY, _rand_color(rng, 0.5), None, Rect(x, y, x + rng.uniform(40, 300), y + rng.uniform(40, 300)))
        for k in range(d):
            s.pop_layer()
    return s, 400, 400


def deep_blend(depth: int = 9) -> Tuple[Scene, int, int]:
    """Nested non-clip blend layers deeper than BLEND_STACK_SPLIT=4 (exercises blend_spill)."""
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8(240, 240, 240), None, Rect(0, 0, 200, 200))
    modes = [(MIX_MULTIPLY, COMPOSE_SRC_OVER), (MIX_SCREEN, COMPOSE_SRC_OVER), (MIX_NORMAL, COMPOSE_XOR),
             (MIX_HUE, COMPOSE_SRC_OVER), (MIX_DIFFERENCE, COMPOSE_SRC_OVER), (MIX_NORMAL, COMPOSE_PLUS)]
    for i in range(depth):
        m, c = modes[i % len(modes)]
        s.push_layer(FILL_NON_ZERO, m, c, 0.9, Affine.IDENTITY, Rect(10 + 8 * i, 10 + 6 * i, 190 - 5 * i, 190 - 7 * i))
        s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8((40 * i) % 256, (90 * i + 30) % 256, (160 + 25 * i) % 256, 200),
               None, Circle(60 + 9 * i, 70 + 8 * i, 50.0))
    for i in range(depth):
        s.pop_layer()
    return s, 200, 200


def brushes() -> Tuple[Scene, int, int]:
    """Every brush kind: linear/radial (all 4 kinds)/sweep gradients with all extends, images at the
    three qualities, blurred rounded rect, luminance mask layer."""
    rng = np.random.default_rng(11)
    s = Scene()
    stops = [(0.0, Color.from_rgba8(255, 0, 0)), (0.5, Color.from_rgba8(0, 255, 0, 128)), (1.0, Color.from_rgba8(0, 0, 255))]
    x = 0.0
    for ext in (EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT):
        s.fill(FILL_NON_ZERO, Affine.translate(x, 0.0), Gradient.linear((20, 10), (60, 50), stops, ext), None, Rect(0, 0, 100, 80))
        s.fill(FILL_NON_ZERO, Affine.translate(x, 90.0), Gradient.radial((50, 40), 5.0, (55, 45), 35.0, stops, ext), None, Rect(0, 0, 100, 80))
        s.fill(FILL_NON_ZERO, Affine.translate(x, 180.0), Gradient.sweep((50, 40), 0.3, 5.0, stops, ext), None, Rect(0, 0, 100, 80))
        x += 110.0
    # radial kinds: strip (equal radii), focal on circle, circular (same centre), swapped (r1 == 0)
    s.fill(FILL_NON_ZERO, Affine.translate(0.0, 270.0), Gradient.radial((20, 40), 20.0, (80, 40), 20.0, stops), None, Rect(0, 0, 100, 80))
    s.fill(FILL_NON_ZERO, Affine.translate(110.0, 270.0), Gradient.radial((30, 40), 0.0, (60, 40), 30.0, stops), None, Rect(0, 0, 100, 80))
    s.fill(FILL_NON_ZERO, Affine.translate(220.0, 270.0), Gradient.radial((50, 40), 5.0, (50, 40), 40.0, stops), None, Rect(0, 0, 100, 80))
    s.fill(FILL_NON_ZERO, Affine.translate(330.0, 270.0), Gradient.radial((40, 40), 30.0, (60, 40), 0.0, stops), None, Rect(0, 0, 100, 80))
    img = rng.integers(0, 256, (13, 17, 4), dtype=np.uint8)
    for i, q in enumerate((QUALITY_LOW, QUALITY_MEDIUM, QUALITY_HIGH)):
        im = Image(img, quality=q, x_extend=(EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT)[i], y_extend=EXTEND_REFLECT,
                   alpha=0.9, alpha_type=ALPHA_STRAIGHT if i != 1 else ALPHA_PREMULTIPLIED,
                   format=FORMAT_RGBA8 if i != 2 else FORMAT_BGRA8)
        t = Affine.translate(340.0, 10.0 + 85.0 * i) * Affine.rotate(0.1 * i) * Affine.scale(4.5)
        s.fill(FILL_NON_ZERO, Affine.IDENTITY, im, t, Rect(340, 10 + 85 * i, 440, 85 + 85 * i))
    s.draw_blurred_rounded_rect(Affine.translate(60.0, 400.0), Rect(-40, -25, 40, 25), Color.from_rgba8(250, 200, 30), 8.0, 5.0)
    s.draw_blurred_rounded_rect(Affine.translate(200.0, 400.0) * Affine.rotate(0.3), Rect(-50, -20, 50, 20), Color.from_rgba8(30, 200, 250, 200), 0.0, 2.0)
    s.push_luminance_mask_layer(FILL_NON_ZERO, 1.0, Affine.IDENTITY, Rect(280, 360, 440, 440))
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, Gradient.linear((280, 360), (440, 440), [(0.0, BLACK), (1.0, WHITE)]), None, Rect(280, 360, 440, 440))
    s.pop_layer()
    return s, 450, 450


def random_small(seed: int, n: int = 40, size: int = 200) -> Tuple[Scene, int, int]:
    """Mixed random fills / strokes / clips for randomized parity tests."""
    rng = np.random.default_rng(seed)
    s = Scene()
    depth = 0
    for i in range(n):
        k = rng.random()
        p = BezPath()
        m = int(rng.integers(2, 7))
        x, y = rng.uniform(-20, size + 20, 2)
        p.move_to(x, y)
        for _ in range(m):
            c = rng.uniform(-40, size + 40, 6)
            r = rng.random()
            if r < 0.4:
                p.line_to(c[0], c[1])
            elif r < 0.7:
                p.quad_to(c[0], c[1], c[2], c[3])
            else:
                p.curve_to(*c)
        t = Affine.translate(*rng.uniform(-5, 5, 2)) * Affine.rotate(rng.uniform(-0.3, 0.3)) * Affine.scale(rng.uniform(0.5, 1.5))
        if k < 0.5:
            s.fill(FILL_NON_ZERO if rng.random() < 0.5 else FILL_EVEN_ODD, t, _rand_color(rng, 0.5), None, p)
        elif k < 0.85:
            st = Stroke(float(rng.uniform(0.3, 12.0)), join=(STYLE_JOIN_BEVEL, STYLE_JOIN_MITER, STYLE_JOIN_ROUND)[int(rng.integers(0, 3))],
                        start_cap=(STYLE_CAP_BUTT, STYLE_CAP_SQUARE, STYLE_CAP_ROUND)[int(rng.integers(0, 3))],
                        end_cap=(STYLE_CAP_BUTT, STYLE_CAP_SQUARE, STYLE_CAP_ROUND)[int(rng.integers(0, 3))])
            if rng.random() < 0.3:
                p.close_path()
            s.stroke(st, t, _rand_color(rng, 0.5), None, p)
        elif depth < 5:
            s.push_clip_layer(FILL_NON_ZERO, t, p)
            depth += 1
        elif depth > 0:
            s.pop_layer()
            depth -= 1
    # leave some clips open on purpose: resolve closes them (resolve.rs:127-141)
    return s, size, size


# ---------------------------------------------------------------------------------------------
# more recipes of examples/scenes/src/test_scenes.rs, statement for statement
# ---------------------------------------------------------------------------------------------
YELLOW = Color.from_rgba8(255, 255, 0)
CYAN = Color.from_rgba8(0, 255, 255)
MAGENTA = Color.from_rgba8(255, 0, 255)


def _even_stops(colors):
    """`Gradient::with_stops([c0, c1, ..])`: evenly spaced offsets (peniko ColorStops from a colour slice)."""
    n = len(colors)
    return [(float(np.float32(i) / np.float32(n - 1)) if n > 1 else 0.0, c) for i, c in enumerate(colors)]


def render_blend_square(s: Scene, mix: int, compose: int, transform: Affine):
    """test_scenes.rs:1398-1436 (`blend` = BlendMode { mix, compose })."""
    rect = Rect.from_origin_size((0.0, 0.0), (200.0, 200.0))
    s.fill(FILL_NON_ZERO, transform, Gradient.linear((0.0, 0.0), (200.0, 0.0), _even_stops([BLACK, WHITE])), None, rect)
    for (x, y, c) in ((150.0, 0.0, Color.from_rgba8(255, 240, 64)), (175.0, 100.0, Color.from_rgba8(255, 96, 240)),
                      (125.0, 200.0, Color.from_rgba8(64, 192, 255))):
        # Gradient::new_radial(center, r) = two-point radial with both centres equal, r0 = 0
        s.fill(FILL_NON_ZERO, transform, Gradient.radial((x, y), 0.0, (x, y), 100.0, _even_stops([c, c.with_alpha(0.0)])), None, rect)
    s.push_layer(FILL_NON_ZERO, MIX_NORMAL, COMPOSE_SRC_OVER, 1.0, transform, rect)
    for i, c in enumerate((RED, LIME, BLUE)):
        linear = Gradient.linear((0.0, 0.0), (0.0, 200.0), _even_stops([WHITE, c]))
        s.push_layer(FILL_NON_ZERO, mix, compose, 1.0, transform, rect)
        a = (transform * Affine.translate(100.0, 100.0) * Affine.rotate(math.pi / 3.0 * (i * 2 + 1))
             * Affine.scale(1.0, 0.357) * Affine.translate(-100.0, -100.0))
        s.fill(FILL_NON_ZERO, a, linear, None, Ellipse(100.0, 100.0, 90.0, 90.0, 0.0))
        s.pop_layer()
    s.pop_layer()


def blend_grid() -> Tuple[Scene, int, int]:
    """test_scenes.rs:1213-1239: the 16 Mix modes (in the reference's order), each as a fragment appended with a
    translation (`Scene::append`)."""
    from .encoding import (MIX_OVERLAY, MIX_DARKEN, MIX_LIGHTEN, MIX_COLOR_DODGE, MIX_COLOR_BURN, MIX_HARD_LIGHT, MIX_SOFT_LIGHT,
                           MIX_EXCLUSION, MIX_SATURATION, MIX_COLOR, MIX_LUMINOSITY)
    modes = [MIX_NORMAL, MIX_MULTIPLY, MIX_DARKEN, MIX_SCREEN, MIX_LIGHTEN, MIX_OVERLAY, MIX_COLOR_DODGE, MIX_COLOR_BURN,
             MIX_HARD_LIGHT, MIX_SOFT_LIGHT, MIX_DIFFERENCE, MIX_EXCLUSION, MIX_HUE, MIX_SATURATION, MIX_COLOR, MIX_LUMINOSITY]
    s = Scene()
    for ix, mix in enumerate(modes):
        frag = Scene()
        render_blend_square(frag, mix, COMPOSE_SRC_OVER, Affine.IDENTITY)
        s.append(frag, Affine.translate((ix % 4) * 225.0, (ix // 4) * 225.0))
    return s, 900, 900


def compose_grid() -> Tuple[Scene, int, int]:
    """All 14 Compose operators (peniko `Compose`, blend.wgsl:255-331) on the blend square, each with Mix::Normal and with a
    separable (Multiply) and a non-separable (Luminosity) mix in front of it."""
    from .encoding import MIX_LUMINOSITY
    composes = [COMPOSE_CLEAR, COMPOSE_COPY, COMPOSE_DEST, COMPOSE_SRC_OVER, COMPOSE_DEST_OVER, COMPOSE_SRC_IN, COMPOSE_DEST_IN,
                COMPOSE_SRC_OUT, COMPOSE_DEST_OUT, COMPOSE_SRC_ATOP, COMPOSE_DEST_ATOP, COMPOSE_XOR, COMPOSE_PLUS, COMPOSE_PLUS_LIGHTER]
    s = Scene()
    for ix, comp in enumerate(composes):
        for k, mix in enumerate((MIX_NORMAL, MIX_MULTIPLY, MIX_LUMINOSITY)):
            render_blend_square(s, mix, comp, Affine.translate((ix % 7) * 210.0, (ix // 7) * 630.0 + k * 210.0) * Affine.scale(1.0))
    return s, 7 * 210, 6 * 210


def _cubic_bbox(p):
    """kurbo `CubicBez::bounding_box`: end points + the interior extrema of each coordinate."""
    lo = [min(p[0][k], p[3][k]) for k in (0, 1)]
    hi = [max(p[0][k], p[3][k]) for k in (0, 1)]
    for k in (0, 1):
        p0, p1, p2, p3 = (q[k] for q in p)
        a, b, c = 3.0 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3), 6.0 * (p0 - 2.0 * p1 + p2), 3.0 * (p1 - p0)
        ts = []
        if abs(a) < 1e-12:
            if abs(b) > 1e-12:
                ts.append(-c / b)
        else:
            disc = b * b - 4.0 * a * c
            if disc >= 0.0:
                r = math.sqrt(disc)
                ts += [(-b + r) / (2.0 * a), (-b - r) / (2.0 * a)]
        for t in ts:
            if 0.0 < t < 1.0:
                mt = 1.0 - t
                v = mt * mt * mt * p0 + 3.0 * mt * mt * t * p1 + 3.0 * mt * t * t * p2 + t * t * t * p3
                lo[k], hi[k] = min(lo[k], v), max(hi[k], v)
    return lo[0], lo[1], hi[0], hi[1]


def _map_rect_to_rect(src, dst):
    sw, sh, dw, dh = src[2] - src[0], src[3] - src[1], dst[2] - dst[0], dst[3] - dst[1]
    sx, sy = dw / sw, dh / sh
    scale, x_larger = min(sx, sy), sx > sy
    tx, ty = dst[0] - src[0] * scale, dst[1] - src[1] * scale
    if x_larger:
        tx += 0.5 * (dw - sw * scale)
    else:
        ty += 0.5 * (dh - sh * scale)
    return Affine((scale, 0.0, 0.0, scale, tx, ty)), scale


def tricky_strokes() -> Tuple[Scene, int, int]:
    """test_scenes.rs:513-697 (Skia's trickycubicstrokes): cusps, near-cusps, flat cubics with 180-degree turns, degenerate
    circles, flat conics as quads; butt caps, miter joins, stroke width 30 in cell space."""
    colors = [Color.from_rgba8(140, 181, 236), Color.from_rgba8(246, 236, 202), Color.from_rgba8(201, 147, 206), Color.from_rgba8(150, 195, 160)]
    CELL, SW, NCOLS = 200.0, 30.0, 5
    tricky = [
        [(122., 737.), (348., 553.), (403., 761.), (400., 760.)], [(244., 520.), (244., 518.), (1141., 634.), (394., 688.)],
        [(550., 194.), (138., 130.), (1035., 246.), (288., 300.)], [(226., 733.), (556., 779.), (-43., 471.), (348., 683.)],
        [(268., 204.), (492., 304.), (352., 23.), (433., 412.)], [(172., 480.), (396., 580.), (256., 299.), (338., 677.)],
        [(731., 340.), (318., 252.), (1026., -64.), (367., 265.)], [(475., 708.), (62., 620.), (770., 304.), (220., 659.)],
        [(0., 0.), (128., 128.), (128., 0.), (0., 128.)], [(0., 0.01), (128., 127.999), (128., 0.01), (0., 127.99)],
        [(0., -0.01), (128., 128.001), (128., -0.01), (0., 128.001)], [(0., 0.), (0., -10.), (0., -10.), (0., 10.)],
        [(10., 0.), (0., 0.), (20., 0.), (10., 0.)], [(39., -39.), (40., -40.), (40., -40.), (0., 0.)],
        [(40., 40.), (0., 0.), (200., 200.), (0., 0.)], [(0., 0.), (1e-2, 0.), (-1e-2, 0.), (0., 0.)],
        [(400.75, 100.05), (400.75, 100.05), (100.05, 300.95), (100.05, 300.95)],
        [(0.5, 0.), (0., 0.), (20., 0.), (10., 0.)], [(10., 0.), (0., 0.), (10., 0.), (10., 0.)],
    ]
    flat_quad = [[(2., 1.), (1., 1.)]]
    flat_conic = [[(2.232486, 1.0), (3.471740, 1.0)], [(4.710995, 1.0), (5.949262, 1.0)], [(7.187530, 1.0), (8.417061, 1.0)],
                  [(9.646591, 1.0), (10.859690, 1.0)], [(12.072789, 1.0), (13.261865, 1.0)], [(14.450940, 1.0), (15.608549, 1.0)],
                  [(16.766161, 1.0), (17.885059, 1.0)], [(19.003958, 1.0), (20.077141, 1.0)], [(21.150328, 1.0), (22.171083, 1.0)],
                  [(23.191839, 1.0), (24.153776, 1.0)], [(25.115715, 1.0), (26.012812, 1.0)], [(26.909912, 1.0), (27.736557, 1.0)],
                  [(28.563202, 1.0), (29.314220, 1.0)], [(30.065239, 1.0), (30.735928, 1.0)], [(31.406620, 1.0), (31.992788, 1.0)],
                  [(32.578957, 1.0), (33.076927, 1.0)], [(33.574905, 1.0), (33.981567, 1.0)], [(34.388233, 1.0), (34.701038, 1.0)],
                  [(35.013851, 1.0), (35.230850, 1.0)], [(35.447845, 1.0), (35.567669, 1.0)], [(35.687500, 1.0), (35.709404, 1.0)],
                  [(35.731312, 1.0), (35.655155, 1.0)], [(35.579006, 1.0), (35.405273, 1.0)], [(35.231541, 1.0), (34.961311, 1.0)],
                  [(34.691086, 1.0), (34.326057, 1.0)], [(33.961029, 1.0), (33.503479, 1.0)], [(33.045937, 1.0), (32.498734, 1.0)],
                  [(31.951530, 1.0), (31.318098, 1.0)], [(30.684669, 1.0), (29.968971, 1.0)], [(29.253277, 1.0), (28.459791, 1.0)],
                  [(27.666309, 1.0), (26.800005, 1.0)], [(25.933704, 1.0), (25.000000, 1.0)]]
    bigger_conic = [[(8.979845, 1.0), (15.795975, 1.0)], [(22.612104, 1.0), (28.363287, 1.0)], [(34.114471, 1.0), (38.884045, 1.0)],
                    [(43.653618, 1.0), (47.510696, 1.0)], [(51.367767, 1.0), (54.368233, 1.0)], [(57.368698, 1.0), (59.556030, 1.0)],
                    [(61.743366, 1.0), (63.149269, 1.0)], [(64.555168, 1.0), (65.200005, 1.0)], [(65.844841, 1.0), (65.737961, 1.0)],
                    [(65.631073, 1.0), (64.770912, 1.0)], [(63.910763, 1.0), (62.284878, 1.0)], [(60.658997, 1.0), (58.243816, 1.0)],
                    [(55.828640, 1.0), (52.589172, 1.0)], [(49.349705, 1.0), (45.239006, 1.0)], [(41.128315, 1.0), (36.086826, 1.0)],
                    [(31.045338, 1.0), (25.000000, 1.0)]]
    s = Scene()
    idx = color_idx = 0

    def cell_of(i):
        x, y = (i % NCOLS) * CELL, (i // NCOLS) * CELL
        return (x, y, x + CELL, y + CELL)

    for i, cubic in enumerate(tricky):
        idx += 1
        b = _cubic_bbox(cubic)
        t, sc = _map_rect_to_rect((b[0] - SW, b[1] - SW, b[2] + SW, b[3] + SW), cell_of(i))
        st = Stroke(SW / sc, join=STYLE_JOIN_MITER, start_cap=STYLE_CAP_BUTT, end_cap=STYLE_CAP_BUTT)
        s.stroke(st, t, colors[color_idx], None, BezPath([("M",) + cubic[0], ("C",) + cubic[1] + cubic[2] + cubic[3]]))
        color_idx = (color_idx + 1) % len(colors)
    for quads in (flat_quad, flat_conic, bigger_conic):
        path = BezPath([("M", 1.0, 1.0)] + [("Q",) + q[0] + q[1] for q in quads])
        xs = [1.0] + [q[k][0] for q in quads for k in (0, 1)]  # every control polygon is flat (y == 1): the bbox is the x range of
        # the curve, which for these monotone-per-piece flat quads is reached at on-curve points or inside one piece
        lo, hi = min(1.0, *[q[1][0] for q in quads]), max(1.0, *[q[1][0] for q in quads])
        p0 = 1.0
        for q in quads:  # interior extremum of a quadratic in x
            c, p2 = q[0][0], q[1][0]
            den = p0 - 2.0 * c + p2
            if den != 0.0:
                tt = (p0 - c) / den
                if 0.0 < tt < 1.0:
                    v = (1 - tt) * (1 - tt) * p0 + 2 * (1 - tt) * tt * c + tt * tt * p2
                    lo, hi = min(lo, v), max(hi, v)
            p0 = p2
        t, sc = _map_rect_to_rect((lo - SW, 1.0 - SW, hi + SW, 1.0 + SW), cell_of(idx))
        st = Stroke(SW / sc, join=STYLE_JOIN_MITER, start_cap=STYLE_CAP_BUTT, end_cap=STYLE_CAP_BUTT)
        s.stroke(st, t, colors[color_idx], None, path)
        color_idx = (color_idx + 1) % len(colors)
        idx += 1
    n = len(tricky) + 3
    return s, int(CELL * NCOLS), int(CELL * (1 + n // NCOLS))


def gradient_extend() -> Tuple[Scene, int, int]:
    """test_scenes.rs:978-1043 without the three text labels (glyph runs are out of scope)."""
    s = Scene()
    colors = _even_stops([RED, LIME, BLUE])
    w = h = 300.0
    deg = float(np.float32(np.pi) / np.float32(180.0))
    for x, ext in enumerate((EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT)):
        for y, kind in enumerate(("linear", "radial", "sweep")):
            t = Affine.translate(x * 350.0 + 50.0, y * 350.0 + 100.0)
            if kind == "linear":
                g = Gradient.linear((w * 0.35, h * 0.5), (w * 0.65, h * 0.5), colors, ext)
            elif kind == "radial":
                radius = float(np.float32(w * 0.25))
                g = Gradient.radial((w * 0.5, h * 0.5), float(np.float32(radius) * np.float32(0.25)), (w * 0.5, h * 0.5), radius, colors, ext)
            else:
                g = Gradient.sweep((w * 0.5, h * 0.5), float(np.float32(30.0) * np.float32(deg)), float(np.float32(150.0) * np.float32(deg)), colors, ext)
            s.fill(FILL_NON_ZERO, t, g, None, Rect(0.0, 0.0, w, h))
    return s, 1200, 1200


def two_point_radial() -> Tuple[Scene, int, int]:
    """test_scenes.rs:1045-1211: the COLR radial-gradient cases (disjoint circles both ways, equal radii = strip, nested,
    touching = focal on circle), each with the three extend modes."""
    s = Scene()
    stops = _even_stops([RED, YELLOW, Color.from_rgba8(6, 85, 186)])

    def make(x0, y0, r0, x1, y1, r1, t, ext):
        rect = Rect(0.0, 0.0, 400.0, 200.0)
        s.fill(FILL_NON_ZERO, t, WHITE, None, rect)
        s.fill(FILL_NON_ZERO, t, Gradient.radial((x0, y0), float(np.float32(r0)), (x1, y1), float(np.float32(r1)), stops, ext), None, rect)
        ra, rb = float(np.float32(r0)) - 1.0, float(np.float32(r1)) - 1.0
        s.stroke(Stroke(1.0), t, BLACK, None, Ellipse(x0, y0, ra, ra, 0.0))
        s.stroke(Stroke(1.0), t, BLACK, None, Ellipse(x1, y1, rb, rb, 0.0))

    exts = (EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT)
    for i, m in enumerate(exts):
        make(140.0, 100.0, 20.0, 280.0, 100.0, 50.0, Affine.translate(i * 420.0 + 20.0, 20.0), m)
    for i, m in enumerate(exts):
        make(280.0, 100.0, 50.0, 140.0, 100.0, 20.0, Affine.translate(i * 420.0 + 20.0, 240.0), m)
    for i, m in enumerate(exts):
        make(140.0, 100.0, 50.0, 280.0, 100.0, 50.0, Affine.translate(i * 420.0 + 20.0, 460.0), m)
    for i, m in enumerate(exts):
        make(140.0, 125.0, 20.0, 190.0, 100.0, 95.0, Affine.translate(i * 420.0 + 20.0, 680.0), m)
    for i, m in enumerate(exts):
        x0, y0, r0, x1, y1, r1 = 140.0, 125.0, 20.0, 190.0, 100.0, 96.0
        dx, dy = x0 - x1, y0 - y1
        n = math.hypot(dx, dy)
        make(x1 + dx / n * (r1 - r0), y1 + dy / n * (r1 - r0), r0, x1, y1, r1, Affine.translate(i * 420.0 + 20.0, 900.0), m)
    return s, 1280, 1120


def many_draw_objects(n_wide: int = 300, n_high: int = 300) -> Tuple[Scene, int, int]:
    """test_scenes.rs:1928-1948: 90,000 small circles (351 draw-object partitions, 90 k paths)."""
    s = Scene()
    W, H = 2000.0, 1500.0
    for j in range(n_high):
        y = (j + 0.5) * (H / n_high)
        for i in range(n_wide):
            s.fill(FILL_NON_ZERO, Affine.IDENTITY, YELLOW, None, Circle((i + 0.5) * (W / n_wide), y, 3.0))
    return s, int(W), int(H)


def conflation_artifacts() -> Tuple[Scene, int, int]:
    """test_scenes.rs:1444-1531: shapes that abut with opposite / equal winding at a fractional pixel offset."""
    s = Scene()
    N, S = 50.0, 4.0
    x, y = N + 0.5, N
    bg, fg = Color.from_rgba8(255, 194, 19), Color.from_rgba8(12, 165, 255)
    sc = Affine.scale(S)
    s.fill(FILL_NON_ZERO, Affine.translate(x, y) * sc, fg, None,
           BezPath([("M", 0.0, 0.0), ("L", N, N), ("L", 0.0, N), ("L", 0.0, 0.0), ("M", 0.0, 0.0), ("L", N, N), ("L", N, 0.0), ("L", 0.0, 0.0)]))
    y += S * N + 10.0
    s.fill(FILL_EVEN_ODD, Affine.translate(x, y) * sc, bg, None, Rect(0.0, 0.0, N, N))
    s.fill(FILL_EVEN_ODD, Affine.translate(x, y) * sc, fg, None,
           BezPath([("M", 0.0, 0.0), ("L", 0.0, N), ("L", N * 0.5, N), ("L", N * 0.5, 0.0),
                    ("M", N * 0.5, 0.0), ("L", N, 0.0), ("L", N, N), ("L", N * 0.5, N)]))
    y += S * N + 10.0
    s.fill(FILL_EVEN_ODD, Affine.translate(x, y) * sc, bg, None, Rect(0.0, 0.0, N, N))
    s.fill(FILL_EVEN_ODD, Affine.translate(x, y) * sc, fg, None,
           BezPath([("M", 0.0, 0.0), ("L", 0.0, N), ("L", N * 0.5, N), ("L", N * 0.5, 0.0),
                    ("M", N * 0.5, 0.0), ("L", N * 0.5, N), ("L", N, N), ("L", N, 0.0)]))
    return s, 320, 720


def image_extend_modes(quality: int = QUALITY_MEDIUM) -> Tuple[Scene, int, int]:
    """test_scenes.rs:2168-2213: a 2x2 image (red, blue / cyan, magenta) drawn 100x with a (2, 2) brush offset under
    pad / reflect / repeat / mixed extends. Base colour white in the reference (`params.base_color`)."""
    px = np.array([[[255, 0, 0, 255], [0, 0, 255, 255]], [[0, 255, 255, 255], [255, 0, 255, 255]]], dtype=np.uint8)
    s = Scene()
    off = Affine.translate(2.0, 2.0)
    for (tx, ty, xe, ye) in ((100.0, 100.0, EXTEND_PAD, EXTEND_PAD), (100.0, 800.0, EXTEND_REFLECT, EXTEND_REFLECT),
                             (800.0, 100.0, EXTEND_REPEAT, EXTEND_REPEAT), (800.0, 800.0, EXTEND_REPEAT, EXTEND_REFLECT)):
        im = Image(px, quality=quality, x_extend=xe, y_extend=ye)
        s.fill(FILL_NON_ZERO, Affine.translate(tx, ty) * Affine.scale(100.0), im, off, Rect(0.0, 0.0, 6.0, 6.0))
    return s, 1500, 1500


def longpathdash(cap: int = STYLE_CAP_BUTT) -> Tuple[Scene, int, int]:
    """test_scenes.rs:779-819: 15 rings of 200 spokes, every spoke a polyline of 21 points, stroked 1 px wide with a
    [1, 1] dash pattern (the dashes are cut on the CPU by kurbo::dash; ~200 k dashes). Lines only: the dash arithmetic is
    closed-form."""
    p = BezPath()
    x = 32
    while x < 256:
        a = 0.0
        while a < math.pi * 2.0:
            p0 = (256.0 + math.sin(a) * x, 256.0 + math.cos(a) * x)
            p1 = (256.0 + math.sin(a + math.pi / 3.0) * (x + 64), 256.0 + math.cos(a + math.pi / 3.0) * (x + 64))
            p.move_to(p0[0], p0[1])
            i = 0.0
            while i < 1.0:
                p.line_to(p0[0] * (1.0 - i) + p1[0] * i, p0[1] * (1.0 - i) + p1[1] * i)
                i += 0.05
            a += math.pi * 0.01
        x += 16
    s = Scene()
    st = Stroke(1.0, join=STYLE_JOIN_BEVEL, start_cap=cap, end_cap=cap, dash_pattern=(1.0, 1.0), dash_offset=0.0)
    s.stroke(st, Affine.translate(50.0, 50.0), YELLOW, None, p)
    return s, 612, 612
