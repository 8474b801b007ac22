"""CPU-side check of the guard behind k_flatten's stroked-line fast path (vello_b200/csrc/k_flatten.cu, flatten_tag):
whenever the guard holds, the reference algorithm (the oracle's Euler flattening, which has no shortcut) must emit
exactly ONE line per side of a stroked line-to. The guard is restated here in float32; the GPU test
test_flatten_stroke_line_fast_path_extremes then compares the kernel's `lines` with the oracle's bit for bit."""
import numpy as np
import pytest

from vello_b200.encoding import STYLE_CAP_BUTT, STYLE_JOIN_BEVEL, Color, Scene, Stroke, resolve
from vello_b200.shapes import Affine, BezPath

f32 = np.float32


def guard(p0, p3, half_width, coeffs) -> bool:
    """Mirror of the condition in k_flatten.cu (float32 arithmetic, same operation order)."""
    chx, chy = f32(p3[0] - p0[0]), f32(p3[1] - p0[1])
    c2 = f32(f32(chx * chx) + f32(chy * chy))
    mag = max(abs(p0[0]), abs(p0[1]), abs(p3[0]), abs(p3[1]))
    m0, m1, m2, m3 = (f32(v) for v in coeffs[:4])
    s1 = f32(f32(f32(abs(m0) + abs(m1)) + abs(m2)) + abs(m3))
    u = f32(mag * f32(2.3841858e-07))
    return bool(half_width > 0 and f32(mag * s1) < f32(16384.0) and c2 >= f32(f32(256.0) * u) * u and c2 >= f32(u * half_width)
                and c2 >= f32(1e-10) and c2 < f32(1e30))


@pytest.mark.parametrize("boundary", [False, True])
def test_stroked_line_guard_implies_one_line_per_side(oracle, boundary):
    rng = np.random.default_rng(77 + boundary)
    n = 6000
    s = Scene()
    cases = []
    for k in range(n):
        mag = float(10 ** rng.uniform(-3, 4.2))
        x0, y0 = rng.uniform(-mag, mag, 2)
        width = float(10 ** rng.uniform(-3, 3.5))
        c = float(10 ** rng.uniform(-5.5, 3))
        if boundary:  # chords just around the guard's thresholds
            u = mag * 2.4e-7
            c = float(np.sqrt(max(256 * u * u, u * 0.5 * width, 1e-10)) * rng.uniform(0.9, 2.5))
        th = rng.uniform(0, 2 * np.pi)
        x1, y1 = x0 + c * np.cos(th), y0 + c * np.sin(th)
        if f32(x0) == f32(x1) and f32(y0) == f32(y1):
            x1 = x0 + max(abs(x0), 1e-3) * 1e-3
        sc = float(10 ** rng.uniform(-2, 2))
        t = Affine(tuple(sc * rng.normal(size=4)) + (10.0, -7.0)) if k % 3 == 0 else Affine.scale(sc)
        p = BezPath()
        p.move_to(x0, y0)
        p.line_to(x1, y1)
        s.stroke(Stroke(width, join=STYLE_JOIN_BEVEL, start_cap=STYLE_CAP_BUTT, end_cap=STYLE_CAP_BUTT), t, Color.from_rgba8(1, 2, 3), None, p)
        cases.append(((f32(x0), f32(y0)), (f32(x1), f32(y1)), f32(0.5) * f32(width), t.coeffs))
    packed = resolve(s.encoding)
    assert packed.layout.n_paths == n
    oracle.bind(packed, 256, 256)
    oracle.run("pathtag", "flatten")
    per_path = np.bincount(oracle.buffer("lines")["path_ix"], minlength=n)
    taken = [k for k, cse in enumerate(cases) if guard(*cse)]
    assert len(taken) > n // 3
    # open single-segment stroke with butt caps: two side lines + end cap + start cap
    wrong = [k for k in taken if per_path[k] != 4]
    assert not wrong, (len(wrong), wrong[:5])


def test_filled_line_guard_implies_one_line(oracle):
    """The fill fast path: a line-to of a fill whose device-space end points are below 65536 in magnitude yields
    exactly one line in the reference algorithm (or none, when degenerate)."""
    from vello_b200.encoding import FILL_NON_ZERO
    rng = np.random.default_rng(5)
    n = 8000
    s = Scene()
    expect = []
    for k in range(n):
        mag = float(10 ** rng.uniform(-2, 4.7))
        step = float(10 ** rng.uniform(-5, 4))
        pts = [rng.uniform(-mag, mag, 2)]
        for _ in range(2):
            nxt = pts[-1] + rng.normal(0, step, 2)
            while any(f32(nxt[0]) == f32(o[0]) and f32(nxt[1]) == f32(o[1]) for o in pts):  # the encoder drops empty segments
                nxt = nxt + np.maximum(np.abs(nxt), 1e-3) * 1e-3
            pts.append(nxt)
        p = BezPath()
        p.move_to(*pts[0])
        p.line_to(*pts[1])
        p.line_to(*pts[2])
        p.close_path()
        s.fill(FILL_NON_ZERO, Affine.IDENTITY, Color.from_rgba8(9, 9, 9), None, p)
        q = [(f32(a), f32(b)) for a, b in pts]
        inside = all(abs(v) < 65536.0 for pt in q for v in pt)
        segs = sum(1 for i in range(3) if q[i] != q[(i + 1) % 3])  # zero-length segments emit nothing
        expect.append((inside, segs))
    packed = resolve(s.encoding)
    assert packed.layout.n_paths == n
    oracle.bind(packed, 256, 256)
    oracle.run("pathtag", "flatten")
    per_path = np.bincount(oracle.buffer("lines")["path_ix"], minlength=n)
    wrong = [k for k, (inside, segs) in enumerate(expect) if inside and per_path[k] != segs]
    assert sum(1 for e in expect if e[0]) > n // 2
    assert not wrong, (len(wrong), wrong[:5], [expect[k] for k in wrong[:5]], [int(per_path[k]) for k in wrong[:5]])
