"""Geometry helpers standing in for the parts of `kurbo 0.13.1` the scene builder touches.

kurbo is a third-party dependency that is *not* under /root/reference (Cargo.lock pins
kurbo 0.13.1); what is restated here is its published algorithm for turning shapes into path
elements (`Shape::path_elements(tolerance)`), which `PathEncoder::shape` calls with
tolerance 0.1 (vello_encoding/src/path.rs:655-657):

* Rect            -> M, L, L, L, Z
* Circle          -> cubic arcs, n = 4 with arm 0.551915024494 below the tolerance knee
* RoundedRect     -> quarter-circle arcs joined by lines
* BezPath.from_svg-> SVG path data incl. elliptical arcs -> cubics (tolerance 0.1)

All maths is f64 like kurbo; the encoder rounds to f32.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Sequence, Tuple


@dataclass(frozen=True)
class Affine:
    """Column-major 2x3 affine [a, b, c, d, e, f]: x' = a x + c y + e, y' = b x + d y + f."""

    coeffs: Tuple[float, float, float, float, float, float] = (1.0, 0.0, 0.0, 1.0, 0.0, 0.0)

    IDENTITY = None  # filled below

    @staticmethod
    def translate(x: float, y: float) -> "Affine":
        return Affine((1.0, 0.0, 0.0, 1.0, float(x), float(y)))

    @staticmethod
    def scale(s: float, sy: float | None = None) -> "Affine":
        return Affine((float(s), 0.0, 0.0, float(s if sy is None else sy), 0.0, 0.0))

    @staticmethod
    def rotate(th: float) -> "Affine":
        s, c = math.sin(th), math.cos(th)
        return Affine((c, s, -s, c, 0.0, 0.0))

    def __mul__(self, o: "Affine") -> "Affine":
        a = self.coeffs
        b = o.coeffs
        return Affine(
            (
                a[0] * b[0] + a[2] * b[1],
                a[1] * b[0] + a[3] * b[1],
                a[0] * b[2] + a[2] * b[3],
                a[1] * b[2] + a[3] * b[3],
                a[0] * b[4] + a[2] * b[5] + a[4],
                a[1] * b[4] + a[3] * b[5] + a[5],
            )
        )

    def apply(self, x: float, y: float) -> Tuple[float, float]:
        a = self.coeffs
        return (a[0] * x + a[2] * y + a[4], a[1] * x + a[3] * y + a[5])


Affine.IDENTITY = Affine()


@dataclass(frozen=True)
class Rect:
    x0: float
    y0: float
    x1: float
    y1: float

    @staticmethod
    def from_center_size(c, s) -> "Rect":
        return Rect(c[0] - 0.5 * s[0], c[1] - 0.5 * s[1], c[0] + 0.5 * s[0], c[1] + 0.5 * s[1])

    @staticmethod
    def from_origin_size(o, s) -> "Rect":
        return Rect(o[0], o[1], o[0] + s[0], o[1] + s[1])


@dataclass(frozen=True)
class Circle:
    cx: float
    cy: float
    r: float


@dataclass(frozen=True)
class Ellipse:
    """kurbo `Ellipse::new(center, radii, x_rotation)` (ellipse.rs): path = MoveTo(start) + `Arc` over 2 pi + ClosePath."""
    cx: float
    cy: float
    rx: float
    ry: float
    x_rotation: float = 0.0


@dataclass(frozen=True)
class RoundedRect:
    x0: float
    y0: float
    x1: float
    y1: float
    radius: float


@dataclass(frozen=True)
class Line:
    x0: float
    y0: float
    x1: float
    y1: float


class BezPath:
    """A list of path elements: ("M",x,y) ("L",x,y) ("Q",x1,y1,x,y) ("C",x1,y1,x2,y2,x,y) ("Z",)."""

    def __init__(self, els: Iterable[tuple] = ()):
        self.els: List[tuple] = list(els)

    def move_to(self, x, y):
        self.els.append(("M", float(x), float(y)))

    def line_to(self, x, y):
        self.els.append(("L", float(x), float(y)))

    def quad_to(self, x1, y1, x, y):
        self.els.append(("Q", float(x1), float(y1), float(x), float(y)))

    def curve_to(self, x1, y1, x2, y2, x, y):
        self.els.append(("C", float(x1), float(y1), float(x2), float(y2), float(x), float(y)))

    def close_path(self):
        self.els.append(("Z",))

    @staticmethod
    def from_svg(d: str) -> "BezPath":
        return BezPath(parse_svg_path(d))


def _arc_elements(cx, cy, rx, ry, start, sweep, x_rot, tolerance) -> Iterator[tuple]:
    """kurbo `Arc::append_iter`: n cubic pieces, n from the tolerance; arm = 4/3 tan(sweep/4n)."""
    sign = 1.0 if sweep >= 0 else -1.0
    scaled_err = max(rx, ry) / tolerance
    n_err = max((1.1163 * scaled_err) ** (1.0 / 6.0), 3.999_999)
    n = int(math.ceil(n_err * abs(sweep) * (1.0 / (2.0 * math.pi))))
    n = max(n, 1)
    angle_step = sweep / n
    arm_len = (4.0 / 3.0) * abs(math.tan(0.25 * angle_step)) * sign
    cr, sr = math.cos(x_rot), math.sin(x_rot)

    def sample(a):
        x, y = rx * math.cos(a), ry * math.sin(a)
        return (cr * x - sr * y, sr * x + cr * y)

    angle0 = start
    p0 = sample(angle0)
    for _ in range(n):
        angle1 = angle0 + angle_step
        p1 = (p0[0] - arm_len * _rot_d(rx, ry, angle0, cr, sr)[0], p0[1] - arm_len * _rot_d(rx, ry, angle0, cr, sr)[1])
        p3 = sample(angle1)
        d1 = _rot_d(rx, ry, angle1, cr, sr)
        p2 = (p3[0] + arm_len * d1[0], p3[1] + arm_len * d1[1])
        yield ("C", cx + p1[0], cy + p1[1], cx + p2[0], cy + p2[1], cx + p3[0], cy + p3[1])
        angle0, p0 = angle1, p3


def _rot_d(rx, ry, a, cr, sr):
    # rotated (rx sin a, -ry cos a): minus the derivative of the ellipse sample
    x, y = rx * math.sin(a), -ry * math.cos(a)
    return (cr * x - sr * y, sr * x + cr * y)


def path_elements(shape, tolerance: float = 0.1) -> Iterator[tuple]:
    if isinstance(shape, BezPath):
        yield from shape.els
    elif isinstance(shape, (list, tuple)):
        yield from shape
    elif isinstance(shape, Rect):
        yield ("M", shape.x0, shape.y0)
        yield ("L", shape.x1, shape.y0)
        yield ("L", shape.x1, shape.y1)
        yield ("L", shape.x0, shape.y1)
        yield ("Z",)
    elif isinstance(shape, Line):
        yield ("M", shape.x0, shape.y0)
        yield ("L", shape.x1, shape.y1)
    elif isinstance(shape, Circle):
        r = abs(shape.r)
        scaled_err = r / tolerance
        if scaled_err < 1.0 / 1.9608e-4:
            n, arm = 4, 0.551915024494
        else:
            n = int(math.ceil((1.1163 * scaled_err) ** (1.0 / 6.0)))
            arm = (4.0 / 3.0) * math.tan(math.pi / (2.0 * n))
        x, y = shape.cx, shape.cy
        yield ("M", x + r, y)
        dth = 2.0 * math.pi / n
        for ix in range(1, n + 1):
            th1 = dth * ix
            th0 = th1 - dth
            s0, c0 = math.sin(th0), math.cos(th0)
            if ix == n:
                s1, c1 = 0.0, 1.0
            else:
                s1, c1 = math.sin(th1), math.cos(th1)
            a = arm * r
            yield (
                "C",
                x + r * c0 - a * s0, y + r * s0 + a * c0,
                x + r * c1 + a * s1, y + r * s1 - a * c1,
                x + r * c1, y + r * s1,
            )
        yield ("Z",)
    elif isinstance(shape, Ellipse):
        # Ellipse::path_elements: radii / rotation come back out of the affine through `svd` (kurbo affine.rs), then
        # Arc { start 0, sweep 2 pi }.path_elements(tolerance) chained with ClosePath
        a, b = shape.rx * math.cos(shape.x_rotation), shape.rx * math.sin(shape.x_rotation)
        c, d = -shape.ry * math.sin(shape.x_rotation), shape.ry * math.cos(shape.x_rotation)
        a2, b2, c2, d2 = a * a, b * b, c * c, d * d
        rot = 0.5 * math.atan2(2.0 * (a * c + b * d), a2 - b2 + c2 - d2)
        s1 = a2 + b2 + c2 + d2
        s2 = math.sqrt((a2 - b2 + c2 - d2) ** 2 + 4.0 * (a * c + b * d) ** 2)
        rx, ry = math.sqrt(0.5 * (s1 + s2)), math.sqrt(max(0.5 * (s1 - s2), 0.0))
        cr, sr = math.cos(rot), math.sin(rot)
        yield ("M", shape.cx + cr * rx, shape.cy + sr * rx)
        yield from _arc_elements(shape.cx, shape.cy, rx, ry, 0.0, 2.0 * math.pi, rot, tolerance)
        yield ("Z",)
    elif isinstance(shape, RoundedRect):
        x0, y0, x1, y1 = shape.x0, shape.y0, shape.x1, shape.y1
        rad = min(abs(shape.radius), 0.5 * abs(x1 - x0), 0.5 * abs(y1 - y0))
        if rad <= 0.0:
            yield from path_elements(Rect(x0, y0, x1, y1), tolerance)
            return
        hp = 0.5 * math.pi
        # start at the top edge after the top-left corner, go clockwise (y down)
        yield ("M", x0 + rad, y0)
        corners = [
            (x1 - rad, y0 + rad, -hp),  # top-right: from angle -90 to 0
            (x1 - rad, y1 - rad, 0.0),  # bottom-right
            (x0 + rad, y1 - rad, hp),  # bottom-left
            (x0 + rad, y0 + rad, 2 * hp),  # top-left
        ]
        for (cx, cy, a0) in corners:
            yield ("L", cx + rad * math.cos(a0), cy + rad * math.sin(a0))
            yield from _arc_elements(cx, cy, rad, rad, a0, hp, 0.0, tolerance)
        yield ("Z",)
    else:
        raise TypeError(f"unsupported shape {type(shape)}")


# ---------------------------------------------------------------------------------------------
# SVG path data (the subset kurbo's `BezPath::from_svg` accepts: MmLlHhVvCcSsQqTtAaZz)
# ---------------------------------------------------------------------------------------------
_NUM = re.compile(r"[+-]?(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?")


class _Lexer:
    def __init__(self, s: str):
        self.s = s
        self.i = 0

    def skip(self):
        while self.i < len(self.s) and self.s[self.i] in " \t\r\n,":
            self.i += 1

    def peek_cmd(self):
        self.skip()
        if self.i < len(self.s) and self.s[self.i].isalpha():
            return self.s[self.i]
        return None

    def more_numbers(self) -> bool:
        self.skip()
        return self.i < len(self.s) and (self.s[self.i] in "+-." or self.s[self.i].isdigit())

    def num(self) -> float:
        self.skip()
        m = _NUM.match(self.s, self.i)
        if not m:
            raise ValueError(f"bad number at {self.i}: {self.s[self.i:self.i+16]!r}")
        self.i = m.end()
        return float(m.group(0))

    def flag(self) -> bool:
        self.skip()
        c = self.s[self.i]
        if c not in "01":
            raise ValueError("bad arc flag")
        self.i += 1
        return c == "1"


def _svg_arc_to_cubics(x0, y0, rx, ry, x_rot_deg, large, sweep, x, y, tolerance=0.1):
    """SVG implementation notes F.6.5 (endpoint -> centre) then `_arc_elements`."""
    if rx == 0.0 or ry == 0.0 or (x0 == x and y0 == y):
        if not (x0 == x and y0 == y):
            yield ("L", x, y)
        return
    rx, ry = abs(rx), abs(ry)
    phi = math.radians(x_rot_deg)
    cp, sp = math.cos(phi), math.sin(phi)
    dx2, dy2 = 0.5 * (x0 - x), 0.5 * (y0 - y)
    x1p = cp * dx2 + sp * dy2
    y1p = -sp * dx2 + cp * dy2
    lam = (x1p * x1p) / (rx * rx) + (y1p * y1p) / (ry * ry)
    if lam > 1.0:
        s = math.sqrt(lam)
        rx *= s
        ry *= s
    num = rx * rx * ry * ry - rx * rx * y1p * y1p - ry * ry * x1p * x1p
    den = rx * rx * y1p * y1p + ry * ry * x1p * x1p
    coef = math.sqrt(max(num / den, 0.0)) if den != 0.0 else 0.0
    if large == sweep:
        coef = -coef
    cxp = coef * rx * y1p / ry
    cyp = -coef * ry * x1p / rx
    cx = cp * cxp - sp * cyp + 0.5 * (x0 + x)
    cy = sp * cxp + cp * cyp + 0.5 * (y0 + y)
    a0 = math.atan2((y1p - cyp) / ry, (x1p - cxp) / rx)
    a1 = math.atan2((-y1p - cyp) / ry, (-x1p - cxp) / rx)
    d = a1 - a0
    if sweep and d < 0:
        d += 2 * math.pi
    elif not sweep and d > 0:
        d -= 2 * math.pi
    els = list(_arc_elements(cx, cy, rx, ry, a0, d, phi, tolerance))
    if els:
        last = els[-1]
        els[-1] = ("C", last[1], last[2], last[3], last[4], x, y)  # land exactly on the endpoint
    yield from els


def parse_svg_path(d: str) -> List[tuple]:
    lx = _Lexer(d)
    els: List[tuple] = []
    cx = cy = 0.0  # current point
    sx = sy = 0.0  # subpath start
    last_ctrl = None  # for S/T reflection
    last_cmd = ""
    cmd = None
    while True:
        c = lx.peek_cmd()
        if c is not None:
            cmd = c
            lx.i += 1
        elif not lx.more_numbers():
            break
        elif cmd is None:
            raise ValueError("path data must start with a command")
        elif cmd in "Mm":
            cmd = "L" if cmd == "M" else "l"  # implicit line-to after move-to
        rel = cmd.islower()
        u = cmd.upper()
        if u == "Z":
            els.append(("Z",))
            cx, cy = sx, sy
            last_ctrl = None
            last_cmd = u
            if lx.more_numbers():
                raise ValueError("numbers after close-path")
            continue
        if u == "M":
            x, y = lx.num(), lx.num()
            if rel:
                x, y = cx + x, cy + y
            els.append(("M", x, y))
            cx, cy = sx, sy = x, y
            last_ctrl = None
        elif u == "L":
            x, y = lx.num(), lx.num()
            if rel:
                x, y = cx + x, cy + y
            els.append(("L", x, y))
            cx, cy = x, y
            last_ctrl = None
        elif u == "H":
            x = lx.num()
            if rel:
                x += cx
            els.append(("L", x, cy))
            cx = x
            last_ctrl = None
        elif u == "V":
            y = lx.num()
            if rel:
                y += cy
            els.append(("L", cx, y))
            cy = y
            last_ctrl = None
        elif u == "C":
            v = [lx.num() for _ in range(6)]
            if rel:
                v = [v[0] + cx, v[1] + cy, v[2] + cx, v[3] + cy, v[4] + cx, v[5] + cy]
            els.append(("C", *v))
            last_ctrl = (v[2], v[3])
            cx, cy = v[4], v[5]
        elif u == "S":
            v = [lx.num() for _ in range(4)]
            if rel:
                v = [v[0] + cx, v[1] + cy, v[2] + cx, v[3] + cy]
            if last_cmd in ("C", "S") and last_ctrl is not None:
                x1, y1 = 2 * cx - last_ctrl[0], 2 * cy - last_ctrl[1]
            else:
                x1, y1 = cx, cy
            els.append(("C", x1, y1, v[0], v[1], v[2], v[3]))
            last_ctrl = (v[0], v[1])
            cx, cy = v[2], v[3]
        elif u == "Q":
            v = [lx.num() for _ in range(4)]
            if rel:
                v = [v[0] + cx, v[1] + cy, v[2] + cx, v[3] + cy]
            els.append(("Q", *v))
            last_ctrl = (v[0], v[1])
            cx, cy = v[2], v[3]
        elif u == "T":
            v = [lx.num() for _ in range(2)]
            if rel:
                v = [v[0] + cx, v[1] + cy]
            if last_cmd in ("Q", "T") and last_ctrl is not None:
                x1, y1 = 2 * cx - last_ctrl[0], 2 * cy - last_ctrl[1]
            else:
                x1, y1 = cx, cy
            els.append(("Q", x1, y1, v[0], v[1]))
            last_ctrl = (x1, y1)
            cx, cy = v[0], v[1]
        elif u == "A":
            rx, ry, rot = lx.num(), lx.num(), lx.num()
            large, sweep = lx.flag(), lx.flag()
            x, y = lx.num(), lx.num()
            if rel:
                x, y = cx + x, cy + y
            els.extend(_svg_arc_to_cubics(cx, cy, rx, ry, rot, large, sweep, x, y))
            cx, cy = x, y
            last_ctrl = None
        else:
            raise ValueError(f"unsupported path command {cmd!r}")
        last_cmd = u
    return els
