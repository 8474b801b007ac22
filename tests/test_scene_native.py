"""The native scene front end (include/vello_b200_scene.h, vb_scene.cpp) against the Python statement of the encoder
(vello_b200/encoding.py, which the reference's golden images pin through the renderer): every `Scene` call of every
test recipe is issued to both, and the resolved packed scene, layout, gradient ramps and image atlas must be identical
byte for byte. CPU only (no GPU call is made)."""
import numpy as np
import pytest

from vello_b200 import encoding, scenes
from vello_b200.encoding import resolve
from vello_b200.scene_native import NativeScene, SCENE_SYMBOLS


class MirrorScene(encoding.Scene):
    """A Scene that repeats every high-level call on a NativeScene."""

    def __init__(self):
        super().__init__()
        self.native = NativeScene()

    def fill(self, *a):
        super().fill(*a)
        self.native.fill(*a)

    def stroke(self, *a):
        super().stroke(*a)
        self.native.stroke(*a)

    def push_layer(self, *a):
        super().push_layer(*a)
        self.native.push_layer(*a)

    def push_luminance_mask_layer(self, *a):
        super().push_luminance_mask_layer(*a)
        self.native.push_luminance_mask_layer(*a)

    def push_clip_layer(self, *a):
        super().push_clip_layer(*a)
        self.native.push_clip_layer(*a)

    def pop_layer(self):
        super().pop_layer()
        self.native.pop_layer()

    def draw_image(self, image, transform):
        # Scene.draw_image is fill() with an image brush in both implementations; call the dedicated native entry point
        encoding.Scene.fill(self, encoding.FILL_NON_ZERO, transform, image, None,
                            scenes._shapes.Rect(0.0, 0.0, float(image.width), float(image.height)) if hasattr(scenes, "_shapes") else
                            __import__("vello_b200.shapes", fromlist=["Rect"]).Rect(0.0, 0.0, float(image.width), float(image.height)))
        self.native.draw_image(image, transform)

    def draw_blurred_rounded_rect(self, *a):
        self._quiet = True  # the Python method is a wrapper around ..._in; the native side has its own entry point
        try:
            encoding.Scene.draw_blurred_rounded_rect(self, *a)
        finally:
            self._quiet = False
        self.native.draw_blurred_rounded_rect(*a)

    def draw_blurred_rounded_rect_in(self, *a):
        encoding.Scene.draw_blurred_rounded_rect_in(self, *a)
        if not getattr(self, "_quiet", False):
            self.native.draw_blurred_rounded_rect_in(*a)


@pytest.fixture()
def mirror(monkeypatch):
    monkeypatch.setattr(scenes, "Scene", MirrorScene)
    return None


def assert_same(scene: MirrorScene):
    a = resolve(scene.encoding)
    b = scene.native.resolve()
    assert a.layout == b.layout, (a.layout, b.layout)
    assert a.scene.shape == b.scene.shape
    assert a.scene.tobytes() == b.scene.tobytes(), f"packed scene differs at words {np.nonzero(a.scene != b.scene)[0][:8]}"
    assert a.ramps.shape == b.ramps.shape and a.ramps.tobytes() == b.ramps.tobytes()
    assert a.atlas.shape == b.atlas.shape and a.atlas.tobytes() == b.atlas.tobytes()


def test_symbols_exported():
    import ctypes
    from vello_b200.renderer import load_library
    lib = load_library()
    for s in SCENE_SYMBOLS:
        assert hasattr(lib, s) and isinstance(getattr(lib, s), ctypes._CFuncPtr), s


RECIPES = ["filled_square", "filled_circle", "simple_square", "layer_size", "robust_paths", "funky_paths", "fill_types", "stroke_styles",
           "many_clips", "deep_blend", "brushes"]


@pytest.mark.parametrize("name", RECIPES)
def test_recipe_bytes_identical(mirror, name):
    s = getattr(scenes, name)()[0]
    assert isinstance(s, MirrorScene)
    assert_same(s)


@pytest.mark.parametrize("premul", [True, False])
def test_gradient_recipes(mirror, premul):
    assert_same(scenes.gradient_color_alpha(premul)[0])


@pytest.mark.parametrize("extend", [encoding.EXTEND_PAD, encoding.EXTEND_REPEAT, encoding.EXTEND_REFLECT])
def test_image_recipe(mirror, extend):
    img = (np.arange(16 * 12 * 4, dtype=np.uint32) * 37 % 256).astype(np.uint8).reshape(12, 16, 4)
    assert_same(scenes.image_roundtrip(img, extend)[0])


@pytest.mark.parametrize("seed", range(10))
def test_random_scenes(mirror, seed):
    assert_same(scenes.random_small(seed)[0])


def test_tiger(mirror):
    assert_same(scenes.tiger(512, 512))


def test_open_clips_and_empty_scene(mirror):
    from vello_b200.shapes import Affine, Rect, Circle
    s = MirrorScene()
    assert_same(s)  # empty
    s.push_clip_layer(encoding.FILL_NON_ZERO, Affine.IDENTITY, Rect(0, 0, 50, 50))
    s.push_layer(encoding.Stroke(3.0), encoding.MIX_MULTIPLY, encoding.COMPOSE_SRC_OVER, 0.5, Affine.scale(2.0), Circle(10.0, 10.0, 8.0))
    s.fill(encoding.FILL_EVEN_ODD, Affine.IDENTITY, encoding.Color(0.2, 0.4, 0.6, 0.8), Affine.rotate(0.3), Circle(5.0, 5.0, 4.0))
    s.push_luminance_mask_layer(encoding.FILL_NON_ZERO, 0.25, Affine.IDENTITY, Rect(1, 2, 3, 4))
    assert_same(s)  # three layers left open: trailing PATH tags / END_CLIP draw tags
    s.pop_layer()
    s.pop_layer()
    s.pop_layer()
    s.pop_layer()  # one too many: ignored
    s.stroke(encoding.Stroke(0.0), Affine.IDENTITY, encoding.RED, None, Rect(0, 0, 1, 1))  # zero width: nothing
    s.draw_blurred_rounded_rect(Affine.translate(3.0, 4.0), Rect(10, 10, 60, 40), encoding.Color(0.9, 0.1, 0.1, 0.7), 6.0, 3.5)
    s.draw_blurred_rounded_rect_in(Circle(40.0, 30.0, 25.0), Affine.rotate(0.2), Rect(10, 10, 60, 40), encoding.Color(0.2, 0.1, 0.9, 1.0), 3.0, 2.0)
    assert_same(s)


def test_c_example_builds(tmp_path):
    """examples/native_demo.c uses nothing but the two C headers and libvello_b200.so."""
    import os, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    exe = tmp_path / "native_demo"
    subprocess.run([cc, "-O1", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "native_demo.c"),
                    "-L" + os.path.join(root, "vello_b200"), "-lvello_b200", "-Wl,-rpath," + os.path.join(root, "vello_b200"), "-lm",
                    "-o", str(exe)], check=True)
    assert exe.exists()


@pytest.mark.parametrize("seed", range(8))
def test_path_encoder_state_machine_fuzz(seed):
    """Degenerate element sequences -- repeated move-tos, closes with nothing to close, zero-length and near-zero
    segments, paths that start with a curve, empty paths, several subpaths -- drive every branch of the PathEncoder state
    machine (path.rs:425-838), for fills and for strokes (cap markers), in both implementations."""
    from vello_b200.shapes import Affine, BezPath
    rng = np.random.default_rng(1000 + seed)
    s = MirrorScene()
    for k in range(120):
        p = BezPath()
        last = [float(v) for v in rng.integers(0, 50, 2)]
        for _ in range(int(rng.integers(0, 9))):
            op = int(rng.integers(0, 8))
            def pt(repeat_prob=0.35):
                nonlocal last
                if rng.random() < repeat_prob:
                    return list(last)
                if rng.random() < 0.2:  # within float noise of the last point
                    return [last[0] + float(rng.choice([0.0, 1e-13, 1e-7])), last[1]]
                last = [float(v) for v in rng.uniform(-20, 80, 2)]
                return list(last)
            if op == 0:
                p.move_to(*pt())
            elif op in (1, 2):
                p.line_to(*pt())
            elif op == 3:
                p.quad_to(*pt(), *pt())
            elif op == 4:
                p.curve_to(*pt(), *pt(), *pt())
            elif op == 5:
                p.close_path()
            elif op == 6:
                p.move_to(*pt())
                p.move_to(*pt())
            else:
                p.close_path()
                p.line_to(*pt())
        t = Affine.translate(float(k % 5), 0.0) if k % 3 else Affine.IDENTITY
        col = encoding.Color.from_rgba8(*[int(v) for v in rng.integers(0, 256, 4)])
        if k % 2:
            s.fill(encoding.FILL_NON_ZERO if k % 4 == 1 else encoding.FILL_EVEN_ODD, t, col, None, p)
        else:
            st = encoding.Stroke(float(rng.choice([0.0, 0.5, 3.0])), join=(encoding.STYLE_JOIN_BEVEL, encoding.STYLE_JOIN_MITER, encoding.STYLE_JOIN_ROUND)[k % 3],
                                 start_cap=(encoding.STYLE_CAP_BUTT, encoding.STYLE_CAP_SQUARE, encoding.STYLE_CAP_ROUND)[(k // 2) % 3],
                                 end_cap=(encoding.STYLE_CAP_BUTT, encoding.STYLE_CAP_SQUARE, encoding.STYLE_CAP_ROUND)[(k // 3) % 3],
                                 miter_limit=float(rng.choice([1.0, 4.0, 1e4])))
            s.stroke(st, t, col, None, p)
        if k % 17 == 0:
            s.push_clip_layer(encoding.FILL_NON_ZERO, t, p)  # clip with a possibly empty shape -> the empty-path sentinel
        if k % 17 == 9:
            s.pop_layer()
    assert_same(s)


def test_append(mirror, oracle):
    """Scene::append: fragments (with gradients and images, i.e. late-bound patches) appended with and without a
    transform; native == Python, and a translated fragment renders like the same content drawn at the translated place."""
    from vello_b200.shapes import Affine, Rect, Circle
    frag = scenes.brushes()[0]
    frag2 = scenes.stroke_styles()[0]
    s = MirrorScene()
    s.fill(encoding.FILL_NON_ZERO, Affine.IDENTITY, encoding.Color(0.1, 0.1, 0.1, 1.0), None, Rect(0, 0, 900, 900))
    for sc in (s,):
        encoding.Scene.append(sc, frag, None)
        sc.native.append(frag.native, None)
        encoding.Scene.append(sc, frag2, Affine.translate(450.0, 20.0) * Affine.scale(0.5))
        sc.native.append(frag2.native, Affine.translate(450.0, 20.0) * Affine.scale(0.5))
        encoding.Scene.append(sc, frag, Affine((0.5, 0.1, -0.1, 0.5, 30.0, 470.0)))
        sc.native.append(frag.native, Affine((0.5, 0.1, -0.1, 0.5, 30.0, 470.0)))
    s.fill(encoding.FILL_NON_ZERO, Affine.IDENTITY, encoding.Color(1.0, 1.0, 1.0, 0.5), None, Circle(400.0, 400.0, 50.0))
    assert_same(s)
    # semantics: appending with an integral translation == drawing the same shapes translated (all f32 products exact)
    part = encoding.Scene()
    part.fill(encoding.FILL_NON_ZERO, Affine.IDENTITY, encoding.Color(0.9, 0.2, 0.1, 0.8), None, Circle(40.0, 40.0, 30.0))
    part.stroke(encoding.Stroke(4.0), Affine.translate(3.0, 5.0), encoding.Color(0.1, 0.9, 0.3, 1.0), None, Rect(10, 10, 90, 60))
    whole = encoding.Scene()
    whole.append(part, Affine.translate(64.0, 32.0))
    direct = encoding.Scene()
    direct.fill(encoding.FILL_NON_ZERO, Affine.translate(64.0, 32.0), encoding.Color(0.9, 0.2, 0.1, 0.8), None, Circle(40.0, 40.0, 30.0))
    direct.stroke(encoding.Stroke(4.0), Affine.translate(67.0, 37.0), encoding.Color(0.1, 0.9, 0.3, 1.0), None, Rect(10, 10, 90, 60))
    from vello_b200.config import AA_MSAA16
    a = oracle.render(resolve(whole.encoding), 200, 128, encoding.BLACK.premul_rgba8_u32(), AA_MSAA16)
    b = oracle.render(resolve(direct.encoding), 200, 128, encoding.BLACK.premul_rgba8_u32(), AA_MSAA16)
    assert np.array_equal(a, b) and a[..., :3].max() > 100


def test_native_shapes_equal_python_statement_of_kurbo():
    """vb_pathbuf's Rect / Line / Circle / RoundedRect / Ellipse / Arc -> Bezier conversions (kurbo `Shape::path_elements`) produce exactly
    the doubles vello_b200.shapes does, over radii that take every branch of the subdivision-count formulas."""
    from vello_b200.scene_native import NativePath, PATHBUF_SYMBOLS
    from vello_b200.renderer import load_library
    from vello_b200.shapes import Arc, Circle, Ellipse, Line, Rect, RoundedRect, BezPath, path_elements
    lib = load_library()
    for s in PATHBUF_SYMBOLS:
        assert hasattr(lib, s), s
    rng = np.random.default_rng(9)
    shapes = [Rect(1.5, -2.0, 30.25, 17.0), Line(0.0, 1.0, -5.0, 9.5), RoundedRect(0, 0, 10, 10, 0.0), RoundedRect(0, 0, 10, 4, 50.0)]
    for _ in range(200):
        r = float(10 ** rng.uniform(-2, 6))
        shapes.append(Circle(float(rng.normal(0, 100)), float(rng.normal(0, 100)), r * float(rng.choice([1.0, -1.0]))))
        x0, y0 = (float(v) for v in rng.normal(0, 50, 2))
        w, h = (float(v) for v in 10 ** rng.uniform(-1, 5, 2))
        shapes.append(RoundedRect(x0, y0, x0 + w, y0 + h, float(10 ** rng.uniform(-2, 5))))
        rx, ry = (float(v) for v in 10 ** rng.uniform(-1, 4, 2))
        shapes.append(Ellipse(x0, y0, rx, ry, float(rng.uniform(-4, 4))))
        shapes.append(Arc(x0, y0, rx, ry, float(rng.uniform(-7, 7)), float(rng.uniform(-7, 7)), float(rng.uniform(-4, 4))))
    for tol in (0.1, 0.01, 3.0):
        for sh in shapes:
            want = [tuple(float(v) if not isinstance(v, str) else v for v in e) for e in path_elements(sh, tol)]
            got = NativePath().add(sh, tol).elements()
            assert got == want, (sh, tol)
    p = BezPath()
    p.move_to(1, 2); p.quad_to(3, 4, 5, 6); p.curve_to(7, 8, 9, 10, 11, 12); p.line_to(0, 0); p.close_path()
    assert NativePath().add(p).elements() == [tuple(float(v) if not isinstance(v, str) else v for v in e) for e in p.els]


def test_native_svg_path_parser():
    """vb_pathbuf_svg == vello_b200.shapes.parse_svg_path (kurbo BezPath::from_svg): every Ghostscript-tiger path string,
    and synthetic data using every command, relative forms, implicit repeats, packed flags and numbers, arcs of all four
    flag combinations, degenerate arcs; malformed data is rejected by both."""
    import gzip, json, os
    from vello_b200.scene_native import NativePath
    from vello_b200.shapes import parse_svg_path
    root = os.path.dirname(os.path.abspath(__file__))
    tiger = json.load(gzip.open(os.path.join(root, "golden", "tiger_paths.json.gz")))
    strings = [it["d"] for it in tiger["items"]]
    strings += [
        "M10 20L30 40H50V60h-5v-5l1 1z",
        "m1,2 3,4 5,6 z m10 10 l1 1",
        "M0 0C1 2 3 4 5 6S9 10 11 12s1 1 2 2Q1 1 2 2T5 5t1 1 2 2",
        "M10-20.5.5-3e1,4E-1 1e+2.25.75",
        "M0 0A10 20 30 0 1 50 60a5 5 0 1 0 10 10 5 5 0 0110 10A5 5 0 11 80 80A0 5 0 0 0 90 90A5 5 0 0 0 90 90 1000 1 45 1 0 95 300",
        "M1 1S2 2 3 3T4 4",
        "", "  \t\n", "Z", "M5 5zz",
    ]
    for d in strings:
        want = [tuple(float(v) if not isinstance(v, str) else v for v in e) for e in parse_svg_path(d)]
        got = NativePath().svg(d).elements()
        assert got == want, d[:60]
    for bad in ["10 10", "M10", "M1 1 L", "M0 0A1 1 0 2 0 5 5", "M0 0 X 1 1", "M 1 1 Z 5", "M1e"]:
        with pytest.raises(ValueError):
            parse_svg_path(bad)
        with pytest.raises(ValueError):
            NativePath().svg(bad)


def test_invalid_arguments_are_rejected_not_crashed():
    """NULL handles / pointers and malformed inputs come back as VB_E_INVALID."""
    import ctypes as C
    from vello_b200.scene_native import _lib, _Brush, _Path, _Stroke, _Packed
    lib = _lib()
    ident = (C.c_double * 6)(1, 0, 0, 1, 0, 0)
    s = C.c_void_p(lib.vb_scene_new())
    verbs = (C.c_uint8 * 2)(ord("M"), ord("X"))  # unknown verb
    coords = (C.c_double * 4)(0, 0, 1, 1)
    bad_path = _Path(C.cast(verbs, C.c_void_p), 2, C.cast(coords, C.c_void_p))
    brush = _Brush()
    assert lib.vb_scene_fill(None, 0, ident, C.byref(brush), None, C.byref(bad_path)) == -1
    assert lib.vb_scene_fill(s, 0, None, C.byref(brush), None, C.byref(bad_path)) == -1
    assert lib.vb_scene_fill(s, 0, ident, None, None, C.byref(bad_path)) == -1
    assert lib.vb_scene_fill(s, 0, ident, C.byref(brush), None, None) == -1
    assert lib.vb_scene_fill(s, 0, ident, C.byref(brush), None, C.byref(bad_path)) == -1  # the verb
    brush.kind = 99
    ok_verbs = (C.c_uint8 * 3)(ord("M"), ord("L"), ord("L"))
    ok_coords = (C.c_double * 6)(0, 0, 5, 0, 5, 5)
    ok_path = _Path(C.cast(ok_verbs, C.c_void_p), 3, C.cast(ok_coords, C.c_void_p))
    assert lib.vb_scene_fill(s, 0, ident, C.byref(brush), None, C.byref(ok_path)) == -1  # unknown brush kind
    brush.kind = 4  # image brush without an image
    assert lib.vb_scene_fill(s, 0, ident, C.byref(brush), None, C.byref(ok_path)) == -1
    brush.kind = 1  # gradient claiming stops it does not have
    brush.n_stops = 3
    assert lib.vb_scene_fill(s, 0, ident, C.byref(brush), None, C.byref(ok_path)) == -1
    st = _Stroke(2.0, 0, 0, 0, 4.0)
    assert lib.vb_scene_stroke(s, None, ident, C.byref(brush), None, C.byref(ok_path)) == -1
    assert lib.vb_scene_stroke(None, C.byref(st), ident, C.byref(brush), None, C.byref(ok_path)) == -1
    assert lib.vb_scene_push_clip_layer(s, 0, None, None, C.byref(ok_path)) == -1
    assert lib.vb_scene_pop_layer(None) == -1
    assert lib.vb_scene_append(s, s, None) == -1  # a scene cannot be appended to itself
    assert lib.vb_scene_resolve(s, None) == -1
    pk = _Packed()
    assert lib.vb_scene_resolve(s, C.byref(pk)) == 0  # still a valid (if odd) scene
    assert lib.vb_pathbuf_circle(None, 0.0, 0.0, 1.0, 0.1) == -1
    pb = C.c_void_p(lib.vb_pathbuf_new())
    assert lib.vb_pathbuf_circle(pb, 0.0, 0.0, 1.0, 0.0) == -1  # tolerance must be positive
    assert lib.vb_pathbuf_rounded_rect(pb, 0.0, 0.0, 1.0, 1.0, 0.2, float("nan")) == -1
    assert lib.vb_pathbuf_svg(pb, None) == -1
    lib.vb_pathbuf_free(pb)
    lib.vb_scene_free(s)
    lib.vb_scene_free(None)
    lib.vb_pathbuf_free(None)


def test_dash_native_equals_python():
    """kurbo::dash in both front ends: the native (C++) dash expansion produces the Python statement's stream byte for byte --
    lines (closed forms), closed subpaths (the stashed first dash joins the last), offsets, odd patterns, quads and cubics."""
    import ctypes as C
    from vello_b200 import shapes
    from vello_b200.encoding import Stroke, Scene, resolve, STYLE_JOIN_BEVEL, STYLE_CAP_BUTT, Color
    from vello_b200.scene_native import NativeScene
    from vello_b200.shapes import Affine, BezPath, Circle, Rect
    rng = np.random.default_rng(3)
    cases = [(Rect(10, 10, 90, 70), (7.0, 3.0), 0.0), (Rect(10, 10, 90, 70), (7.0, 3.0, 1.0), 4.5), (Circle(50, 50, 30), (5.0, 2.5), 1.0),
             (BezPath([("M", 0.0, 0.0), ("L", 100.0, 0.0), ("Q", 120.0, 40.0, 60.0, 80.0), ("C", 20.0, 120.0, 10.0, 20.0, 90.0, 90.0)]), (9.0, 4.0), 30.0),
             (BezPath([("M", 5.0, 5.0), ("L", 50.0, 5.0), ("L", 50.0, 50.0), ("Z",), ("M", 60.0, 60.0), ("L", 90.0, 95.0)]), (4.0, 4.0), 0.0)]
    p = BezPath()
    p.move_to(*rng.uniform(0, 100, 2))
    for _ in range(30):
        p.line_to(*rng.uniform(0, 100, 2))
    cases.append((p, (3.0, 1.0, 0.5, 1.0), 2.0))
    py, nat = Scene(), NativeScene()
    for shape, pat, off in cases:
        st = Stroke(2.0, join=STYLE_JOIN_BEVEL, start_cap=STYLE_CAP_BUTT, end_cap=STYLE_CAP_BUTT, dash_pattern=pat, dash_offset=off)
        for s in (py, nat):
            s.stroke(st, Affine.translate(3.0, 4.0), Color.from_rgba8(255, 255, 0), None, shape)
    a, b = resolve(py.encoding), nat.resolve()
    assert a.scene.tobytes() == b.scene.tobytes()
    assert a.layout.as_array().tolist() == b.layout.as_array().tolist()
    # known answers of the state machine on a straight line
    d = shapes.dash([("M", 0.0, 0.0), ("L", 10.0, 0.0)], 0.0, [1.0, 1.0])
    assert len(d) == 10 and d[0] == ("M", 2.0, 0.0) and d[-2:] == [("M", 0.0, 0.0), ("L", 1.0, 0.0)]
    d = shapes.dash([("M", 0.0, 0.0), ("L", 4.0, 0.0), ("L", 4.0, 4.0), ("L", 0.0, 4.0), ("Z",)], 0.5, [3.0, 1.0])
    assert d[-3:] == [("M", 0.0, 0.5), ("L", 0.0, 0.0), ("L", 2.5, 0.0)]  # last dash joined to the stashed first one


@pytest.mark.parametrize("seed", range(12))
def test_dash_fuzz_native_equals_python(seed):
    """Random subpaths (lines, quads, cubics, closes, degenerate segments), random patterns and offsets (also negative and
    longer than the pattern): the C++ dash expansion and the Python one produce the same byte stream."""
    from vello_b200.encoding import Stroke, Scene, resolve, STYLE_JOIN_MITER, STYLE_CAP_SQUARE, Color
    from vello_b200.scene_native import NativeScene
    from vello_b200.shapes import Affine, BezPath
    rng = np.random.default_rng(1000 + seed)
    py, nat = Scene(), NativeScene()
    for _ in range(6):
        p = BezPath()
        for _sub in range(int(rng.integers(1, 4))):
            p.move_to(*rng.uniform(0, 200, 2))
            for _seg in range(int(rng.integers(1, 8))):
                k = int(rng.integers(0, 8))
                if k < 4:
                    p.line_to(*rng.uniform(0, 200, 2))
                elif k < 5:
                    last = p.els[-1]
                    p.line_to(last[-2], last[-1])  # zero-length segment
                elif k < 6:
                    p.quad_to(*rng.uniform(0, 200, 4))
                else:
                    p.curve_to(*rng.uniform(0, 200, 6))
            if rng.random() < 0.4:
                p.close_path()
        n = int(rng.integers(1, 6))
        pat = tuple(float(v) for v in np.round(rng.uniform(0.5, 25.0, n), 3))
        off = float(np.round(rng.uniform(-40.0, 80.0), 3))
        st = Stroke(float(np.round(rng.uniform(0.5, 6.0), 2)), join=STYLE_JOIN_MITER, start_cap=STYLE_CAP_SQUARE, end_cap=STYLE_CAP_SQUARE,
                    dash_pattern=pat, dash_offset=off)
        for s in (py, nat):
            s.stroke(st, Affine.IDENTITY, Color.from_rgba8(10, 200, 90), None, p)
    a, b = resolve(py.encoding), nat.resolve()
    assert a.layout.as_array().tolist() == b.layout.as_array().tolist()
    assert a.scene.tobytes() == b.scene.tobytes()
