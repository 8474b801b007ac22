"""Host-side encoding checks: the format known answers the reference's unit tests hold for this path
(vello_encoding/src/math.rs:147-281 f16, draw.rs:281-297 colour packing, path.rs:848-877 style flags), the tag
monoid bit magic (doc/pathseg.md), and bulk == generic encoders."""
import numpy as np

from vello_b200 import scenes
from vello_b200.config import make_mask_lut, make_mask_lut_16
from vello_b200.encoding import (BLUE, Color, Encoding, FILL_NON_ZERO, PathEncoder, Scene, Stroke, f16_to_f32, f32_to_f16, resolve,
                                 STYLE_FLAGS_STYLE_BIT, style_from_stroke, TAG_LINE_TO_F32, TAG_PATH, TAG_SUBPATH_END_BIT, TAG_QUAD_TO_F32)
from vello_b200.shapes import Affine, BezPath, Rect


def test_f16_known_answers():
    # math.rs:147-281: exact halves, rounding, inf / nan, denormals
    for f, h in [(0.0, 0x0000), (1.0, 0x3C00), (-2.0, 0xC000), (65504.0, 0x7BFF), (6.1035156e-5, 0x0400), (5.9604645e-8, 0x0001),
                 (float("inf"), 0x7C00), (4.0, 0x4400), (0.5, 0x3800)]:
        assert f32_to_f16(f) == h, (f, hex(f32_to_f16(f)))
        assert f16_to_f32(h) == np.float32(f)
    assert f32_to_f16(float("nan")) & 0x7C00 == 0x7C00 and f32_to_f16(float("nan")) & 0x3FF != 0
    assert f32_to_f16(1e6) == 0x7C00  # overflow clamps to infinity


def test_draw_color_is_premultiplied_little_endian():
    # draw.rs:281-297: premultiplied, r in the low byte
    assert Color.from_rgba8(255, 0, 0, 255).premul_rgba8_u32() == 0xFF0000FF
    assert Color.from_rgba8(0, 255, 0, 255).premul_rgba8_u32() == 0xFF00FF00
    assert Color(1.0, 1.0, 1.0, 0.5).premul_rgba8_u32() == 0x80808080


def test_stroke_style_flags_roundtrip():
    # path.rs:848-877
    st = style_from_stroke(Stroke(3.5, miter_limit=4.0))
    assert st[0] & STYLE_FLAGS_STYLE_BIT and st[1] == 3.5
    assert f16_to_f32(st[0] & 0xFFFF) == 4.0
    assert style_from_stroke(Stroke(0.0)) is None


def test_tag_stream_of_a_fill_and_an_open_stroke():
    e = Encoding()
    pe = PathEncoder(e, True)
    pe.path_elements([("M", 0, 0), ("L", 10, 0), ("L", 10, 10)])
    assert pe.finish(True) == 3  # closing line added
    assert e.path_tags == [TAG_LINE_TO_F32, TAG_LINE_TO_F32, TAG_LINE_TO_F32 | TAG_SUBPATH_END_BIT, TAG_PATH]
    e = Encoding()
    pe = PathEncoder(e, False)
    pe.path_elements([("M", 0, 0), ("L", 10, 0), ("L", 10, 10)])
    pe.finish(True)
    assert e.path_tags == [TAG_LINE_TO_F32, TAG_LINE_TO_F32, TAG_QUAD_TO_F32 | TAG_SUBPATH_END_BIT, TAG_PATH]  # quad-to cap marker


def test_bulk_appenders_match_generic_encoder():
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 100, (17, 2)).astype(np.float32)
    a, b = Scene(), Scene()
    p = BezPath([("M", *pts[0])] + [("L", *q) for q in pts[1:]] + [("Z",)])
    a.fill(FILL_NON_ZERO, Affine.IDENTITY, BLUE, None, p)
    b.encoding.encode_transform(Affine.IDENTITY)
    b.encoding.encode_fill_style(FILL_NON_ZERO)
    scenes._bulk_fill_polygon(b.encoding, pts)
    b.encoding.encode_color(BLUE)
    assert np.array_equal(resolve(a.encoding).scene, resolve(b.encoding).scene)
    a, b = Scene(), Scene()
    st = Stroke(2.5)
    a.stroke(st, Affine.IDENTITY, BLUE, None, BezPath([("M", *pts[0])] + [("L", *q) for q in pts[1:]]))
    b.encoding.encode_transform(Affine.IDENTITY)
    b.encoding.encode_stroke_style(st)
    scenes._bulk_stroke_polyline(b.encoding, pts)
    b.encoding.encode_color(BLUE)
    assert np.array_equal(resolve(a.encoding).scene, resolve(b.encoding).scene)


def test_layout_and_padding():
    s = Scene()
    s.fill(FILL_NON_ZERO, Affine.IDENTITY, BLUE, None, Rect(0, 0, 4, 4))
    s.push_clip_layer(FILL_NON_ZERO, Affine.IDENTITY, Rect(0, 0, 2, 2))  # left open: resolve closes it (resolve.rs:127-141)
    p = resolve(s.encoding)
    L = p.layout
    assert L.path_tag_base == 0 and L.path_data_base == 256  # tags padded to 1024 B
    # as the reference: the closing END_CLIP / PATH tags are appended to the streams, but Layout keeps the
    # encoding's own counts (resolve.rs:113-117,151)
    assert L.n_paths == L.n_draw_objects == 2 and L.n_clips == 1
    assert p.scene[L.draw_tag_base: L.draw_tag_base + 3].tolist() == [0x44, 0x49, 0x21]
    assert L.bin_data_start == 2


def test_mask_luts_shape_and_symmetry():
    l8, l16 = make_mask_lut(), make_mask_lut_16()
    assert l8.shape == (256,) and l16.shape == (2048,)
    b8 = l8.view(np.uint8)
    # a full-coverage entry and an empty entry must exist (mask.rs one_mask over slope/translation grid)
    assert 0xFF in b8 and 0x00 in b8
