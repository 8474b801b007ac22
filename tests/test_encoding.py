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


def test_f16_reference_vectors():
    """The exact inputs / outputs of the reference's own unit tests, vello_encoding/src/math.rs:151-281."""
    import math
    import struct

    def bits(u):
        return struct.unpack("<f", struct.pack("<I", u))[0]

    assert f32_to_f16(math.pi) == 0x4248                      # test_f32_to_f16_simple
    assert f32_to_f16(bits(0x7F800001)) == 0x7E00             # signalling NaN -> quiet NaN, not infinity
    assert f32_to_f16(bits(0x7F800000)) == 0x7C00             # +inf
    assert f32_to_f16(0.00003051758) == 0x0200                # exponent rebias (a half denormal)
    assert f32_to_f16(1.701412e38) == 0x7C00 and f32_to_f16(-1.701412e38) == 0xFC00
    assert f16_to_f32(0x4248) == np.float32(3.140625)
    assert np.isinf(f16_to_f32(0x7C00)) and np.isinf(f16_to_f32(0xFC00)) and f16_to_f32(0xFC00) < 0
    assert np.isnan(f16_to_f32(0x7C01)) and np.isnan(f16_to_f32(f32_to_f16(f16_to_f32(0x7C01))))
    for h in (0x7C00, 0xFC00, 0x7BFF, 0xFBFF, 0x0001, 0x8001):  # the *_roundtrip tests
        assert f32_to_f16(f16_to_f32(h)) == h, hex(h)
    assert abs(math.pi - float(f16_to_f32(f32_to_f16(math.pi)))) < 0.001


def test_extend_and_quality_values():
    # vello_encoding/src/encoding.rs:623-645: the numeric values the shaders decode
    from vello_b200.encoding import EXTEND_PAD, EXTEND_REFLECT, EXTEND_REPEAT, QUALITY_HIGH, QUALITY_LOW, QUALITY_MEDIUM
    assert (EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT) == (0, 1, 2)
    assert (QUALITY_LOW, QUALITY_MEDIUM, QUALITY_HIGH) == (0, 1, 2)


def test_draw_color_is_premultiplied_little_endian():
    # draw.rs:281-297: premultiplied, r in the low byte
    assert Color.from_rgba8(255, 0, 0, 255).premul_rgba8_u32() == 0xFF0000FF
    assert Color.from_rgba8(0, 255, 0, 255).premul_rgba8_u32() == 0xFF00FF00
    assert Color(1.0, 1.0, 1.0, 0.5).premul_rgba8_u32() == 0x80808080
    # the reference's own vectors (draw_color_endianness, draw_color_premultiplied)
    assert Color.from_rgba8(0x00, 0xCA, 0xFE, 0xFF).premul_rgba8_u32().to_bytes(4, "little") == bytes([0x00, 0xCA, 0xFE, 0xFF])
    assert Color.from_rgba8(0x00, 0xCA, 0xFE, 0x00).premul_rgba8_u32() == 0


def test_stroke_style_flags_roundtrip():
    # path.rs:848-877
    st = style_from_stroke(Stroke(3.5, miter_limit=4.0))
    assert st[0] & STYLE_FLAGS_STYLE_BIT and st[1] == 3.5
    assert f16_to_f32(st[0] & 0xFFFF) == 4.0
    assert style_from_stroke(Stroke(0.0)) is None


def test_fill_and_stroke_style_fields():
    """path.rs:846-877 (test_fill_style, test_stroke_style): every cap x cap x join combination decodes through the masks the
    shaders use (flatten.wgsl:20-32)."""
    from vello_b200.encoding import (FILL_EVEN_ODD, STYLE_CAP_BUTT, STYLE_CAP_ROUND, STYLE_CAP_SQUARE, STYLE_FLAGS_FILL_BIT, STYLE_JOIN_BEVEL,
                                     STYLE_JOIN_MITER, STYLE_JOIN_ROUND, style_from_fill)
    assert style_from_fill(FILL_NON_ZERO)[0] == 0 and style_from_fill(FILL_EVEN_ODD)[0] == STYLE_FLAGS_FILL_BIT
    assert not style_from_fill(FILL_NON_ZERO)[0] & STYLE_FLAGS_STYLE_BIT  # a fill has no stroke width
    assert style_from_stroke(Stroke(1.0))[0] & STYLE_FLAGS_STYLE_BIT      # and a stroke is not a fill
    caps, joins = (STYLE_CAP_BUTT, STYLE_CAP_SQUARE, STYLE_CAP_ROUND), (STYLE_JOIN_BEVEL, STYLE_JOIN_MITER, STYLE_JOIN_ROUND)
    for start in caps:
        for end in caps:
            for join in joins:
                flags, width = style_from_stroke(Stroke(1.0, join=join, miter_limit=0.0, start_cap=start, end_cap=end))
                assert width == 1.0
                assert flags & 0x3000_0000 == join
                assert (flags & 0x0C00_0000) >> 2 == start
                assert flags & 0x0300_0000 == end
                assert flags & 0xFFFF == 0


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
