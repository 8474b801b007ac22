"""The 16-bit coordinate path (vello_encoding/src/path.rs:244-316: tags without the F32 bit carry points as two i16 in
one word; read at flatten.wgsl:639-644). The reference's encoder never emits it, but the shader supports it and so do the
oracle and k_flatten: a scene with integral coordinates is re-packed in the i16 form (half the path data) and must give
the same lines, bounding boxes and pixels as its f32 form."""
import copy

import numpy as np
import pytest

from vello_b200.config import AA_MSAA16, RenderParams
from vello_b200.encoding import BLACK, FILL_EVEN_ODD, FILL_NON_ZERO, Color, Packed, Scene, resolve
from vello_b200.shapes import Affine, BezPath


def integral_scene(seed=3, n=120) -> Scene:
    rng = np.random.default_rng(seed)
    s = Scene()
    for k in range(n):
        p = BezPath()
        x, y = (int(v) for v in rng.integers(-40, 300, 2))
        p.move_to(x, y)
        for _ in range(int(rng.integers(2, 7))):
            kind = int(rng.integers(0, 3))
            pts = [int(v) for v in rng.integers(-60, 320, 6)]
            if kind == 0:
                p.line_to(pts[0], pts[1])
            elif kind == 1:
                p.quad_to(*pts[:4])
            else:
                p.curve_to(*pts)
        if k % 3:
            p.close_path()
        if k % 7 == 0:  # a second subpath in the same path
            p.move_to(*[int(v) for v in rng.integers(0, 256, 2)])
            p.line_to(*[int(v) for v in rng.integers(0, 256, 2)])
            p.line_to(*[int(v) for v in rng.integers(0, 256, 2)])
        t = Affine.translate(float(rng.integers(-5, 5)), 0.5) if k % 4 else Affine.scale(1.25)
        s.fill(FILL_NON_ZERO if k % 2 else FILL_EVEN_ODD, t, Color.from_rgba8(*[int(v) for v in rng.integers(0, 256, 4)]), None, p)
    return s


def to_i16(packed: Packed) -> Packed:
    """Re-pack every path segment with 16-bit coordinates: clear the F32 bit (0x8) of the segment tags, store each point as
    x | y << 16, and shift the stream offsets that follow the path data."""
    L = packed.layout
    words = packed.scene
    n_tag_words = L.path_data_base - L.path_tag_base
    tags = words[L.path_tag_base:L.path_data_base].copy().view(np.uint8)
    seg = (tags & 3) != 0
    assert np.all((tags[seg] & 8) != 0), "expected an all-f32 scene"
    tags[seg] &= np.uint8(0xF7)
    pd = words[L.path_data_base:L.draw_tag_base].view(np.float32)
    assert pd.size % 2 == 0 and np.all(pd == np.round(pd)) and np.all(np.abs(pd) < 32768)
    xy = pd.astype(np.int32).reshape(-1, 2)
    pts = ((xy[:, 0] & 0xFFFF) | ((xy[:, 1] & 0xFFFF) << 16)).astype(np.uint32)
    rest = words[L.draw_tag_base:]
    scene = np.concatenate([tags.view(np.uint32), pts, rest])
    lay = copy.copy(L)
    delta = pd.size - pts.size
    lay.draw_tag_base -= delta
    lay.draw_data_base -= delta
    lay.transform_base -= delta
    lay.style_base -= delta
    assert n_tag_words + pts.size == lay.draw_tag_base
    return Packed(scene=np.ascontiguousarray(scene), layout=lay, ramps=packed.ramps, atlas=packed.atlas)


def test_oracle_i16_equals_f32(oracle):
    f32p = resolve(integral_scene().encoding)
    i16p = to_i16(f32p)
    assert i16p.scene.nbytes < f32p.scene.nbytes
    out = []
    for pk in (f32p, i16p):
        oracle.bind(pk, 256, 256)
        oracle.run("pathtag", "flatten")
        out.append((oracle.buffer("lines").copy(), oracle.buffer("path_bboxes").copy()))
    assert out[0][0].tobytes() == out[1][0].tobytes() and len(out[0][0]) > 1000
    assert out[0][1].tobytes() == out[1][1].tobytes()
    a = oracle.render(f32p, 256, 256, BLACK.premul_rgba8_u32(), AA_MSAA16)
    b = oracle.render(i16p, 256, 256, BLACK.premul_rgba8_u32(), AA_MSAA16)
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_gpu_i16_equals_f32_and_oracle(oracle):
    from oracle.vbo import DTYPES
    from vello_b200.renderer import Renderer
    r = Renderer()
    f32p = resolve(integral_scene(seed=8).encoding)
    i16p = to_i16(f32p)
    p = RenderParams(BLACK, 256, 256, AA_MSAA16)
    img_f = r.render_to_texture(f32p, p)
    lines_f = r.download("lines", DTYPES["lines"]).copy()
    img_i = r.render_to_texture(i16p, p)
    lines_i = r.download("lines", DTYPES["lines"])
    assert lines_f.tobytes() == lines_i.tobytes()
    assert np.array_equal(img_f, img_i)
    assert np.array_equal(img_i, oracle.render(i16p, 256, 256, BLACK.premul_rgba8_u32(), AA_MSAA16))
    r.close()
