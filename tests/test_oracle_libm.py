"""The oracle built against libm (the literal arithmetic of vello_shaders/src/cpu: Rust std -> platform libm) versus
the default build (vb_detmath.h). Different last-ulp results in sin/cos/atan2 can change a line COUNT, so buffers are
not bit-identical, but rendered pixels must agree within 1 LSB almost everywhere -- this bounds how far the
reproducible-math convention moves us from the reference's own CPU path."""
import numpy as np
import pytest

from vello_b200 import scenes
from vello_b200.encoding import BLACK, resolve


@pytest.mark.parametrize("name", ["stroke_styles", "fill_types", "many_clips"])
def test_libm_vs_detmath_pixels(oracle, oracle_libm, name):
    s, w, h = getattr(scenes, name)()
    p = resolve(s.encoding)
    for aa in (0, 2):
        a = oracle.render(p, w, h, BLACK.premul_rgba8_u32(), aa)
        b = oracle_libm.render(p, w, h, BLACK.premul_rgba8_u32(), aa)
        d = np.abs(a.astype(int) - b.astype(int))
        assert (d > 1).mean() < 2e-4, f"{name} aa={aa}: {(d > 1).sum()} channel values differ by more than 1 LSB"
        assert d.max() <= 40


def test_libm_tiger(oracle, oracle_libm):
    p = resolve(scenes.tiger(512, 512).encoding)
    a = oracle.render(p, 512, 512, BLACK.premul_rgba8_u32(), 0)
    b = oracle_libm.render(p, 512, 512, BLACK.premul_rgba8_u32(), 0)
    d = np.abs(a.astype(int) - b.astype(int))
    assert (d > 1).mean() < 2e-4
    la, lb = oracle.buffer("bump")["lines"][0], oracle_libm.buffer("bump")["lines"][0]
    assert abs(int(la) - int(lb)) <= max(4, int(la) // 2000)
