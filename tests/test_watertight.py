"""Watertightness of flatten's output, the reference's own debug validation (vello/src/debug/validate.rs:47-64): every
end point of every line of a path, compared bit for bit, must be matched by another end point of the same path. Checked
here on the oracle's line soup (CPU); tests/parity.compare_all applies the same check to the CUDA path's lines."""
import numpy as np
import pytest

from vello_b200 import scenes
from vello_b200.encoding import resolve

from .parity import unpaired_endpoints


@pytest.mark.parametrize("name", ["tiger", "stroke_styles", "fill_types", "tricky_strokes", "robust_paths", "funky_paths", "many_clips",
                                  "blend_grid", "two_point_radial", "paris"])
def test_oracle_lines_are_watertight(oracle, name):
    if name == "tiger":
        s, w, h = scenes.tiger(512, 512), 512, 512
    elif name == "paris":
        s, w, h = scenes.paris_like(1500, 1024, seed=3), 1024, 1024
    else:
        s, w, h = getattr(scenes, name)()
    oracle.bind(resolve(s.encoding), w, h)
    oracle.run("pathtag", "flatten")
    lines = oracle.buffer("lines")
    assert len(lines) > 0
    assert len(unpaired_endpoints(lines)) == 0


def test_checker_sees_a_gap(oracle):
    s, w, h = scenes.fill_types()
    oracle.bind(resolve(s.encoding), w, h)
    oracle.run("pathtag", "flatten")
    lines = oracle.buffer("lines").copy()
    lines["p1"][5, 0] = np.nextafter(lines["p1"][5, 0], np.float32(1e9))  # one ulp
    assert len(unpaired_endpoints(lines)) == 2
