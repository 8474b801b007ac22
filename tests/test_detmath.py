"""vb_detmath.h (the IEEE-only elementary functions shared by the CUDA kernels and the oracle) against float64
libm: accuracy in ulps on the argument ranges the pipeline uses. WGSL allows far more (e.g. atan2 4096 ulp,
sin/cos absolute error 2^-11), so a few ulp keeps every use inside the reference's own tolerance."""
import math

import numpy as np
import pytest

from oracle.vbo import Oracle

FN = dict(sin=0, cos=1, atan2=2, asin=3, acos=4, pow23=5, exp=6, pow=7, cbrt=8, log=9)


def ulp_err(got, want):
    want32 = np.float32(want)
    if want32 == 0 or not np.isfinite(want32):
        return abs(float(got) - float(want)) / float(np.finfo(np.float32).tiny) if want32 == 0 and got != 0 else 0.0
    return abs(float(got) - float(want)) / float(np.spacing(abs(want32)))


@pytest.fixture(scope="module")
def o():
    return Oracle()


def test_not_libm(o):
    assert o.lib.vbo_uses_libm() == 0


@pytest.mark.parametrize("name,ref,lo,hi,tol", [
    ("sin", math.sin, -20.0, 20.0, 2.5), ("cos", math.cos, -20.0, 20.0, 2.5),
    ("asin", math.asin, -1.0, 1.0, 2.5), ("acos", math.acos, -1.0, 1.0, 2.5),
    ("cbrt", lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x), -1e6, 1e6, 2.5),
    ("pow23", lambda x: abs(x) ** (2.0 / 3.0), 0.0, 100.0, 4.0),
    ("exp", math.exp, -80.0, 80.0, 2.5), ("log", math.log, 1e-30, 1e30, 2.5),
])
def test_unary(o, name, ref, lo, hi, tol):
    rng = np.random.default_rng(1)
    xs = rng.uniform(lo, hi, 4000).astype(np.float32)
    if name == "log":
        xs = np.exp(rng.uniform(math.log(lo), math.log(hi), 4000)).astype(np.float32)
    worst = 0.0
    for x in xs:
        x = float(x)
        if name in ("sin", "cos") and abs(ref(x)) < 1e-3:
            assert abs(o.math(FN[name], x) - ref(x)) < 2e-7  # near zeros: absolute error
            continue
        worst = max(worst, ulp_err(o.math(FN[name], x), ref(x)))
    assert worst <= tol, f"{name}: {worst} ulp"


def test_atan2(o):
    rng = np.random.default_rng(2)
    worst = 0.0
    for y, x in rng.normal(0, 3, (4000, 2)).astype(np.float32):
        worst = max(worst, ulp_err(o.math(FN["atan2"], float(y), float(x)), math.atan2(float(y), float(x))))
    assert worst <= 3.0
    assert o.math(FN["atan2"], 0.0, -1.0) == pytest.approx(math.pi, rel=1e-7)
    assert o.math(FN["atan2"], 1.0, 0.0) == pytest.approx(math.pi / 2, rel=1e-7)


def test_pow_for_blur(o):
    rng = np.random.default_rng(3)
    for x, y in zip(rng.uniform(0.01, 300, 2000), rng.uniform(0.2, 4.0, 2000)):
        got, want = o.math(FN["pow"], float(np.float32(x)), float(np.float32(y))), float(np.float32(x)) ** float(np.float32(y))
        assert abs(got - want) <= 4e-6 * abs(want)
    assert o.math(FN["pow"], 0.0, 2.5) == 0.0
