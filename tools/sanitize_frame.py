"""Frames run under compute-sanitizer (tools/sanitize.sh): graph replay off (the sanitizer tools instrument plain launches),
every kernel of the pipeline, all three AA modes, clips / blends / gradients / images / strokes, a stripe window and an arena
overflow with grow-and-retry. Pixels are checked against the oracle so a 'clean' log is also a correct frame."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VELLO_B200_NO_GRAPH"] = "1"
from vello_b200 import scenes  # noqa: E402
from vello_b200.config import RenderParams  # noqa: E402
from vello_b200.encoding import BLACK, resolve  # noqa: E402
from vello_b200.renderer import Renderer  # noqa: E402
from oracle.vbo import Oracle  # noqa: E402

r, o = Renderer(), Oracle(threads=4)
cases = [(scenes.tiger(256, 256), 256, 256), scenes.brushes(), scenes.deep_blend(), scenes.stroke_styles(), scenes.many_clips(),
         (scenes.paris_like(300, 512, seed=1), 512, 512)]
for s, w, h in cases:
    packed = resolve(s.encoding)
    for aa in (2, 1, 0):
        img = r.render_to_texture(packed, RenderParams(BLACK, w, h, aa))
        ref = o.render(packed, w, h, BLACK.premul_rgba8_u32(), aa)
        d = int(np.abs(img.astype(int) - ref.astype(int)).max())
        assert d <= (1 if aa == 0 else 0), (w, h, aa, d)
packed = resolve(scenes.paris_like(300, 512, seed=1).encoding)
ref = o.render(packed, 512, 512, BLACK.premul_rgba8_u32(), 2)
assert np.array_equal(r.render_to_texture(packed, RenderParams(BLACK, 512, 512, 2), bin_rows=(1, 2)), ref[256:512])
# one frame on three renderers with the sharded flatten + peer-memory exchange (k_exchange.cu), device frame and host destination
from vello_b200.renderer import RendererGroup  # noqa: E402
g = RendererGroup([0, 0, 0])
g.set_exchange(True)
assert np.array_equal(g.render_to_texture(packed, RenderParams(BLACK, 512, 512, 2)), ref)
g.upload(packed)
for _ in range(2):
    g.render_resident(RenderParams(BLACK, 512, 512, 2))
    assert np.array_equal(g.frame_to_host(RenderParams(BLACK, 512, 512, 2)), ref)
g.close()
# Resolver::resolve on the device (k_resolve.cu)
from vello_b200.scene_native import NativeScene  # noqa: E402
from vello_b200.encoding import FILL_NON_ZERO, Gradient  # noqa: E402
from vello_b200.shapes import Affine, Rect  # noqa: E402
nat = NativeScene()
nat.fill(FILL_NON_ZERO, Affine.IDENTITY, Gradient.linear((0, 0), (200, 100), [(0.0, BLACK), (1.0, BLACK.with_alpha(0.3))]), None, Rect(5, 5, 190, 120))
nat.upload_device(r)
r.render_resident(RenderParams(BLACK, 200, 128, 2))
print("sanitize_frame: all frames match the oracle; launches of the last frame:", r.last_stats.as_dict()["kernel_launches"])
r.close()
