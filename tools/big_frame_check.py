"""C4-sized sanity run (16384^2, 1 Mi tiles, 1 GiB of pixels): the frame renders without arena
failures, bin-row stripes reproduce the rows of the full frame bit for bit, and occlusion culling changes nothing.
usage: python tools/big_frame_check.py [size] [n_paths]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from vello_b200 import scenes
from vello_b200.config import AA_MSAA16, RenderParams
from vello_b200.encoding import BLACK, resolve
from vello_b200.renderer import Renderer, RendererOptions

size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n_paths = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
packed = resolve(scenes.paris_like(n_paths, size, seed=30000).encoding)
p = RenderParams(BLACK, size, size, AA_MSAA16)
r = Renderer(RendererOptions(timing=True))
t0 = time.time()
full = r.render_to_texture(packed, p)
st = r.last_stats.as_dict()
print("full frame", full.shape, "first call %.2f s" % (time.time() - t0), "retries", st["retries"], "arena MB", st["arena_bytes"] >> 20)
r.upload(packed)
for _ in range(3):
    sd = r.render_resident(p).as_dict()
print("resident frame ms", round(sd["total_ms"], 3), {k: round(v, 3) for k, v in sd["stage_ms"].items()})
rows = (size + 255) // 256
for b in (0, rows // 3, rows - 1):
    stripe = r.render_to_texture(packed, p, bin_rows=(b, b + 1))
    assert np.array_equal(stripe, full[b * 256:(b + 1) * 256]), f"stripe {b} differs"
r.set_occlusion_cull(False)
assert np.array_equal(r.render_to_texture(packed, p, bin_rows=(1, 3)), full[256:768]), "cull on/off differ"
print("stripes == full frame rows, cull on == off; checksum", int(full.astype(np.uint64).sum()))
