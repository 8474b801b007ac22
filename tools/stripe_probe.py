"""Per-stage times of ONE stripe of the headline frame on one GPU (what a rank of an N-GPU run executes):
    VELLO_B200_LIB=... python tools/stripe_probe.py [n_stripes=8] [stripe=3] [scene=paris|beziers]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vello_b200 import scenes  # noqa: E402
from vello_b200.config import RenderParams  # noqa: E402
from vello_b200.encoding import BLACK, resolve  # noqa: E402
from vello_b200.renderer import Renderer, RendererOptions  # noqa: E402
from vello_b200.stripes import even_tile_bounds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = sys.argv[3] if len(sys.argv) > 3 else "paris"
sc = scenes.paris_like(30000, 4096, 30000) if which == "paris" else scenes.beziers_clips(100000, 1000, 4096, seed=100000)
packed = resolve(sc.encoding)
p = RenderParams(BLACK, 4096, 4096, 2)
b = even_tile_bounds(n, 4096)
rows = (b[k], b[k + 1])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
rt = Renderer(RendererOptions(device=0, timing=True))
rt.upload(packed)
rt.render_resident(p, 0, tile_rows=rows)
acc = {}
for _ in range(10):
    flush.fill_(1)
    torch.cuda.synchronize()
    sd = rt.render_resident(p, 0, tile_rows=rows).as_dict()
    for s, v in sd["stage_ms"].items():
        acc[s] = acc.get(s, 0.0) + v / 10
rg = Renderer(RendererOptions(device=0))
rg.upload(packed)
rg.render_resident(p, 0, tile_rows=rows)
ms = []
for _ in range(20):
    flush.fill_(1)
    torch.cuda.synchronize()
    rg.render_resident(p, 0, tile_rows=rows)
    ms.append(rg.lib.vb_last_frame_ms(rg.handle))
print(os.environ.get("VELLO_B200_LIB", "default"), f"stripe {k}/{n} rows {rows}: graph frame {np.median(ms):.3f} ms; stages",
      {s: round(v, 3) for s, v in acc.items()}, "sum", round(sum(acc.values()), 3))
