"""Where does a host-to-host frame spend its time? Times the pieces of vb_render separately (paris-like-30k 4096^2 MSAA16)."""
import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from vello_b200 import scenes
from vello_b200.config import AA_MSAA16, RenderParams
from vello_b200.encoding import BLACK, resolve
from vello_b200.renderer import Renderer, RendererOptions, FrameStats, _Layout, _params_struct

size = 4096
packed = resolve(scenes.paris_like(30000, size, seed=30000).encoding)
p = RenderParams(BLACK, size, size, AA_MSAA16)
r = Renderer(RendererOptions(timing=True))
scene_h = torch.from_numpy(np.ascontiguousarray(packed.scene)).pin_memory()
outs = [torch.empty((size, size, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
lay = _Layout(*[int(v) for v in packed.layout.as_array()])
ps = _params_struct(p, (0, 0))
fs = FrameStats()
atlas = np.ascontiguousarray(packed.atlas)

def T(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def upload():
    r.lib.vb_scene_upload(r.handle, scene_h.data_ptr(), scene_h.numel() * 4, C.byref(lay), None, 512, 0, atlas.ctypes.data, 0, 0)
def resident():
    r.lib.vb_render_resident(r.handle, C.byref(ps), None, C.byref(fs))
def readback():
    r.lib.vb_copy_to_host(r.handle, C.c_void_p(r.target_ptr()), C.c_void_p(outs[0].data_ptr()), C.c_size_t(size * size * 4))
def blocking():
    rc = r.lib.vb_render(r.handle, scene_h.data_ptr(), scene_h.numel() * 4, C.byref(lay), None, 512, 0, atlas.ctypes.data, 0, 0, C.byref(ps), outs[0].data_ptr(), 0, C.byref(fs))
    assert rc == 0
k = [0]
def streaming():
    rc = r.lib.vb_render_begin(r.handle, scene_h.data_ptr(), scene_h.numel() * 4, C.byref(lay), None, 512, 0, atlas.ctypes.data, 0, 0, C.byref(ps), outs[k[0] & 1].data_ptr(), C.byref(fs))
    k[0] += 1
    assert rc == 0

blocking()
print("upload (H2D %.1f MB)      %.3f ms" % (scene_h.numel() * 4 / 1e6, T(upload)))
print("render_resident           %.3f ms" % T(resident), " device total_ms %.3f" % fs.total_ms)
print("read-back 64 MiB          %.3f ms" % T(readback))
print("vb_render (blocking)      %.3f ms" % T(blocking), " device total_ms %.3f" % fs.total_ms, {k2: round(v, 3) for k2, v in fs.as_dict()["stage_ms"].items() if k2 in ("fine", "flatten")})
for nb in (8,):
    r.lib.vb_set_readback_bands(r.handle, nb)
    print("vb_render, %d band(s)      %.3f ms" % (nb, T(blocking)), " device total_ms %.3f fine %.3f" % (fs.total_ms, fs.as_dict()["stage_ms"]["fine"]))
print("vb_render_begin (stream)  %.3f ms" % T(streaming), " device total_ms %.3f" % fs.total_ms)
print("   streamed stage_ms", {k2: round(v, 3) for k2, v in fs.as_dict()["stage_ms"].items()})
resident()
print("   resident stage_ms", {k2: round(v, 3) for k2, v in fs.as_dict()["stage_ms"].items()})
r.lib.vb_readback_wait(r.handle)
