"""A/B + parity of several builds of libvello_b200.so in ONE process (one GPU call):

    python tools/variant_check.py variants/a.so variants/b.so ...   [--json out.json]

For every library: (1) a set of parity scenes (all three AA modes; strokes, clips, blends, gradients, images, the
headline paris-like frame at full size) is rendered through the C ABI and compared with the CPU oracle (computed once);
(2) the headline frame is timed scene-resident: frame time with timing off, per-stage times with timing on.
Development tool: the oracle is used as the checker only."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vello_b200 import renderer as vr, scenes  # noqa: E402
from vello_b200.config import AA_AREA, AA_MSAA8, AA_MSAA16, RenderParams  # noqa: E402
from vello_b200.encoding import BLACK, Scene, resolve  # noqa: E402


def parity_cases(full):
    cases = []
    def add(name, sc, w, h, aas):
        packed = resolve(sc.encoding) if isinstance(sc, Scene) else sc
        for aa in aas:
            cases.append((f"{name}/aa{aa}", packed, w, h, aa))
    add("tiger", scenes.tiger(512, 512), 512, 512, (AA_AREA, AA_MSAA8, AA_MSAA16))
    for fn in (scenes.stroke_styles, scenes.fill_types, scenes.brushes, scenes.many_clips, scenes.deep_blend, scenes.blend_grid,
               scenes.gradient_extend, scenes.two_point_radial, scenes.image_extend_modes, scenes.longpathdash, scenes.robust_paths):
        s, w, h = fn()
        add(fn.__name__, s, w, h, (AA_AREA, AA_MSAA16) if fn not in (scenes.stroke_styles, scenes.brushes) else (AA_AREA, AA_MSAA8, AA_MSAA16))
    add("beziers_clips", scenes.beziers_clips(3000, 60, 1024), 1024, 1024, (AA_MSAA16, AA_MSAA8))
    if full:
        add("paris30k", scenes.paris_like(30000, 4096), 4096, 4096, (AA_MSAA16,))
    return cases


def check_one(lib, cases, refs, headline, frames):
        hp = RenderParams(BLACK, 4096, 4096, AA_MSAA16)
        vr._lib = None
        os.environ["VELLO_B200_LIB"] = os.path.abspath(lib)
        res = {"lib": lib}
        try:
            r = vr.Renderer()
            bad = []
            worst = 0
            for (name, p, w, h, aa), ref in zip(cases, refs):
                img = r.render_to_texture(p, RenderParams(BLACK, w, h, aa))
                d = int(np.abs(img.astype(np.int32) - ref.astype(np.int32)).max())
                worst = max(worst, d)
                if d > (1 if aa == AA_AREA else 0):
                    bad.append((name, d))
            res["parity_ok"] = not bad
            res["parity_bad"] = bad
            res["max_diff"] = worst
            r.upload(headline)
            for _ in range(5):
                r.render_resident(hp)
            tot = 0.0
            r.lib.vb_last_frame_ms.restype = __import__("ctypes").c_float
            for _ in range(frames):
                r.render_resident(hp)
                tot += float(r.lib.vb_last_frame_ms(r.handle))
            res["frame_ms"] = tot / frames
            assert r.lib.vb_set_timing(r.handle, 1) == 0
            st = {}
            for _ in range(10):
                sd = r.render_resident(hp).as_dict()["stage_ms"]
                for k, v in sd.items():
                    st[k] = st.get(k, 0.0) + v / 10
            r.lib.vb_set_timing(r.handle, 0)
            res["stage_ms"] = {k: round(v, 4) for k, v in st.items()}
            r.close()
        except Exception as e:  # a variant that fails to load / render is reported, the others still run
            res["error"] = repr(e)
        return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--json", default=None)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--no-full", action="store_true")
    ap.add_argument("--worker", default=None, help="internal: pickle with the cases / references")
    a = ap.parse_args()
    import pickle
    import subprocess
    if a.worker:
        with open(a.worker, "rb") as f:
            cases, refs, headline = pickle.load(f)
        print("RESULT " + json.dumps(check_one(a.libs[0], cases, refs, headline, a.frames)), flush=True)
        return
    from oracle.vbo import Oracle
    o = Oracle(threads=max(1, min(32, os.cpu_count() or 1)))
    t0 = time.time()
    cases = parity_cases(not a.no_full)
    refs = [o.render(p, w, h, BLACK.premul_rgba8_u32(), aa) for (_, p, w, h, aa) in cases]
    headline = resolve(scenes.paris_like(30000, 4096).encoding) if a.no_full else cases[-1][1]
    pkl = "/tmp/variant_check.pkl"
    with open(pkl, "wb") as f:
        pickle.dump((cases, refs, headline), f, protocol=4)
    print(f"# {len(cases)} parity cases, oracle + scenes in {time.time() - t0:.1f} s", flush=True)
    results = {}
    for lib in a.libs:
        # one process per library: a build that faults takes only its own CUDA context with it
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), lib, "--worker", pkl, "--frames", str(a.frames)],
                                 capture_output=True, text=True, timeout=120)
            line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")]
            res = json.loads(line[-1][7:]) if line else {"lib": lib, "error": (out.stderr or out.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            res = {"lib": lib, "error": "timeout"}
        results[lib] = res
        print(json.dumps(res), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(results, f, indent=1)
    ok = [v for v in results.values() if v.get("parity_ok")]
    if ok:
        best = min(ok, key=lambda v: v["frame_ms"])
        print("BEST", best["lib"], round(best["frame_ms"], 4), "ms")


if __name__ == "__main__":
    main()
