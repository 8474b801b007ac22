#!/usr/bin/env python
"""bench.py -- frames/s on the paris-30k-like scene at 4096x4096 MSAA16 (BASELINE.json configs[2]) and the
fine-stage HBM roofline, next to the CPU baseline (the oracle = restated reference CPU shaders).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W          (one rank per GPU, bin-row stripes, no data-path collective)
    python bench.py --impl reference ...                (the CPU arm: oracle on the host cores)

One step = one frame of the hot path (pathtag .. fine) over the synthetic scene.
`value`  : frames/s with the packed scene already resident in HBM (vb_render_resident).
`e2e`    : frames/s through the C ABI with pinned HOST buffers (streaming `vb_render_begin`; the blocking one-call
           `vb_render` figure is reported next to it): scene H2D + render +
           full image D2H inside the timed region.
Timing   : CUDA events on the renderer's stream around every step; L2 is flushed (256 MiB write) between
           steps outside the event pairs; max over ranks; clocks sampled with nvidia-smi during the run.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="paris-like-30k 4096x4096 MSAA16", n_paths=30000, size=4096, seed=30000, aa=2)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_scene(args):
    from vello_b200 import scenes
    from vello_b200.encoding import resolve
    t = time.time()
    if args.scene == "tiger":        # BASELINE.json configs[1]: tiger at args.size x args.height
        sc = scenes.tiger(args.size, args.height or args.size)
    elif args.scene == "beziers":    # configs[4]: cubic paths + nested clips
        sc = scenes.beziers_clips(args.paths, max(1, args.paths // 100), args.size, seed=100000)
    else:                            # configs[2] (default, the headline workload)
        sc = scenes.paris_like(args.paths, args.size, args.seed)
    packed = resolve(sc.encoding)
    return packed, time.time() - t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the single JSON line (no NCCL version banner)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


from vello_b200.stripes import stripe_for  # noqa: E402  (bin-row stripe of a rank)


def run_cpu_arm(args, packed, as_reference):
    """Time the oracle (restated reference CPU shaders + fine) on the host cores."""
    from oracle.vbo import Oracle
    from vello_b200.encoding import BLACK
    cores = os.cpu_count() or 1
    o = Oracle(threads=cores)
    # bounded sample: one full frame is ~3-5 s of CPU work on 8 cores, so the reference arm runs at most
    # one warm-up frame and stops after ~2 minutes of timed frames
    steps = max(1, min(args.steps, 20) if as_reference else 1)
    warm = min(args.warmup, 1) if as_reference else 0
    times, serial, fine = [], [], []
    w, h = args.size, args.height or args.size
    frame = None
    for i in range(warm + steps):
        t = time.perf_counter()
        o.bind(packed, w, h, BLACK.premul_rgba8_u32(), args.aa)
        o.run("pathtag", "path_tiling")  # one host thread, like the reference's CPU shaders (RendererOptions::use_cpu)
        t1 = time.perf_counter()
        frame = o.run("fine", "fine")    # the reference has no CPU fine; ours runs one thread per tile row, all cores
        t2 = time.perf_counter()
        if i >= warm:
            times.append(t2 - t)
            serial.append(t1 - t)
            fine.append(t2 - t1)
        if sum(times) > 120:
            break
    fps = len(times) / sum(times)
    fine_threads = min(cores, 1024, (h + 15) // 16)
    return fps, dict(value=fps, unit="frames/s", cores=cores, kind="port",
                     sample=f"{len(times)} full frame(s) of the workload; pathtag..path_tiling on ONE thread as the reference's CPU shaders run "
                            f"({1000 * sum(serial) / len(serial):.0f} ms/frame), fine on {fine_threads} threads ({1000 * sum(fine) / len(fine):.0f} ms/frame)"), \
        1000.0 * sum(times) / len(times), len(times), frame


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to fd 1 behind Python's back (NCCL's version banner, for one). Everything that is not the JSON
    # line is sent to stderr: fd 1 is re-pointed at fd 2 for the run and the line is written to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=WORKLOAD["n_paths"])
    ap.add_argument("--size", type=int, default=WORKLOAD["size"])
    ap.add_argument("--seed", type=int, default=WORKLOAD["seed"])
    ap.add_argument("--aa", type=int, default=WORKLOAD["aa"])
    ap.add_argument("--scene", default="paris", choices=["paris", "tiger", "beziers"], help="paris = the headline workload")
    ap.add_argument("--height", type=int, default=0, help="frame height (default: --size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="just run warmup+steps resident frames (for ncu)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if not args.profile_only else args.warmup

    H = args.height or args.size
    wl = {"paris": f"paris-like-{args.paths // 1000}k", "tiger": "Ghostscript tiger", "beziers": f"beziers-{args.paths // 1000}k+clips"}[args.scene]
    config = {"workload": f"{wl} {args.size}x{H} " + ["Area", "MSAA8", "MSAA16"][args.aa],
              "n_paths": args.paths, "seed": args.seed, "parallelism": f"bin-row stripes x{args.gpus}",
              "l2": "flushed between steps (256 MiB write) outside the per-step CUDA-event pairs"}

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        packed, _ = build_scene(args)
        fps, cb, ms, n, _ = run_cpu_arm(args, packed, True)
        emit({"impl": "reference", "metric": "frames/sec paris-30k@4K", "value": fps, "unit": "frames/s",
                          "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": cb, "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0})
        return

    rank, world, local, dist = dist_setup(args.gpus)
    import torch
    from vello_b200.config import RenderParams
    from vello_b200.encoding import BLACK
    from vello_b200.renderer import Renderer, RendererOptions, FrameStats, _Layout, _Params

    packed, gen_s = build_scene(args)
    params = RenderParams(BLACK, args.size, H, args.aa)
    bin_rows = stripe_for(rank, world, H) if world > 1 else (0, 0)
    torch.cuda.set_device(local)
    r = Renderer(RendererOptions(device=local))
    r.upload(packed)
    h0, h1 = r.stripe_rows(params, bin_rows)
    out = torch.empty((max(h1 - h0, 1), args.size, 4), dtype=torch.uint8, device=f"cuda:{local}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    stream = torch.cuda.ExternalStream(r.stream, device=f"cuda:{local}")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-scene throughput ("value") -------------------------------------------------
    st = r.render_resident(params, out.data_ptr(), bin_rows)  # sizes the arenas (may retry)
    for _ in range(args.warmup):
        r.render_resident(params, out.data_ptr(), bin_rows)
    if args.profile_only:
        for _ in range(args.steps):
            r.render_resident(params, out.data_ptr(), bin_rows)
        return
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    t_wall = time.perf_counter()
    for a, b in evs:
        flush.fill_(1)          # L2 flush on the default stream ...
        torch.cuda.synchronize()  # ... finished before the step starts
        a.record(stream)
        r.enqueue(params, out.data_ptr(), bin_rows)
        b.record(stream)
        s = r.finish()
        assert s.failed == 0
        launches += int(s.kernel_launches)
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    fps = args.steps / (total_ms / 1000.0)

    # ---- end-to-end through vb_render with pinned host buffers ("e2e") ---------------------------
    scene_h = torch.from_numpy(np.ascontiguousarray(packed.scene)).pin_memory()
    ramps_h = torch.from_numpy(np.ascontiguousarray(packed.ramps.reshape(-1))).pin_memory() if packed.ramps.size else None
    atlas_np = np.ascontiguousarray(packed.atlas)
    out_h = torch.empty((max(h1 - h0, 1), args.size, 4), dtype=torch.uint8).pin_memory()
    lay = _Layout(*[int(v) for v in packed.layout.as_array()])
    ps = _Params(BLACK.premul_rgba8_u32(), args.size, H, args.aa, bin_rows[0], bin_rows[1])
    fs = FrameStats()

    out_h2 = torch.empty_like(out_h).pin_memory()
    outs = (out_h, out_h2)

    def e2e_args(o):
        return (r.handle, scene_h.data_ptr(), scene_h.numel() * 4, C.byref(lay), ramps_h.data_ptr() if ramps_h is not None else None, 512,
                packed.ramps.shape[0], atlas_np.ctypes.data, atlas_np.shape[1], atlas_np.shape[0], C.byref(ps), o.data_ptr())

    def e2e_sync_step():  # one blocking call per frame: upload + render + read-back
        rc = r.lib.vb_render(*e2e_args(out_h), 0, C.byref(fs))
        assert rc == 0 and fs.failed == 0

    def e2e_stream(n):  # streaming form: frame k's read-back tail overlaps frame k+1's upload and geometry stages
        for k in range(n):
            rc = r.lib.vb_render_begin(*e2e_args(outs[k & 1]), C.byref(fs))
            assert rc == 0 and fs.failed == 0
        assert r.lib.vb_readback_wait(r.handle) == 0  # every frame's pixels are in host memory when the clock stops

    e2e_steps = max(4, args.steps // 2)
    e2e_fps_by_mode = {}
    for mode in ("sync", "stream"):
        for _ in range(3):
            e2e_sync_step() if mode == "sync" else e2e_stream(2)
        barrier()
        t0 = time.perf_counter()
        if mode == "sync":
            for _ in range(e2e_steps):
                e2e_sync_step()  # vb_render synchronises internally: host wall clock == device completion
        else:
            e2e_stream(e2e_steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_fps_by_mode[mode] = e2e_steps / float(t.item())
    assert torch.equal(out_h, out_h2), "streamed frames differ"
    e2e_fps = e2e_fps_by_mode["stream"]
    h2d = int(packed.scene.nbytes + packed.ramps.nbytes + packed.atlas.nbytes)
    d2h = int((h1 - h0) * args.size * 4 + 32)

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (fine), measured live with per-stage CUDA events ----------
    rt = Renderer(RendererOptions(device=local, timing=True))
    rt.upload(packed)
    rt.render_resident(params, out.data_ptr(), bin_rows)
    stage_ms = {}
    n_t = 10
    for _ in range(n_t):
        flush.fill_(1)
        torch.cuda.synchronize()
        sd = rt.render_resident(params, out.data_ptr(), bin_rows).as_dict()
        for k, v in sd["stage_ms"].items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v / n_t
    ptcl_words, seg_refs, fill_cmds = rt.fine_traffic()  # what the interpreter reads, from each tile's occlusion start
    rt.set_occlusion_cull(False)
    full_words, full_segs, full_fills = rt.fine_traffic()  # the whole command lists, as the reference executes them
    rt.set_occlusion_cull(True)
    px = (h1 - h0) * args.size
    seg_bytes = 24 * seg_refs * (2 if args.aa else 1)  # MSAA reads each segment twice (count + rasterise), fine.wgsl:177,225
    alg_bytes = 4 * px + 4 * ptcl_words + 24 * seg_refs  # each datum once
    fine_s = stage_ms["fine"] / 1000.0
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / fine_s / 1e9
    roofline = {"kernel": f"k_fine<{args.aa}>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes": alg_bytes, "bytes_breakdown": {"pixels": 4 * px, "ptcl": 4 * ptcl_words, "segments": 24 * seg_refs},
                "fine_ms": stage_ms["fine"], "fill_cmds_executed": fill_cmds,
                "whole_list": {"bytes": 4 * px + 4 * full_words + 24 * full_segs, "fill_cmds": full_fills},
                "note": "bytes = pixels + the PTCL words and segments fine reads from each tile's occlusion start (last opaque "
                "full-tile cover, noted by coarse); whole_list = the same count over the complete lists the reference executes. "
                "MSAA16 fine is shared-memory-atomic / ALU bound (SURVEY.md 8d caveat); segments re-read from L2 by the second "
                "MSAA pass are not counted"}
    tp = os.path.join(ROOT, "profiles", "fine_traffic.json")
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    rt.close()

    cpu_baseline, parity = None, {"checked": False}
    if not args.no_cpu_baseline:
        # the CPU arm renders this very frame: keep it and compare the GPU's pixels with it (the oracle is the checker here,
        # after every timed region; it is never on the measured path)
        _, cpu_baseline, _, _, cpu_frame = run_cpu_arm(args, packed, False)
        r.render_resident(params, out.data_ptr(), bin_rows)
        gpu_frame = out.cpu().numpy()
        ref_rows = cpu_frame[h0:h1]
        d = np.abs(gpu_frame.astype(np.int16) - ref_rows.astype(np.int16))
        parity = {"checked": True, "against": "oracle (cpu_baseline frame)", "rows": [int(h0), int(h1)], "max_diff": int(d.max()) if d.size else 0,
                  "differing_channel_values": int((d > 0).sum()), "tolerance": 0 if args.aa else 1}

    line = {"metric": "frames/sec paris-30k@4K", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "vb_render_begin x steps + vb_readback_wait (host scene in, host pixels out, every frame)",
                    "blocking_vb_render_value": e2e_fps_by_mode["sync"]},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
            "stage_ms": stage_ms, "bump": {k: int(getattr(st, k)) for k in ("lines", "tile", "seg_counts", "segments", "ptcl", "binning")},
            "scene_bytes": int(packed.scene.nbytes), "wall_s_timed_region": wall, "scene_build_s": gen_s}
    emit(line)
    r.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
