"""Device->host (and host->device) copy bandwidth with 1, 2, 4 concurrent streams (pinned memory)."""
import time, torch
n = 64 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
for direction in ("d2h", "h2d"):
    for k in (1, 2, 4, 8):
        streams = [torch.cuda.Stream() for _ in range(k)]
        chunk = n // k
        def go():
            for i, s in enumerate(streams):
                with torch.cuda.stream(s):
                    a, b = i * chunk, (i + 1) * chunk
                    if direction == "d2h": host[a:b].copy_(dev[a:b], non_blocking=True)
                    else: dev[a:b].copy_(host[a:b], non_blocking=True)
        for _ in range(3): go()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps): go()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(direction, "streams", k, "%.1f GB/s" % (n / dt / 1e9), "%.3f ms per 64 MiB" % (dt * 1e3))
