"""Host-link ceiling of this box: aggregate pinned D2H (and H2D) rate of G GPUs copying concurrently, no kernels.
The end-to-end number of bench.py is bounded by it (the 152 MB of saveat output per 1M Float32 trajectories have to reach
host memory).  One process, one copy stream per GPU:  python tools/host_ceiling.py [MB per GPU]  -> JSON lines for
G = 1, 2, 4, 8 (as many as visible)."""
import json
import sys
import time

import torch

MB = int(sys.argv[1]) if len(sys.argv) > 1 else 152
ndev = torch.cuda.device_count()
bufs = []
for g in range(ndev):
    with torch.cuda.device(g):
        d = torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{g}")
        h = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
        bufs.append((d, h, torch.cuda.Stream(device=g)))
for G in [g for g in (1, 2, 4, 8) if g <= ndev]:
    for direction in ("d2h", "h2d"):
        best = 1e9
        for rep in range(5):
            for g in range(G):
                torch.cuda.synchronize(g)
            t = time.perf_counter()
            for it in range(4):
                for g in range(G):
                    d, h, st = bufs[g]
                    with torch.cuda.stream(st):
                        if direction == "d2h":
                            h.copy_(d, non_blocking=True)
                        else:
                            d.copy_(h, non_blocking=True)
            for g in range(G):
                bufs[g][2].synchronize()
            best = min(best, (time.perf_counter() - t) / 4)
        print(json.dumps({"gpus": G, "direction": direction, "mb_per_gpu": MB, "ms": round(best * 1e3, 3),
                          "aggregate_gbs": round(G * (MB << 20) / best / 1e9, 2), "per_gpu_gbs": round((MB << 20) / best / 1e9, 2)}), flush=True)
