"""The box's bare device->host ceiling: N processes (one per GPU), each copying `bytes` from its GPU into its own
page-locked buffer with cudaMemcpyAsync, all starting together, no kernels running.  This is the number the
full-rate outlet copy-back of a multi-GPU solve cannot beat.

  python scripts/d2h_ceiling.py --gpus N [--gb 0.66 5.3] [--reps 5]

Prints one JSON line: per size, per-process GB/s (min / max) and the aggregate (total bytes / slowest process).
torch is used for the plumbing only (pinned allocation, copies, events)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.multiprocessing as mp


def worker(rank, world, sizes, reps, chunk_mb, barrier, q):
    torch.cuda.set_device(rank)
    res = {}
    for gb in sizes:
        n = int(gb * 1e9) // 8 * 8
        d = torch.empty(n, dtype=torch.uint8, device=f"cuda:{rank}")
        d.fill_(1)
        h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h.fill_(0)  # first touch
        s = torch.cuda.Stream()
        best = None
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            barrier.wait()
            t0 = time.perf_counter()
            with torch.cuda.stream(s):
                if chunk_mb:
                    c = chunk_mb << 20
                    for o in range(0, n, c):
                        h[o:o + c].copy_(d[o:o + c], non_blocking=True)
                else:
                    h.copy_(d, non_blocking=True)
            s.synchronize()
            dt = time.perf_counter() - t0
            barrier.wait()
            best = dt if best is None else min(best, dt)
        res[str(gb)] = {"bytes": n, "best_s": best, "gbs": n / best / 1e9}
        del d, h
        torch.cuda.empty_cache()
    q.put((rank, res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gb", type=float, nargs="+", default=[0.66, 5.3])
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--chunk-mb", type=int, default=0, help="issue the copy in chunks of this size (0 = one call)")
    a = ap.parse_args()
    n = min(a.gpus, torch.cuda.device_count())
    mp.set_start_method("spawn", force=True)
    barrier = mp.Barrier(n)
    q = mp.Queue()
    ps = [mp.Process(target=worker, args=(r, n, a.gb, a.reps, a.chunk_mb, barrier, q)) for r in range(n)]
    for p in ps:
        p.start()
    out = dict(q.get() for _ in range(n))
    for p in ps:
        p.join()
    line = {"what": "bare pinned D2H ceiling, one process per GPU, simultaneous start, no kernels", "n_gpus": n,
            "host_cpus": len(os.sched_getaffinity(0)), "chunk_mb": a.chunk_mb, "sizes": {}}
    for gb in a.gb:
        k = str(gb)
        per = [out[r][k]["gbs"] for r in range(n)]
        slow = max(out[r][k]["best_s"] for r in range(n))
        tot = sum(out[r][k]["bytes"] for r in range(n))
        line["sizes"][k] = {"bytes_per_gpu": out[0][k]["bytes"], "per_gpu_gbs_min": min(per), "per_gpu_gbs_max": max(per),
                            "aggregate_gbs": tot / slow / 1e9, "slowest_ms": slow * 1e3}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    sys.exit(main())
