"""Measures what the box can move between pinned host memory and the GPUs: H2D alone, D2H alone and both at once,
on 1 .. N GPUs concurrently (one host thread and two streams per GPU, 1 GiB buffers, contiguous copies).  This is
the ceiling the end-to-end numbers of bench.py are compared with (`e2e.ceiling_frac`): a cfg2 subject moves
4.22 GB in and 4.87 GB out, so its floor is max(4.22 / h2d, 4.87 / d2h) with both directions running together.

    python tools/gpu/pcie_probe.py [--gpus 1,2,4,8] [--json out.json]
"""
import argparse
import json
import threading
import time

import torch


def probe(ngpu, nbytes=1 << 30, reps=6):
    devs = list(range(ngpu))
    bufs = []
    for d in devs:
        torch.cuda.set_device(d)
        bufs.append(dict(h_in=torch.empty(nbytes, dtype=torch.uint8).pin_memory(), h_out=torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                         d_in=torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"), d_out=torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}"),
                         s1=torch.cuda.Stream(device=d), s2=torch.cuda.Stream(device=d)))
    out = {}
    for mode in ("h2d", "d2h", "both"):
        barrier = threading.Barrier(ngpu + 1)
        times = [0.0] * ngpu

        def work(i):
            b = bufs[i]
            torch.cuda.set_device(i)

            def run(n):
                for _ in range(n):
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(b["s1"]):
                            b["d_in"].copy_(b["h_in"], non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(b["s2"]):
                            b["h_out"].copy_(b["d_out"], non_blocking=True)
                torch.cuda.synchronize(i)
            run(1)
            barrier.wait()
            t = time.perf_counter(); run(reps); times[i] = time.perf_counter() - t
            barrier.wait()
        th = [threading.Thread(target=work, args=(i,)) for i in devs]
        [t.start() for t in th]
        barrier.wait(); barrier.wait()
        [t.join() for t in th]
        dt = max(times)
        gb = nbytes * reps / 1e9
        out[mode] = {"per_gpu_GBps_each_direction": gb / dt, "aggregate_GBps": gb * ngpu * (2 if mode == "both" else 1) / dt}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default=None)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    have = torch.cuda.device_count()
    ns = [int(x) for x in a.gpus.split(",")] if a.gpus else [n for n in (1, 2, 4, 8) if n <= have]
    res = {"gpus_visible": have, "host_cpus": len(__import__("os").sched_getaffinity(0)), "results": {}}
    for n in ns:
        if n > have:
            continue
        r = probe(n)
        res["results"][str(n)] = r
        # floor of one cfg2 subject per GPU (4.218 GB in, 4.873 GB out), both directions concurrent
        both = r["both"]["per_gpu_GBps_each_direction"]
        r["cfg2_subject_floor_ms"] = max(4.218, 4.873) / both * 1e3
        print(f"{n} GPU(s): H2D {r['h2d']['per_gpu_GBps_each_direction']:.1f} GB/s/GPU, D2H {r['d2h']['per_gpu_GBps_each_direction']:.1f}, "
              f"both {both:.1f} each way ({r['both']['aggregate_GBps']:.0f} GB/s aggregate); cfg2 subject floor {r['cfg2_subject_floor_ms']:.0f} ms", flush=True)
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
