#!/usr/bin/env python
"""Experiment (not product): where does the enumerating search kernel spend its time?

Runs the bench workload's search on an S3_ITEM_STATS build of the library
(`python soap3-dp_b200/build.py --variant=stats -DS3_ITEM_STATS`), reads back the number of
LF-mapping steps every (read, case) item took in s3_search_kernel, and prints the distribution
per case, plus the launch time at several batch sizes (a time that does not shrink with the
batch is a tail: the longest single item).

usage (GPU box): S3_LIB_PATH=$PWD/soap3-dp_b200/libsoap3dp_b200.stats.so python tools/search_tail.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from soap3dp_b200 import api, formats  # noqa: E402


def main():
    genome_bp = int(os.environ.get("S3_GENOME_BP", 3_100_000_000))
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    genome, host = bench.get_index(genome_bp, 3, device, 0, 1)
    gi = bench.upload_index(host, 0)
    lib = api.load_library()
    has_stats = hasattr(lib, "s3_debug_item_stats")
    stream = torch.cuda.ExternalStream(gi.stream, device=device)
    out = {}
    for pairs in (524288, 262144, 131072, 65536):
        b = bench.make_batch(genome, pairs, 100, seed=100)
        allowed, ncases = formats.SA_RANGES_ROUND1[2], formats.NUM_CASES[2]
        wpa = 2 * allowed
        answers = [torch.empty(formats.ceil32(b.n) * wpa, dtype=torch.int32, device=device) for _ in range(ncases)]
        ptrs = [a.data_ptr() for a in answers]
        torch.cuda.synchronize()               # the library runs on its own stream
        for _ in range(2):
            api.search_round1_device(gi, b.queries.data_ptr(), b.lens.data_ptr(), b.n, b.wpq, 2, ncases, allowed, wpa, ptrs, 0)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(3):
                api.search_round1_device(gi, b.queries.data_ptr(), b.lens.data_ptr(), b.n, b.wpq, 2, ncases, allowed, wpa, ptrs, 0)
            e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[f"ms_reads_{b.n}"] = ms
        print(f"reads {b.n}: search {ms:.3f} ms", flush=True)
        if has_stats and pairs == 524288:
            n_items = b.n * ncases
            st = np.zeros(n_items, np.uint32)
            hard = C.c_uint32(0)
            lib.s3_debug_item_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
            api._check(lib.s3_debug_item_stats(gi.handle, st.ctypes.data, n_items, C.byref(hard)), "stats")
            print(f"items {n_items}, left to the enumerating kernel {hard.value} ({100.0 * hard.value / n_items:.2f} %)")
            out["hard_items"] = int(hard.value)
            for c in range(ncases):
                s = st[c * b.n:(c + 1) * b.n]
                nz = s[s > 0]
                if nz.size == 0:
                    continue
                pc = np.percentile(nz, [50, 90, 99, 99.9, 99.99])
                print(f"case {c}: enumerated {nz.size}, steps sum {int(nz.sum())}, p50 {pc[0]:.0f} p90 {pc[1]:.0f} p99 {pc[2]:.0f} "
                      f"p99.9 {pc[3]:.0f} p99.99 {pc[4]:.0f} max {int(nz.max())}")
                out[f"case{c}"] = {"n": int(nz.size), "sum": int(nz.sum()), "pct": [float(x) for x in pc], "max": int(nz.max())}
            allnz = st[st > 0]
            srt = np.sort(allnz)[::-1]
            print("top 20 items (steps):", srt[:20].tolist())
            for thr in (256, 512, 1024, 2048, 4096):
                m = allnz > thr
                print(f"items with more than {thr} steps: {int(m.sum())}, holding {100.0 * allnz[m].sum() / max(allnz.sum(), 1):.1f} % of all steps")
            out["top"] = srt[:50].tolist()
            out["total_steps"] = int(allnz.sum())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "search_tail.json"), "w"))


if __name__ == "__main__":
    main()
