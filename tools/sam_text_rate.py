"""Host throughput of s3_sam_single_batch_text (records + text lines per second) on this machine's cores -- the writers are host code like
the reference's output threads; this says how many reads per second they turn into SAM text next to what the GPU path aligns.
    python tools/sam_text_rate.py [reads] [threads ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402  (pack_text, pointer helpers; no oracle is called)
from soap3dp_b200 import api  # noqa: E402
from test_cpu_sam import Config, Genome, SamReads, Segment, U8P  # noqa: E402


def main():
    num = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
    threads = [int(x) for x in sys.argv[2:]] or [1, 2, 4, 8, 16, 0]
    lib = api.load_library()
    lib.s3_sam_single_batch_text.restype = C.c_int
    lib.s3_free.restype = None
    lib.s3_free.argtypes = [C.c_void_p]
    rng = np.random.default_rng(1)
    n, L, row = 50_000_000, 100, 104
    G = rng.integers(0, 4, n).astype(np.uint8)
    pac = helpers.pack_text(G)
    amb = np.zeros((n >> 18) + 2, np.uint32)
    chr_end = np.array([n - 1], np.uint32)
    segs = (Segment * 1)(Segment(0, 1, 0xFFFFFFFF))
    cnames = (C.c_char_p * 1)(b"chrSynth")
    gen = Genome(helpers.u32p(pac), n, segs, 1, helpers.u32p(amb), helpers.u32p(chr_end), 1, cnames)
    pos0 = rng.integers(0, n - L, num)
    bases = np.zeros((num, row), np.uint8)
    bases[:, :L] = G[pos0[:, None] + np.arange(L)[None, :]]
    sub = rng.random((num, L)) < 0.01
    bases[:, :L][sub] = (bases[:, :L][sub] + 1) & 3
    quals = np.ascontiguousarray(rng.integers(2, 41, (num, row)).astype(np.uint8))
    lens = np.full(num, L, np.uint32)
    names = (C.c_char_p * num)(*[b"read%d" % r for r in range(num)])
    rd = SamReads(bases.ctypes.data_as(U8P), C.cast(quals.ctypes.data, C.c_char_p), row, lens.ctypes.data_as(C.POINTER(C.c_uint32)), names)
    counts = rng.choice([int(x) for x in os.environ.get("CNT","0,1,1,1,1,2,3").split(",")], num)                       # 1/7 unmapped, most reads unique
    off = np.zeros(num + 1, np.uint32)
    off[1:] = np.cumsum(counts)
    tot = int(off[-1])
    pos = rng.integers(0, n - L, tot).astype(np.uint32)
    if tot: pos[off[:-1][counts > 0]] = pos0[counts > 0]                          # the first occurrence is where the read came from (MD has content)
    flags = np.ascontiguousarray(np.stack([np.ones(tot), rng.integers(0, 3, tot)], 1).astype(np.uint8)) if tot else np.zeros((1, 2), np.uint8)
    cfg = Config(1, 0, 1, -2, 1, 40, 1, 1, 1, 1000, b"rg")
    for t in threads:
        text, size = C.c_void_p(), C.c_uint64()
        t0 = time.perf_counter()
        rc = lib.s3_sam_single_batch_text(C.byref(gen), C.byref(cfg), C.byref(rd), C.c_uint64(num), helpers.u32p(off), helpers.u32p(pos), flags.ctypes.data_as(U8P), t,
                                          C.byref(text), C.byref(size))
        dt = time.perf_counter() - t0
        assert rc == 0, lib.s3_last_error()
        lib.s3_free(text)
        print(f"threads {t if t else os.cpu_count()}: {num / dt / 1e6:.2f} M reads/s, {size.value / dt / 1e6:.0f} MB/s of SAM text ({size.value / num:.0f} bytes per line)")


if __name__ == "__main__":
    main()
