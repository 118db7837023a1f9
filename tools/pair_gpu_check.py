"""Torch-free GPU check of s3_pair_occurrences against oracle/pair_oracle.c (for a GPU session with seconds to spare;
the regular test is tests/test_zz_pair_gpu.py).  Needs tools/_tiny_index.npz, written by `python tools/pair_gpu_check.py
--make-index` on a box with torch.  Prints PASS / FAIL lines and writes them to gpurun_out/pair_check.txt."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NPZ = os.path.join(ROOT, "tools", "_tiny_index.npz")


def make_index():
    import _pkg
    _pkg.load()
    from soap3dp_b200 import fmindex, synth
    idx = fmindex.build_index(synth.random_genome(20_000, seed=3))
    n = lambda x: np.ascontiguousarray((x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)).astype(np.uint32))
    np.savez(NPZ, bwt=n(idx.fwd.bwt_words), occ=n(idx.fwd.occ), rbwt=n(idx.rev.bwt_words), rocc=n(idx.rev.occ),
             meta=np.array([idx.fwd.num_occ, idx.fwd.inverse_sa0, idx.rev.inverse_sa0, idx.text_length], np.uint64))
    print("wrote", NPZ)


def lists(rng, num_pairs, max_occ=16, near_edges=False):
    p1, s1, m1, o1, p2, s2, m2, o2 = [], [], [], [0], [], [], [], [0]
    for _ in range(num_pairs):
        n1, n2 = int(rng.integers(0, max_occ)), int(rng.integers(0, max_occ))
        base = int(rng.integers(0, 600)) if near_edges and rng.random() < 0.5 else \
            (0xFFFFFFFF - int(rng.integers(0, 900)) if near_edges else int(rng.integers(1000, 1 << 20)))
        a = (base + rng.integers(-300, 300, n1)) & 0xFFFFFFFF
        b = (base + rng.integers(-700, 700, n2)) & 0xFFFFFFFF
        if n1 > 2:
            a[1] = a[0]
        p1 += list(a); p2 += list(b)
        s1 += list(rng.integers(1, 3, n1)); s2 += list(rng.integers(1, 3, n2))
        m1 += list(rng.integers(0, 5, n1)); m2 += list(rng.integers(0, 5, n2))
        o1.append(len(p1)); o2.append(len(p2))
    f = lambda x, t: np.ascontiguousarray(np.array(x, dtype=np.int64).astype(t))
    return (f(p1, np.uint32), f(s1, np.uint8), f(m1, np.uint8), f(o1, np.uint64),
            f(p2, np.uint32), f(s2, np.uint8), f(m2, np.uint8), f(o2, np.uint64))


def main():
    import _pkg
    _pkg.load()
    from soap3dp_b200 import api
    U32P, U8P, U64P = api.U32P, api.U8P, api.U64P
    z = np.load(NPZ)
    lib = api.load_library()
    h = C.c_void_p()
    u = lambda a: np.ascontiguousarray(a, np.uint32).ctypes.data_as(U32P)
    arrs = [np.ascontiguousarray(z[k], np.uint32) for k in ("bwt", "occ", "rbwt", "rocc")]
    meta = [int(x) for x in z["meta"]]
    api._check(lib.s3_index_upload(*(a.ctypes.data_as(U32P) for a in arrs), meta[0], meta[1], meta[2], meta[3], None, None, 0, C.byref(h)),
               "s3_index_upload")
    gi = api.GpuIndex(h.value, meta[3])
    olib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    olib.s3o_pair_occurrences.restype = C.c_uint64
    olib.s3o_pair_occurrences.argtypes = [U32P, U8P, U8P, U64P, U32P, U8P, U8P, U64P, U32P, C.c_uint64, C.c_int32, C.c_int32, C.c_int, C.c_int,
                                          C.c_int, U64P, U32P, U32P, U32P, U8P, C.c_uint64, U32P, U32P, U32P]
    rng = np.random.default_rng(4)
    lines = []
    for legs in ((1, 2), (2, 1), (1, 1)):
        for one in (0, 1):
            for near, npairs in ((False, 3000), (True, 500)):
                L = lists(rng, npairs, near_edges=near)
                pl = rng.integers(60, 151, npairs).astype(np.uint32)
                got = api.pair_occurrences(gi, *L, pl, 200, 500, *legs, bool(one))
                b8 = lambda x: x.ctypes.data_as(U8P)
                args = (u(L[0]), b8(L[1]), b8(L[2]), L[3].ctypes.data_as(U64P), u(L[4]), b8(L[5]), b8(L[6]), L[7].ctypes.data_as(U64P),
                        u(pl), npairs, 200, 500, legs[0], legs[1], one)
                offs = np.zeros(npairs + 1, np.uint64)
                tot = olib.s3o_pair_occurrences(*args, offs.ctypes.data_as(U64P), None, None, None, None, 0, None, None, None)
                a, b, ins = np.zeros(tot, np.uint32), np.zeros(tot, np.uint32), np.zeros(tot, np.uint32)
                fl = np.zeros((tot, 4), np.uint8)
                opt, sub, st = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32), np.zeros((npairs, 32), np.uint32)
                olib.s3o_pair_occurrences(*args, offs.ctypes.data_as(U64P), u(a), u(b), u(ins), b8(fl), tot, u(opt), u(sub), u(st))
                want = dict(offsets=offs, pos1=a, pos2=b, insertion=ins, flags=fl, optimal=opt, suboptimal=sub, stats=st)
                bad = [k for k in want if not np.array_equal(want[k], got[k])]
                lines.append(f"{'PASS' if not bad else 'FAIL ' + ','.join(bad)} legs={legs} one={one} near_edges={near} pairs={tot}")
                print(lines[-1], flush=True)
    # empty batches and lists
    z32, z8 = np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    g = api.pair_occurrences(gi, z32, z8, z8, np.zeros(1, np.uint64), z32, z8, z8, np.zeros(1, np.uint64), z32, 200, 500)
    ok = g["offsets"].tolist() == [0] and len(g["pos1"]) == 0
    z64 = np.zeros(4, np.uint64)
    g = api.pair_occurrences(gi, z32, z8, z8, z64, z32, z8, z8, z64, np.full(3, 100, np.uint32), 200, 500)
    ok = ok and g["offsets"].tolist() == [0, 0, 0, 0] and bool((g["optimal"] == 0xFFFFFFFF).all())
    L = list(lists(rng, 200))
    L[4:] = [z32, z8, z8, np.zeros(201, np.uint64)]
    g = api.pair_occurrences(gi, *L, np.full(200, 100, np.uint32), 200, 500)
    ok = ok and int(g["offsets"][-1]) == 0 and bool((g["suboptimal"] == 0xFFFFFFFF).all())
    lines.append(f"{'PASS' if ok else 'FAIL'} empty batch / empty lists / second reads without hits")
    print(lines[-1], flush=True)
    # best-hit filters (s3_retain_best) against oracle/retain_oracle.c
    olib.s3o_retain_best.restype = None
    olib.s3o_retain_best.argtypes = [C.c_int, C.c_int32, U32P, U32P, U8P, U8P, U64P, U32P, U8P, U8P, U64P, C.c_uint64,
                                     U64P, U32P, U32P, U8P, U64P, U32P, U8P, U32P]
    b8 = lambda x: x.ctypes.data_as(U8P)
    for mode, cap, nreads, mx in ((0, 0, 20000, (7, 7)), (1, 1, 5000, (7, 7)), (1, 3, 5000, (9, 2)), (1, 40, 5000, (2, 9)), (2, 0, 20000, (7, 7)),
                                  (0, 0, 1048576, (3, 3)), (0, 0, 3, (1, 1))):
        n_sa, n_occ = rng.integers(0, mx[0], nreads), rng.integers(0, mx[1], nreads)
        so, oo = np.zeros(nreads + 1, np.uint64), np.zeros(nreads + 1, np.uint64)
        so[1:], oo[1:] = np.cumsum(n_sa), np.cumsum(n_occ)
        ts_, to_ = int(so[-1]), int(oo[-1])
        sl = rng.integers(0, 1 << 31, ts_).astype(np.uint32)
        sr = (sl + rng.integers(0, 40, ts_)).astype(np.uint32)
        ss, sm = rng.integers(1, 3, ts_).astype(np.uint8), rng.integers(0, 5, ts_).astype(np.uint8)
        op_, os_, om = rng.integers(0, 1 << 32, to_).astype(np.uint32), rng.integers(1, 3, to_).astype(np.uint8), rng.integers(0, 5, to_).astype(np.uint8)
        import time
        t0 = time.perf_counter()
        got = api.retain_best(gi, mode, sl, sr, ss, sm, so, op_, os_, om, oo, cap)
        t_gpu = time.perf_counter() - t0
        w_so, w_oo = np.zeros(nreads + 1, np.uint64), np.zeros(nreads + 1, np.uint64)
        w_l, w_r, w_sf = np.zeros(ts_, np.uint32), np.zeros(ts_, np.uint32), np.zeros((ts_, 2), np.uint8)
        w_p, w_of, w_n = np.zeros(to_, np.uint32), np.zeros((to_, 2), np.uint8), np.zeros(nreads, np.uint32)
        t0 = time.perf_counter()
        olib.s3o_retain_best(mode, cap, u(sl), u(sr), b8(ss), b8(sm), so.ctypes.data_as(U64P), u(op_), b8(os_), b8(om), oo.ctypes.data_as(U64P), nreads,
                             w_so.ctypes.data_as(U64P), u(w_l), u(w_r), b8(w_sf), w_oo.ctypes.data_as(U64P), u(w_p), b8(w_of), u(w_n))
        t_cpu = time.perf_counter() - t0
        ks, ko = int(w_so[-1]), int(w_oo[-1])
        want = dict(sa_off=w_so, sa_l=w_l[:ks], sa_r=w_r[:ks], sa_flags=w_sf[:ks], occ_off=w_oo, occ_pos=w_p[:ko], occ_flags=w_of[:ko], num=w_n)
        bad = [k for k in want if not np.array_equal(want[k], got[k])]
        lines.append(f"{'PASS' if not bad else 'FAIL ' + ','.join(bad)} retain mode={mode} cap={cap} reads={nreads}: {ts_} ranges + {to_} occurrences -> "
                     f"{ks} + {ko} kept; call {1e3 * t_gpu:.1f} ms (first call of a size includes module load), oracle {1e3 * t_cpu:.1f} ms")
        print(lines[-1], flush=True)
    # bench size: the 524,288 read pairs of one bench step, 1-3 occurrences per read around a common locus
    import time
    npairs = 524288
    c1, c2 = rng.integers(1, 4, npairs), rng.integers(1, 4, npairs)
    o1, o2 = np.zeros(npairs + 1, np.uint64), np.zeros(npairs + 1, np.uint64)
    o1[1:], o2[1:] = np.cumsum(c1), np.cumsum(c2)
    base = rng.integers(1000, 3_000_000_000, npairs)
    p1 = (np.repeat(base, c1) + rng.integers(-40, 40, int(o1[-1]))).astype(np.uint32)
    p2 = (np.repeat(base, c2) + rng.integers(150, 420, int(o2[-1]))).astype(np.uint32)
    s1, s2 = rng.integers(1, 3, len(p1)).astype(np.uint8), rng.integers(1, 3, len(p2)).astype(np.uint8)
    m1, m2 = rng.integers(0, 3, len(p1)).astype(np.uint8), rng.integers(0, 3, len(p2)).astype(np.uint8)
    big = (p1, s1, m1, o1, p2, s2, m2, o2)
    pl = np.full(npairs, 100, np.uint32)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        got = api.pair_occurrences(gi, *big, pl, 200, 500, 1, 2, False)
        ts.append(time.perf_counter() - t0)
    b8 = lambda x: x.ctypes.data_as(U8P)
    args = (u(p1), b8(s1), b8(m1), o1.ctypes.data_as(U64P), u(p2), b8(s2), b8(m2), o2.ctypes.data_as(U64P), u(pl), npairs, 200, 500, 1, 2, 0)
    tot = int(got["offsets"][-1])
    offs = np.zeros(npairs + 1, np.uint64)
    a, b, ins, fl = np.zeros(tot, np.uint32), np.zeros(tot, np.uint32), np.zeros(tot, np.uint32), np.zeros((tot, 4), np.uint8)
    opt, sub, st = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32), np.zeros((npairs, 32), np.uint32)
    t0 = time.perf_counter()
    olib.s3o_pair_occurrences(*args, offs.ctypes.data_as(U64P), u(a), u(b), u(ins), b8(fl), tot, u(opt), u(sub), u(st))
    t_cpu = time.perf_counter() - t0
    want = dict(offsets=offs, pos1=a, pos2=b, insertion=ins, flags=fl, optimal=opt, suboptimal=sub, stats=st)
    bad = [k for k in want if not np.array_equal(want[k], got[k])]
    lines.append(f"{'PASS' if not bad else 'FAIL ' + ','.join(bad)} bench size: {npairs} read pairs, {len(p1)} + {len(p2)} occurrences -> {tot} pairs; "
                 f"s3_pair_occurrences (pageable host arrays in and out, ctypes wrapper included) {1e3 * min(ts[1:]):.1f} ms "
                 f"= {npairs / min(ts[1:]) / 1e6:.1f} M read pairs/s; oracle port on one host core {1e3 * t_cpu:.1f} ms")
    print(lines[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "pair_check.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    make_index() if "--make-index" in sys.argv else main()
