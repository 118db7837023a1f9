"""Torch-free parity + timing of the DP step on a GPU (numpy + ctypes + the CUDA runtime only), for iterating on the DP
kernels in 15-25 s GPU slots.  `--make-case [n]` (needs torch and the oracle; run where they are) writes a mate-rescue
batch shaped like bench.py's (100-base reads, 401-column windows, n alignments, default 65,536) and the oracle's outputs
to tools/_dp_case.npz.  Without flags: uploads the batch, runs s3_dp_align_device `S3_DP_CHECK_STEPS` (default 6) times with
the library's timing hooks on, compares scores, hit locations, tie counts and traced patterns with the oracle's, and
prints per-kernel milliseconds and GCUPS; also written to gpurun_out/dp_check.txt."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NPZ = os.path.join(ROOT, "tools", "_dp_case.npz")
SCORES = (1, -2, -3, -1)
FIELDS = ("dna", "dna_len", "read", "read_len", "cutoff", "clip_lt", "clip_rt", "anchor_l", "anchor_r")


def make_case(n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from soap3dp_b200 import synth
    G = synth.random_genome(4_000_000, seed=41)
    b = helpers.make_dp_batch(G, n, 100, "rescue", seed=2, indel_rate=0.006)
    t0 = time.time()
    sc, hit, cnt, pat, cells = helpers.oracle_dp(helpers.load_oracle_dp(), b, SCORES)
    print(f"oracle: {cells / 1e9:.2f} G cells in {time.time() - t0:.1f} s")
    np.savez_compressed(NPZ, meta=np.array([b.n, b.max_read, b.max_dna, b.pat_len, cells], np.int64),
                        scores=sc, hit=hit, cnt=cnt, pattern=pat, **{k: getattr(b, k) for k in FIELDS})
    print("wrote", NPZ, os.path.getsize(NPZ) >> 20, "MiB")


class Cuda:
    """the few CUDA runtime calls needed, through ctypes (the product library has already loaded libcudart)"""
    def __init__(self):
        self.rt = C.CDLL("libcudart.so.12")
        self.rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        self.rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        self.rt.cudaStreamSynchronize.argtypes = [C.c_void_p]
        self.rt.cudaFree.argtypes = [C.c_void_p]

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed with CUDA error {rc}")

    def upload(self, a):
        a = np.ascontiguousarray(a)
        p = C.c_void_p()
        self.check(self.rt.cudaMalloc(C.byref(p), max(a.nbytes, 4)), "cudaMalloc")
        self.check(self.rt.cudaMemcpy(p, a.ctypes.data_as(C.c_void_p), a.nbytes, 1), "cudaMemcpy H2D")
        return p.value

    def alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.rt.cudaMalloc(C.byref(p), max(nbytes, 4)), "cudaMalloc")
        return p.value

    def download(self, p, like):
        out = np.empty_like(like)
        self.check(self.rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(p), out.nbytes, 2), "cudaMemcpy D2H")
        return out

    def sync(self, stream):
        self.check(self.rt.cudaStreamSynchronize(C.c_void_p(stream)), "cudaStreamSynchronize")


def pattern_bytes(p):
    i = 0
    while i < len(p) and p[i] != 0:
        i += 2 if p[i] == ord('V') else 1
    return i + 1


def main():
    import _pkg
    _pkg.load()
    from soap3dp_b200 import api
    z = np.load(NPZ)
    n, max_read, max_dna, pat_len, cells = (int(x) for x in z["meta"])
    lib = api.load_library()
    al = api.SemiGlobalAligner(max_read, max_dna, n, *SCORES)
    cu = Cuda()
    d_in = {k: cu.upload(z[k]) for k in FIELDS}
    want = {k: z[k] for k in ("scores", "hit", "cnt", "pattern")}
    d_out = {k: cu.alloc(want[k].nbytes) for k in want}
    stream = al.stream
    steps = int(os.environ.get("S3_DP_CHECK_STEPS", 6))

    def step():
        al.align_device(d_in["dna"], d_in["dna_len"], d_in["read"], d_in["read_len"], d_in["cutoff"], d_out["scores"], d_out["hit"],
                        d_out["cnt"], d_out["pattern"], n, d_in["clip_lt"], d_in["clip_rt"], d_in["anchor_l"], d_in["anchor_r"])
    for _ in range(2):
        step()
    cu.sync(stream)
    api.set_timing(al.handle, True, dp=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    cu.sync(stream)
    wall = (time.perf_counter() - t0) / steps
    ms, launches = api.read_timing(al.handle, dp=True)
    got = {k: cu.download(d_out[k], want[k]) for k in want}
    bad = {k: int((got[k][:n] != want[k][:n]).sum()) for k in ("scores", "hit", "cnt")}
    traced = np.nonzero(want["scores"][:n] >= z["cutoff"][:n])[0]
    bad_pat = 0
    for t in traced:
        w = want["pattern"][t * pat_len:(t + 1) * pat_len]
        k = pattern_bytes(w)
        bad_pat += int(not np.array_equal(got["pattern"][t * pat_len:t * pat_len + k], w[:k]))
    ok = not any(bad.values()) and bad_pat == 0
    per = [m / steps for m in ms[:4]]
    lines = [f"{'PASS' if ok else 'FAIL'} DP parity: {n} alignments ({len(traced)} traced), differ: {bad}, patterns {bad_pat}",
             f"DP step {sum(per):.3f} ms (sweep {per[0]:.3f}, second sweep of the traceback windows {per[1]:.3f}, traceback {per[2]:.3f}, "
             f"pass 2 {per[3]:.3f}; launches per step {[c // steps for c in launches[:4]]}), wall {1e3 * wall:.3f} ms; "
             f"{cells / 1e9:.2f} G cells -> {cells / (sum(per) * 1e-3) / 1e9:.0f} GCUPS (sweep alone {cells / (per[0] * 1e-3) / 1e9:.0f})"]
    print("\n".join(lines), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "dp_check.txt"), "w").write("\n".join(lines) + "\n")
    al.freeMemory()


if __name__ == "__main__":
    if "--make-case" in sys.argv:
        i = sys.argv.index("--make-case")
        make_case(int(sys.argv[i + 1]) if len(sys.argv) > i + 1 else 65536)
    else:
        main()
