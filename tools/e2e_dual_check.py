import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import _pkg; _pkg.load()
from soap3dp_b200 import api, fmindex, formats, packing, synth
import bench
dev = torch.device("cuda", 0)
G = synth.random_genome(300_000_000, seed=3, device=dev)
idx = fmindex.build_index(G, keep_sa=True)
gi = api.GPUINDEXUpload(idx, device=0, with_text=True, with_sa=True)
gi2 = api.index_clone(gi)
L, pairs = 100, 524288
N = 2 * pairs
par = api.pe_params(read_length=L, max_windows=N // 2)
pe, pe2 = api.PairAligner(gi, N, L, par), api.PairAligner(gi2, N, L, par)
batches = [bench.make_batch(G, pairs, L, seed=100 + s) for s in range(8)]
wpq = batches[0].wpq
def pinned(t):
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h
sets = [(pinned(b.queries), pinned(b.lens)) for b in batches]
def run(al, ss, prefetch, log=None):
    for k, (q, l) in enumerate(ss):
        t0 = time.perf_counter()
        if prefetch and k + 1 < len(ss):
            al.prefetch(ss[k + 1][0].data_ptr(), ss[k + 1][1].data_ptr(), N, wpq)
        t1 = time.perf_counter()
        al.align(q.data_ptr(), l.data_ptr(), N, wpq, copy=False)
        if log is not None: log.append((t1 - t0, time.perf_counter() - t1))
for al in (pe, pe2):
    run(al, sets[:2], True)
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3 / 8
def dual(prefetch, logs=None):
    ta = threading.Thread(target=run, args=(pe, sets[0::2], prefetch, logs[0] if logs else None))
    tb = threading.Thread(target=run, args=(pe2, sets[1::2], prefetch, logs[1] if logs else None))
    ta.start(); tb.start(); ta.join(); tb.join()
print("one thread, prefetch   ms/step", timed(lambda: run(pe, sets, True)))
print("one thread, no prefetch ms/step", timed(lambda: run(pe, sets, False)))
print("two threads, prefetch   ms/step", timed(lambda: dual(True)))
print("two threads, no prefetch ms/step", timed(lambda: dual(False)))
logs = ([], [])
timed(lambda: dual(True, logs))
print("per call (prefetch s, align s) A:", [(round(a*1e3,2), round(b*1e3,2)) for a, b in logs[0]])
print("per call B:", [(round(a*1e3,2), round(b*1e3,2)) for a, b in logs[1]])
def dev_run(al, bs):
    for b in bs: al.align_device(b.queries.data_ptr(), b.lens.data_ptr(), N, wpq)
def dual_dev():
    ta = threading.Thread(target=dev_run, args=(pe, batches[0::2])); tb = threading.Thread(target=dev_run, args=(pe2, batches[1::2]))
    ta.start(); tb.start(); ta.join(); tb.join()
print("device one", timed(lambda: dev_run(pe, batches)), "device two", timed(dual_dev))
