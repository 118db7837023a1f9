#!/usr/bin/env python
"""bench.py -- reads/s of the SOAP3-dp GPU alignment hot path on B200.

Workload (BASELINE.json metric = config 4, one batch per step):
  paired-end 2x100 bp, insert 200-500, FR, against a 3.1 Gbp synthetic genome.
  One STEP = one batch of P pairs (2P reads, default 1,048,576 reads = the reference's
  NUM_BLOCKS*THREADS_PER_BLOCK launch size, definitions.h:75-77) through
    (1) the paired-end chain (s3_pe_align[_device]): GPU-2BWT search, <=2 mismatches (Soap3MisMatchAllow=2 when DP is on,
        SOAP3-DP.cu:210-213; 4 cases x both strands, round-1 slots), answer collection, routing, locate, pairing, mate-rescue
        windows (HalfEndAlgnBatch::pack), semi-global DP, CIGAR runs -- nothing taken from the simulator's truth;
    (2) deep DP (s3_pe_deep_dp = DPForUnalignPairs2) of the pairs (1) left with no occurrence of either read.
  `value`  : queries resident in HBM, CUDA-event timed, S3_IN_FLIGHT (4) batches in flight on handles that share the index.
  `e2e`    : the same steps through the host-pointer C ABI, queries from pinned host memory, every result into host memory,
             H2D + D2H inside the timed region; `e2e.pageable_value`: caller buffers from malloc.
  `--impl reference` : the reference's own CPU search (ProcessReadDoubleStrand2, oracle/_ref) + its DP kernels compiled for the
             host, on all host threads, on a bounded sample of the same workload.
  `--config` : pe100_chain (without (2)), se100_k4 (BASELINE config 2), se150_dp (config 3).

Launch:  python bench.py --gpus N --steps K --warmup W
         (N > 1: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from soap3dp_b200 import api, fmindex, formats, packing, sharding, synth  # noqa: E402

INSERT_LO, INSERT_HI = 200, 500
K_MISMATCH = 2
DP_SCORES = (1, -2, -3, -1)                   # soap3-dp.ini:57-66
RESCUE_FRACTION_HINT = 0.364                  # rescue windows per pair of the GPU arm on this workload (190.8 k per 524,288 pairs): the reference arm's DP share


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------
# index (built on the GPU with torch, cached on local disk in the reference's array format)
# ---------------------------------------------------------------------------
def bind_to_gpu_numa_node(dev_index):
    """One process per GPU: run (and first-touch the pinned host buffers) on the CPUs of the NUMA node the GPU hangs off,
    so that the e2e copies of N ranks do not cross the socket interconnect.  Best effort; returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(dev_index)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                                # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:                                    # noqa: BLE001
        pass
    return None


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cache_dir():
    d = os.environ.get("S3_CACHE", "/tmp/s3_bench_cache")
    os.makedirs(d, exist_ok=True)
    return d


def get_index(n_bp, seed, device, rank, world, repeat_fraction=0.2):
    """-> (genome codes on device, dict of host numpy arrays in the reference format)"""
    tag = os.path.join(cache_dir(), f"idx_{n_bp}_{seed}_{repeat_fraction}")
    names = ["bwt", "occ", "rbwt", "rocc", "meta", "sa", "pac"]
    t0 = time.time()
    genome = synth.random_genome(n_bp, seed=seed, device=device, repeat_fraction=repeat_fraction)
    torch.cuda.synchronize()
    log(f"genome {n_bp} bp generated in {time.time() - t0:.1f}s")
    have = all(os.path.exists(f"{tag}.{x}.npy") for x in names)
    if not have and rank == 0:
        t0 = time.time()
        idx = fmindex.build_index(genome, keep_sa=True, verbose=bool(os.environ.get("S3_VERBOSE")))
        torch.cuda.synchronize()
        log(f"2BWT index built on the GPU in {time.time() - t0:.1f}s")
        arrs = {"bwt": idx.fwd.bwt_words, "occ": idx.fwd.occ, "rbwt": idx.rev.bwt_words, "rocc": idx.rev.occ,
                "pac": idx.packed_text}
        for k, v in arrs.items():
            np.save(f"{tag}.{k}.tmp.npy", v.cpu().numpy().view(np.uint32))
            os.replace(f"{tag}.{k}.tmp.npy", f"{tag}.{k}.npy")
        # the suffix array of the n + 1 BWT rows, as bwt->saValue with SaValueFreq = 1 (soap3-dp-builder.ini:29)
        np.save(f"{tag}.sa.tmp.npy", idx.fwd.sa.cpu().numpy().astype(np.uint32))
        os.replace(f"{tag}.sa.tmp.npy", f"{tag}.sa.npy")
        np.save(f"{tag}.meta.tmp.npy", np.array([idx.fwd.inverse_sa0, idx.rev.inverse_sa0, n_bp], np.int64))
        os.replace(f"{tag}.meta.tmp.npy", f"{tag}.meta.npy")
        del idx, arrs
        torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.barrier()
    host = {k: np.load(f"{tag}.{k}.npy", mmap_mode=None) for k in names}
    return genome, host


def upload_index(host, device):
    lib = api.load_library()
    meta = host["meta"]
    n = int(meta[2])
    out = C.c_void_p()
    num_occ = (n + 127) // 128 + 1
    with_locate = not os.environ.get("S3_NO_CHECK_EXTEND")
    rc = lib.s3_index_upload(api._u32(host["bwt"]), api._u32(host["occ"]), api._u32(host["rbwt"]), api._u32(host["rocc"]),
                             num_occ, int(meta[0]), int(meta[1]), n,
                             api._u32(host["pac"]) if with_locate else None, api._u32(host["sa"]) if with_locate else None,
                             device, C.byref(out))
    api._check(rc, "s3_index_upload")
    return api.GpuIndex(out.value, n)


# ---------------------------------------------------------------------------
# batches
# ---------------------------------------------------------------------------
class Batch:
    pass


def make_batch(genome, pairs, L, seed):
    """P pairs -> interleaved reads (mate1, mate2, mate1, ...) packed on the device"""
    m1, m2, bad = synth.simulate_paired_end(genome, pairs, L, seed=seed, insert_lo=INSERT_LO, insert_hi=INSERT_HI)
    b = Batch()
    b.pairs, b.n, b.L = pairs, 2 * pairs, L
    reads = torch.stack([m1.reads, m2.reads], dim=1).reshape(2 * pairs, L)
    b.reads = reads
    b.pos = torch.stack([m1.pos, m2.pos], dim=1).reshape(-1)
    b.strand = torch.stack([m1.strand, m2.strand], dim=1).reshape(-1)
    b.wpq = formats.word_per_query(L)
    up = formats.ceil32(b.n)
    lens = torch.zeros(up, dtype=torch.int32, device=genome.device)
    lens[:b.n] = L
    b.lens = lens
    b.queries = packing.pack_queries(reads, lens[:b.n], b.wpq)
    return b


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, dev_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                      # noqa: BLE001
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:                       # noqa: BLE001
                pass
            time.sleep(0.02)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------
# CPU arm.  Search: the reference's own CPU search (oracle/_ref/libref_cpu_search.so = ProcessReadDoubleStrand2 on the models of
# SRAModelConstruct, 13-mer lookup tables, check-and-extend: what SOAP3-dp runs on the host for a read the GPU left over).
# DP: the reference has no CPU implementation of it, so its DP kernels compiled for the host (oracle/_ref/libref_dp.so) stand in.
# The reference's search kernels compiled for the host are timed beside it as `kernel_code`.  Without oracle/_ref: the oracle port.
# ---------------------------------------------------------------------------
_REF_CPU = {}
MAX_OUTPUT_PER_READ = 1000


def ref_cpu_handle(host, threads):
    """-> helpers.RefCpuSearch on the bench index (built once: occurrence tables + two 13-mer lookup tables), or None"""
    if "h" in _REF_CPU:
        return _REF_CPU["h"]
    import helpers
    lib = helpers.load_ref_cpu_search()
    h = None
    if lib is not None and "sa" in host and "pac" in host:
        t0 = time.time()
        meta = host["meta"]
        h = helpers.RefCpuSearch(lib, host["bwt"], host["rbwt"], int(meta[0]), int(meta[1]), int(meta[2]), host["pac"], host["sa"], threads)
        log(f"reference CPU index structs (occ tables, 13-mer lookup tables of both directions) built in {time.time() - t0:.1f}s")
    _REF_CPU["h"] = h
    return h


def cpu_arm(host, genome, L, sample_reads, seed, dp_fraction, threads, kernel_code_reads=None, gpu_legs=True, keep_hits=0):
    """-> dict(value reads/s, kind, cores, sample, ...)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    ref_s, ref_d = helpers.load_ref_search(), helpers.load_ref_dp()
    ref_c = ref_cpu_handle(host, threads)
    kind = "reference" if (ref_s is not None and ref_d is not None) else "port"
    pairs = sample_reads // 2
    b = make_batch(genome, pairs, L, seed)
    q = b.queries.cpu().numpy().view(np.uint32)
    lens = b.lens.cpu().numpy().view(np.uint32)
    n = b.n
    ncases = formats.NUM_CASES[K_MISMATCH]
    out = {}

    # --- the reference's CPU search
    t_cpu_search = None
    if ref_c is not None:
        reads = np.ascontiguousarray(b.reads.cpu().numpy().astype(np.uint8))
        t0 = time.perf_counter()
        res = ref_c.search(reads, K_MISMATCH, ncases, MAX_OUTPUT_PER_READ, threads)
        t_cpu_search = time.perf_counter() - t0
        cnt = res["counts"]
        out["cpu_search"] = {"reads": int(n), "seconds": t_cpu_search, "reads_per_s": n / t_cpu_search,
                             "reads_with_a_hit": int((cnt[:, 3] > 0).sum()), "occurrences": int(cnt[:, 3].sum()),
                             "found_by_check_and_extend": int(cnt[:, 2].sum()), "sa_ranges": int(cnt[:, 0].sum())}
        if keep_hits:
            m = min(keep_hits, n)
            out["_hit_reads"] = reads[:m]
            out["_hits"] = ref_c.search(reads[:m], K_MISMATCH, ncases, MAX_OUTPUT_PER_READ, threads, out_cap=64)["hits"]

    # --- the reference's search kernels compiled for the host (or the port)
    class HI:
        pass
    hi = HI()
    hi.bwt, hi.occ, hi.rbwt, hi.rocc = host["bwt"], host["occ"], host["rbwt"], host["rocc"]
    hi.isa0, hi.risa0, hi.n = int(host["meta"][0]), int(host["meta"][1]), int(host["meta"][2])
    allowed = formats.SA_RANGES_ROUND1[K_MISMATCH]
    wpa = 2 * allowed
    nk = n if (kernel_code_reads is None or ref_c is None) else min(n, formats.ceil32(kernel_code_reads))
    bad = np.zeros(formats.ceil32(nk), np.uint8)
    ans = []
    nrank = 0
    qq = q[:formats.ceil32(nk) * b.wpq].copy() if nk < n else q.copy()
    lk = lens[:formats.ceil32(nk)].copy()
    lk[nk:] = 0
    q0 = qq.copy()                                            # the reference's kernels turn the queries around in place
    t0 = time.perf_counter()
    for case in range(ncases):
        a = np.zeros(formats.ceil32(nk) * wpa, np.uint32)
        if kind == "reference":
            nrank += helpers.ref_launch(ref_s, hi, case, qq, lk, nk, b.wpq, a, bad, 0, K_MISMATCH, allowed, wpa, nthreads=threads)
        else:
            nrank += helpers.oracle_launch(helpers.load_oracle(), hi, case, qq, lk, nk, b.wpq, a, bad, 0, K_MISMATCH, allowed, wpa)
        ans.append(a)
    t_kernel_search = time.perf_counter() - t0

    # --- DP on the same fraction of pairs the GPU arm rescues
    m = max(int(pairs * dp_fraction), 32)
    dpb = helpers.make_dp_batch(genome[:4_000_000].cpu(), m, L, "rescue", seed=seed + 1, insert=(INSERT_LO, INSERT_HI))
    t0 = time.perf_counter()
    if kind == "reference":
        dp_out = helpers.ref_dp(ref_d, dpb, DP_SCORES, nthreads=threads)
    else:
        dp_out = helpers.oracle_dp(helpers.load_oracle_dp(), dpb, DP_SCORES)
    t_dp = time.perf_counter() - t0

    # --- the reference's own CUDA kernels compiled for sm_100a, on this GPU, on the same batches: the kernels to beat
    gpu_ref = None
    ref_cu = helpers.load_ref_dp_cuda() if (kind == "reference" and gpu_legs) else None
    if ref_cu is not None and torch.cuda.is_available():
        try:
            helpers.ref_dp_cuda(ref_cu, dpb, DP_SCORES)                       # warm-up
            out_cu, ms_cu = helpers.ref_dp_cuda(ref_cu, dpb, DP_SCORES)
            same = helpers.compare_dp(dpb, out_cu, dp_out, "reference CUDA kernels vs reference host build")
            gpu_ref = {"kind": "reference DP kernels (DV-DPfunctions.cu:35-512, textures -> array reads) compiled for sm_100a, "
                               "8192 alignments per launch as SemiGlobalAligner::performAlignment",
                       "alignments": int(dpb.n), "kernel_ms": ms_cu, "gcups": dpb.n * 400 * L / (ms_cu * 1e-3) / 1e9,
                       "tracebacks_equal_to_host_build": int(same)}
        except Exception as e:                       # noqa: BLE001
            gpu_ref = {"error": str(e)[:200]}
    gpu_ref_s = None
    ref_scu = helpers.load_ref_search_cuda() if (kind == "reference" and gpu_legs) else None
    if ref_scu is not None and torch.cuda.is_available():
        try:
            if ref_scu.ref_search_cuda_upload(helpers.u32p(hi.bwt), helpers.u32p(hi.rbwt), len(hi.bwt), helpers.u32p(hi.occ),
                                              helpers.u32p(hi.rocc), len(hi.occ)) != 0:
                raise RuntimeError("index upload failed")
            warm = min(nk, 65536)
            helpers.ref_search_cuda_round1(ref_scu, hi, q0.copy(), lk, warm, b.wpq, K_MISMATCH, allowed, wpa)
            ans_cu, ms_cu = helpers.ref_search_cuda_round1(ref_scu, hi, q0.copy(), lk, nk, b.wpq, K_MISMATCH, allowed, wpa)
            ref_scu.ref_search_cuda_free()
            equal = all(np.array_equal(formats.answers_view(x, nk, wpa), formats.answers_view(y, nk, wpa)) for x, y in zip(ans_cu, ans))
            flags_path = os.path.join(ROOT, "oracle", "_ref", "libref_search_cuda.so.flags")
            flags = open(flags_path).read().strip() if os.path.exists(flags_path) else "flags not recorded"
            gpu_ref_s = {"kind": "reference search kernels (DV-Kernel.cu, unmodified) compiled for sm_100a (" + flags + "), one launch per "
                                 "case over <= 1,048,576 reads as perform_round1_alignment",
                         "reads": int(nk), "kernel_ms": ms_cu, "reads_per_s": nk / (ms_cu * 1e-3),
                         "slots_equal_to_host_build": bool(equal)}
        except Exception as e:                       # noqa: BLE001
            gpu_ref_s = {"error": str(e)[:200]}

    t_search = t_cpu_search if t_cpu_search is not None else t_kernel_search * (n / nk)
    dp_text = f"{m} mate-rescue DP alignments (the fraction of pairs the GPU arm rescues; the reference has no CPU DP: its DP kernels, DV-DPfunctions.cu:35-512, compiled for the host, OpenMP over alignments)"
    if t_cpu_search is not None:
        sample = (f"{n} reads of the bench workload through the reference's CPU search (ProcessReadDoubleStrand2 per case on SRAModelConstruct's "
                  f"16G models, k<=2, 4 cases, both strands, lookup tables + check-and-extend, MaxOutputPerRead {MAX_OUTPUT_PER_READ}, OpenMP over reads) + " + dp_text)
    elif kind == "reference":
        sample = f"{n} reads (k<=2, 4 cases, both strands) through the reference's search kernels (DV-Kernel.cu) compiled for the host, OpenMP over reads + " + dp_text
    else:
        sample = f"{n} reads + {m} rescue DP alignments through the oracle/ C restatement, single thread"
    out.update({"gpu_reference_dp": gpu_ref, "gpu_reference_search": gpu_ref_s, "value": n / (t_search + t_dp), "unit": "reads/s",
                "cores": threads if kind == "reference" else 1, "kind": kind, "sample": sample,
                "search_path": "ProcessReadDoubleStrand2" if t_cpu_search is not None else "kernel_code",
                "kernel_code": {"what": "the reference's GPU search kernels (DV-Kernel.cu) compiled for the host, OpenMP over reads: the same code "
                                        "path the GPU arm replaces, on CPU cores", "reads": int(nk), "seconds": t_kernel_search,
                                "search_reads_per_s": nk / t_kernel_search, "rank_queries_per_read": nrank / nk,
                                "value_with_this_search": n / (t_kernel_search * (n / nk) + t_dp)},
                "rank_queries_per_read": nrank / nk, "t_search_s": t_search, "t_dp_s": t_dp,
                "dp_gcups": dpb.n * 400 * L / t_dp / 1e9})
    return out


# ---------------------------------------------------------------------------
def cpu_search_parity(gi, cb, L):
    """The first reads of the CPU arm's sample through s3_se_align (search -> collect -> locate on the device) against what the
    reference's own CPU search reported for them: per read the same (position, strand, mismatches) set.  Reads whose round-1
    slot overflowed (flag from the device) or that reach the cap on either side are counted apart, not compared."""
    reads, hits = cb["_hit_reads"], cb["_hits"]
    n = len(hits)
    wpq = formats.word_per_query(L)
    lens = np.zeros(formats.ceil32(n), np.uint32)
    lens[:n] = L
    q = formats.pack_queries(reads, lens[:n], wpq)
    al = api.SingleAligner(gi, n, num_mismatch=K_MISMATCH, max_output_per_read=MAX_OUTPUT_PER_READ, report_best=False)
    try:
        got = al.align(q, lens, n, wpq)
    finally:
        al.free()
    off = got["occ_offsets"]
    same = skipped = with_hits = 0
    for r in range(n):
        a, b = int(off[r]), int(off[r + 1])
        if int(got["read_flags"][r]) or len(hits[r]) >= 64 or b - a >= 64:
            skipped += 1
            continue
        mine = {(int(p), int(f[0]), int(f[1])) for p, f in zip(got["positions"][a:b], got["occ_flags"][a:b])}
        same += mine == set(hits[r])
        with_hits += bool(hits[r])
    return {"reads": n, "compared": n - skipped, "reads_with_hits": with_hits, "equal": same, "bit_exact": bool(same == n - skipped),
            "skipped_overflow_or_over_64_hits": skipped,
            "checker": "the reference's CPU search (ProcessReadDoubleStrand2, oracle/_ref/libref_cpu_search.so) on the full-size index"}


# ---------------------------------------------------------------------------
def chain_parity(gi, host, genome, L, pairs, seed, par, device_index):
    """A sample of the workload through s3_pe_align, compared record by record with the composition of the oracles on the
    host (oracle/pe_chain_oracle.py) at full genome size."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import helpers
    import decode_oracle
    import pe_chain_oracle
    b = make_batch(genome, pairs, L, seed)
    n = b.n
    q = b.queries.cpu().numpy().view(np.uint32)
    lens = b.lens.cpu().numpy().view(np.uint32)
    al = api.PairAligner(gi, n, L, par)
    try:
        got = al.align(q, lens, n, b.wpq)
    finally:
        al.free()

    class HI:
        pass
    hi = HI()
    hi.bwt, hi.occ, hi.rbwt, hi.rocc = host["bwt"], host["occ"], host["rbwt"], host["rocc"]
    hi.isa0, hi.risa0, hi.n = int(host["meta"][0]), int(host["meta"][1]), int(host["meta"][2])
    ref_s = helpers.load_ref_search()
    allowed = formats.SA_RANGES_ROUND1[K_MISMATCH]
    wpa = 2 * allowed
    bad = np.zeros(formats.ceil32(n), np.uint8)
    views = []
    qq = q.copy()
    for case in range(formats.NUM_CASES[K_MISMATCH]):
        a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
        if ref_s is not None:
            helpers.ref_launch(ref_s, hi, case, qq, lens, n, b.wpq, a, bad, 0, K_MISMATCH, allowed, wpa, nthreads=os.cpu_count() or 1)
        else:
            helpers.oracle_launch(helpers.load_oracle(), hi, case, q, lens, n, b.wpq, a, bad, 0, K_MISMATCH, allowed, wpa)
        views.append(formats.answers_view(a, n, wpa))
    max_read = (L // 4 + 1) * 4
    max_dna = INSERT_HI - INSERT_LO + max_read + 1
    ref_d = helpers.load_ref_dp()

    def dp_fn(dna, dna_len, rd, rl, mdna, mread, cutoff, clip_lt, clip_rt, anc_l, anc_r, scores):
        db = helpers.DPBatch(dna, dna_len, rd, rl, mdna, mread, cutoff, clip_lt, clip_rt, anc_l, anc_r)
        if ref_d is not None:
            sc, hit, cnt, pat = helpers.ref_dp(ref_d, db, scores, nthreads=os.cpu_count() or 1)
        else:
            sc, hit, cnt, pat, _ = helpers.oracle_dp(helpers.load_oracle_dp(), db, scores)
        return sc, hit, cnt, pat, db.pat_len
    opar = dict(insert_low=INSERT_LO, insert_high=INSERT_HI, left_leg=1, right_leg=2, max_output_per_read=par.maxOutputPerRead,
                max_hit=par.maxHitNumForDP, keep_second_best=bool(par.keepSecondBest), cutoff=-1, soft_clip_left=par.softClipLeft,
                soft_clip_right=par.softClipRight, max_read=max_read, max_dna=max_dna, scores=DP_SCORES)
    gen = _GenomeView(genome)
    reads = b.reads.cpu().numpy()
    want = pe_chain_oracle.pe_chain(views, allowed, lens[:n], host["sa"], gen, list(reads), opar, helpers.oracle_pair_occurrences, dp_fn,
                                    lambda pat, score, rl, sc4: decode_oracle.decode_one(pat, score, rl, sc4)[0])
    route_ok = bool(np.array_equal(got["route"], want["route"]))
    pair_ok, dp_ok = True, len(got["dp"]) == len(want["dp"])
    for p, w in enumerate(want["pairs"]):
        g = got["pairs"][p]
        if w is None or not w["numPairs"]:
            pair_ok &= int(g["numPairs"]) == 0
            continue
        pair_ok &= all(int(g[k2]) == w[k2] for k2 in ("numPairs", "pos1", "pos2", "insertion", "strand1", "mism1", "strand2", "mism2", "optimalTotal",
                                                     "numOptimal", "suboptimalTotal", "numSuboptimal"))
    traced = 0
    if dp_ok:
        for t, w in enumerate(want["dp"]):
            g = got["dp"][t]
            dp_ok &= all(int(g[k2]) == w[k2] for k2 in ("dpReadID", "alignedPos", "alignedStrand", "alignedMismatches", "dpStrand", "leftOrRight",
                                                       "score", "numSameScore", "dpPos"))
            cig = api.runs_to_cigar(got["runs"][int(g["runOffset"]):int(g["runOffset"]) + int(g["numRuns"])])
            dp_ok &= cig == w["cigar"]
            traced += w["cigar"] != ""
    if not (route_ok and pair_ok and dp_ok):
        log(f"CHAIN PARITY FAILURE at full size: routes {route_ok}, pairings {pair_ok}, rescue alignments {dp_ok}")
    return {"pairs": int(pairs), "routes_bit_exact": route_ok, "pairings_bit_exact": bool(pair_ok), "rescue_alignments": len(want["dp"]),
            "rescue_alignments_bit_exact": bool(dp_ok), "rescue_cigars_compared": int(traced),
            "route_counts": np.bincount(want["route"], minlength=9).tolist(),
            "checker": "oracle/pe_chain_oracle.py over " + ("the reference's kernels compiled for the host (oracle/_ref)" if ref_s is not None and ref_d is not None
                                                           else "the oracle ports")}


class _GenomeView:
    """the genome's base codes for the oracle's window reads, without a 3 GB host copy"""
    def __init__(self, g):
        self.g = g

    def __len__(self):
        return int(self.g.numel())

    def __getitem__(self, sl):
        return self.g[sl].cpu().numpy()


def random_sector_rate(gi, device):
    """independent random 32-byte reads over the index's own bucket array (s3_random_sector_probe): the memory system's
    random-sector rate in this run, the denominator of the search roofline"""
    lib = api.load_library()
    lib.s3_random_sector_probe.restype = C.c_int
    lib.s3_random_sector_probe.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    ms, n = C.c_float(), C.c_uint64()
    api._check(lib.s3_random_sector_probe(gi.handle, 64, C.byref(ms), C.byref(n)), "s3_random_sector_probe")
    api._check(lib.s3_random_sector_probe(gi.handle, 64, C.byref(ms), C.byref(n)), "s3_random_sector_probe")
    return n.value / (ms.value * 1e-3)


def make_se_batch(genome, n, L, seed):
    rs = synth.simulate_single_end(genome, n, L, seed=seed, sub_rate=0.01)
    b = Batch()
    b.n, b.L, b.reads, b.pos, b.strand = n, L, rs.reads, rs.pos, rs.strand
    b.wpq = formats.word_per_query(L)
    lens = torch.zeros(formats.ceil32(n), dtype=torch.int32, device=genome.device)
    lens[:n] = L
    b.lens = lens
    b.queries = packing.pack_queries(rs.reads, lens[:n], b.wpq)
    return b


_PINNED_KEEP = []


def pinned_np(t):
    """device tensor -> numpy uint32 view of a pinned host copy (the tensor is kept alive)"""
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t)
    _PINNED_KEEP.append(h)
    return h.numpy().view(np.uint32)


def stage_parity(args, gi, host, genome, L, wpq, reads, sp, se_mode):
    """a sample at full size through the seeded stage (every read / pair of the sample sent through it) against oracle/seeding_oracle.py"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import seeding_oracle
    from test_stages_gpu import OracleEnv, PAR
    m = min(args.parity_pairs // 16, 2048)
    m = m - (m & 1)
    rd = [np.ascontiguousarray(reads[r]) for r in range(m)]
    lens_m = np.zeros(formats.ceil32(m), np.uint32)
    lens_m[:m] = L
    qm = formats.pack_queries(np.stack(rd), lens_m[:m], wpq)

    class HI:
        pass
    hi = HI()
    hi.bwt, hi.occ, hi.rbwt, hi.rocc = host["bwt"], host["occ"], host["rbwt"], host["rocc"]
    hi.isa0, hi.risa0, hi.n = int(host["meta"][0]), int(host["meta"][1]), int(host["meta"][2])
    env = OracleEnv(None, hi, sa=host["sa"])
    gv = _GenomeView(genome)
    t0 = time.time()
    if se_mode:
        ids = np.arange(m, dtype=np.uint32)
        got = api.single_dp_align(gi, qm, lens_m, m, wpq, ids, sp)
        want = seeding_oracle.single_dp(env, gv, rd, ids.tolist(), PAR)
        same = got["num_seeds"] == want["seeds"] and got["num_candidates"] == want["candidates"] and got["unseeded"].tolist() == want["unseeded"] \
            and len(got["hits"]) == len(want["hits"])
        if same:
            for h, w in zip(got["hits"], want["hits"]):
                cig = api.runs_to_cigar(got["runs"][int(h["runOffset"]):int(h["runOffset"]) + int(h["numRuns"])])
                same &= (int(h["readID"]), int(h["strand"]), int(h["pos"]), int(h["score"]), int(h["numSameScore"]), cig) == w
    else:
        ids = np.arange(0, m, 2, dtype=np.uint32)
        got = api.deep_dp_align(gi, qm, lens_m, m, wpq, ids, sp)
        want = seeding_oracle.deep_dp(env, gv, rd, ids.tolist(), PAR)
        same = got["num_seeds"] == want["seeds"] and got["num_candidates"] == want["candidates"] and got["unseeded"].tolist() == want["unseeded"] \
            and len(got["hits"]) == len(want["hits"])
        if same:
            for h, w in zip(got["hits"], want["hits"]):
                c1 = api.runs_to_cigar(got["runs"][int(h["runOffset1"]):int(h["runOffset1"]) + int(h["numRuns1"])])
                c2 = api.runs_to_cigar(got["runs"][int(h["runOffset2"]):int(h["runOffset2"]) + int(h["numRuns2"])])
                same &= (int(h["readID"]), int(h["strand1"]), int(h["strand2"]), int(h["pos1"]), int(h["pos2"]), int(h["score1"]), int(h["score2"]),
                         int(h["numSame1"]), int(h["numSame2"]), c1, c2) == w
    return {("reads" if se_mode else "pairs"): int(m if se_mode else m // 2), "seeds": int(want["seeds"]),
            "candidates": int(want["candidates"]), "stage_alignments": len(want["hits"]), "stage_bit_exact": bool(same),
            "seconds": time.time() - t0,
            "checker": "oracle/seeding_oracle.py (every read of the sample sent through the seeded stage) over the oracle restatements"}, (m, rd, qm, lens_m, hi)


def run_stage_config(args, gi, host, genome, device, local_rank, rank, world, threads):
    """BASELINE config 3 (se150_dp): the search chain followed by the DP stage that starts from seeds, for the reads the chain left
    unaligned -- 150 bp single-end reads with 1 % substitutions and 0.15 % indels: s3_se_align in long-read mode (first 100 bases searched
    with <= 2 mismatches, occurrences extended over the read: validateAlignments), then s3_single_dp_align (DPForUnalignSingle2) for the
    reads without a valid occurrence.  (The deep-DP stage of config 4 is part of the default step, see main.)
    The seeded stage is a host-orchestrated entry (host arrays in and out), so every number here is end to end: wall clock over
    K steps through the host-pointer C ABI, queries from host memory, results into host memory."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    L = 150
    N = args.pairs
    total = args.warmup + args.steps
    wpq = formats.word_per_query(L)
    sets = []
    for s in range(total):
        rs = synth.simulate_single_end(genome, N, L, seed=500 + 1000 * rank + s, sub_rate=0.01, indel_rate=0.0015)
        lens = torch.zeros(formats.ceil32(N), dtype=torch.int32, device=device)
        lens[:N] = L
        q = packing.pack_queries(rs.reads, lens[:N], wpq)
        sets.append((pinned_np(q), pinned_np(lens), rs.reads.cpu().numpy()))
    sp = api.stage_params(insert_low=INSERT_LO, insert_high=INSERT_HI, scores=DP_SCORES)
    chain = api.SingleAligner(gi, N, num_mismatch=K_MISMATCH, max_output_per_read=MAX_OUTPUT_PER_READ, long_read_mode=True)

    # more handles on the same index, each with its own chain and stage workspace: T batches in flight, one host thread each
    T = max(int(os.environ.get("S3_IN_FLIGHT", "4")), 1)
    handles = [(chain, gi)]
    for _ in range(T - 1):
        g2 = api.index_clone(gi)
        handles.append((api.SingleAligner(g2, N, num_mismatch=K_MISMATCH, max_output_per_read=MAX_OUTPUT_PER_READ, long_read_mode=True), g2))

    def step(q, lens, chain=chain, gi=gi):
        # results stay where the C entries put them (the handle's pinned buffers / malloc'ed arrays): no numpy copies in the timed loop
        t0 = time.perf_counter()
        got = chain.align(q, lens, N, wpq, copy=False)
        t1 = time.perf_counter()
        found = got["occ_offsets"][1:] != got["occ_offsets"][:-1]
        ids = np.nonzero(~found & (got["read_flags"] == 0))[0].astype(np.uint32)
        aligned = int(found.sum())
        t1b = time.perf_counter()
        res = api.single_dp_align(gi, q, lens, N, wpq, ids, sp, counts_only=True)
        t2 = time.perf_counter()
        return got, res, ids, (t1 - t0, t2 - t1b, t1b - t1), aligned

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
    for s in range(args.warmup):
        for ch, g in handles:
            step(sets[s][0], sets[s][1], ch, g)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = api.launch_count()
    t_chain = t_stage = t_pick = 0.0
    n_ids = n_hits = n_seeds = n_cand = n_aligned = n_unseeded = 0
    h2d = d2h = 0
    t0 = time.perf_counter()
    for kk in range(args.steps):
        got, res, ids, (a, b2, c2), aligned = step(sets[args.warmup + kk][0], sets[args.warmup + kk][1])
        t_chain += a; t_stage += b2; t_pick += c2
        n_ids += len(ids); n_hits += res["num_hits"]; n_seeds += res["num_seeds"]; n_cand += res["num_candidates"]; n_aligned += aligned
        n_unseeded += res["num_unseeded"]
        h2d += got["h2d_bytes"]; d2h += got["d2h_bytes"]
    barrier()
    t_one = time.perf_counter() - t0
    launches = api.launch_count() - launches0
    # the same K steps, even ones on the first handle and odd ones on its clone, a host thread each

    def run_half(first, ch, g):
        torch.cuda.set_device(local_rank)
        for kk in range(first, args.steps, T):
            step(sets[args.warmup + kk][0], sets[args.warmup + kk][1], ch, g)
    def warm_half(ch, g):
        torch.cuda.set_device(local_rank)
        for s_ in range(args.warmup):
            step(sets[s_][0], sets[s_][1], ch, g)
    # (warm-up of the threaded shape: with T handles allocating at once the stream-ordered pool reaches its working size here, not in the timed pass)
    wt = [threading.Thread(target=warm_half, args=(ch, g)) for ch, g in handles]
    for t in wt:
        t.start()
    for t in wt:
        t.join()
    th = [threading.Thread(target=run_half, args=(i, ch, g)) for i, (ch, g) in enumerate(handles)]
    barrier()
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    barrier()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    if rank != 0:
        return
    t_total = float(tt[0])
    K = args.steps
    unit = "reads"
    value = world * N * K / t_total
    name = (f"se_{L}bp_indels_genome{args.genome_bp}bp: per step and GPU {N} reads (1 % substitutions, 0.15 % indels) through the long-read search chain "
            "(first 100 bases, k<=2, validateAlignments) and single-read DP from seeds for the rest")
    out = {"metric": "reads/s aligned (SE 150 bp with indels, search + single-read DP, 3.1 Gbp synth ref)",
           "value": value, "unit": "reads/s", "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / K,
           "value_one_batch_in_flight": world * N * K / t_one, "ms_per_step_one_batch_in_flight": 1e3 * t_one / K,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
           "config": {"workload": name, "genome_bp": args.genome_bp, "repeat_fraction": args.repeat_fraction, "reads_per_step_per_gpu": N,
                      "timing": "wall clock over K steps through the host-pointer entries, queries in pinned host memory, every result in host memory "
                                "(the seeded stages' logic runs on the host between device steps): value == e2e; " + str(T) + " batches in flight (step k on handle "
                                "k mod T: the index handle and its s3_index_clones, a host thread each); value_one_batch_in_flight and the stage times: the same steps one after the other",
                      "batches_in_flight": T,
                      "l2": "inputs larger than L2: the 56 GB index is touched at random, a different read batch every step",
                      "parallelism": f"reads sharded over {world} GPU(s), index replicated, no collective"},
           "clocks": sampler.result(),
           "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K), "ms_per_step": 1e3 * t_total / K,
                   "note": "h2d / d2h count the chain's transfers; the seeded stage moves its own seed, candidate and window arrays"},
           "gpu_launches": int(launches),
           "stages_ms_per_step": {"search chain (s3_se_align long-read mode)": 1e3 * t_chain / K,
                                  "picking the reads that go on (host, numpy)": 1e3 * t_pick / K,
                                  "s3_single_dp_align": 1e3 * t_stage / K},
           "pipeline": {f"{unit}_aligned_by_the_chain_per_step": n_aligned / K, f"{unit}_sent_to_the_seeded_stage_per_step": n_ids / K,
                        "seeds_per_step": n_seeds / K, "candidates_per_step": n_cand / K, "stage_alignments_per_step": n_hits / K,
                        f"{unit}_without_a_candidate_per_step": n_unseeded / K},
           "roofline": None}
    if world == 1 and not args.no_cpu_baseline:
        # parity of a sample at full size: the chain and the seeded stage against the compositions of the oracles
        import helpers
        out["parity_at_full_size"], (m, rd, qm, lens_m, hi) = stage_parity(args, gi, host, genome, L, wpq, sets[-1][2], sp, True)
        log("seeded stage parity at full size:", out["parity_at_full_size"])
        # and the long-read chain of the same sample against the reference's CPU search of the seeds + the validation oracle
        ref_c = ref_cpu_handle(host, threads)
        if ref_c is not None:
            seeds = np.ascontiguousarray(np.stack(rd)[:, :100])
            hits = ref_c.search(seeds, K_MISMATCH, formats.NUM_CASES[K_MISMATCH], MAX_OUTPUT_PER_READ, threads, out_cap=64)["hits"]
            al = api.SingleAligner(gi, m, num_mismatch=K_MISMATCH, max_output_per_read=MAX_OUTPUT_PER_READ, long_read_mode=True)
            g2 = al.align(qm, lens_m, m, wpq)
            al.free()
            pac = host["pac"]
            olib = helpers.load_oracle()
            ok = cmp = 0
            for r in range(m):
                if len(hits[r]) >= 64 or int(g2["read_flags"][r]):
                    continue
                a, b2 = int(g2["occ_offsets"][r]), int(g2["occ_offsets"][r + 1])
                mine = sorted((int(p), int(f[0]), int(f[1])) for p, f in zip(g2["positions"][a:b2], g2["occ_flags"][a:b2]))
                hp = [h[0] for h in hits[r]]; hs = [h[1] for h in hits[r]]; hm = [h[2] for h in hits[r]]
                vp, vs, vm = helpers.validate_one(olib.s3o_validate_one, pac, hi.n, rd[r], 100, hp, hs, hm, 0, 0, int(np.ceil(0.02 * L)), MAX_OUTPUT_PER_READ)
                cmp += 1
                ok += mine == sorted(zip(vp, vs, vm))
            out["parity_at_full_size"]["long_read_chain_vs_reference_cpu_search"] = {"reads_compared": cmp, "equal": ok, "bit_exact": bool(ok == cmp)}
            log("long-read chain vs the reference's CPU search + validation oracle:", out["parity_at_full_size"]["long_read_chain_vs_reference_cpu_search"])
            # CPU baseline of this config's search leg: the reference's CPU search of the 100-base seeds
            nb = min(args.cpu_sample // 2, N)
            sd = np.ascontiguousarray(sets[0][2][:nb, :100])
            t0 = time.perf_counter()
            res = ref_c.search(sd, K_MISMATCH, formats.NUM_CASES[K_MISMATCH], MAX_OUTPUT_PER_READ, threads)
            tcs = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nb / tcs, "unit": "reads/s", "cores": threads, "kind": "reference", "cpu_model": cpu_model(),
                                   "sample": f"{nb} reads: the reference's CPU search (ProcessReadDoubleStrand2) of the first 100 bases, k<=2, both strands; "
                                             "the validation and the seeded DP stage are not included (the reference has no CPU DP)",
                                   "reads_with_a_hit": int((res["counts"][:, 3] > 0).sum())}
    print(json.dumps(out), flush=True)
    for ch, g in reversed(handles):                               # clones are freed before the handle they were made from
        ch.free()
        api.GPUINDEXFree(g)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_se(args, gi, host, genome, device, local_rank, rank, world, numa, stream, L, threads, k):
    """BASELINE config 2: single-end 100 bp reads, <= k mismatches (10 cases at k = 4, both strands), search + answer collection +
    locate per step through s3_se_align_device / s3_se_align (alignSingleR's results)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    N = 2 * args.pairs
    total = args.warmup + args.steps
    batches = [make_se_batch(genome, N, L, seed=300 + 1000 * rank + s) for s in range(total)]
    wpq = batches[0].wpq
    se = api.SingleAligner(gi, N, num_mismatch=k, max_output_per_read=1000, report_best=False)
    ncases = formats.NUM_CASES[k]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # more handles on the same index arrays, each with its own chain: T batches in flight, one host thread each (the enumerator of one batch is
    # latency-bound; the other batches' kernels run beside it)
    T = max(int(os.environ.get("S3_IN_FLIGHT", "4")), 1)
    gis, ses, streams = [gi], [se], [stream]
    for _ in range(T - 1):
        g2 = api.index_clone(gi)
        gis.append(g2)
        ses.append(api.SingleAligner(g2, N, num_mismatch=k, max_output_per_read=1000, report_best=False))
        streams.append(torch.cuda.ExternalStream(g2.stream, device=device))

    def in_threads(fns):
        errs = []

        def run(fn):
            try:
                torch.cuda.set_device(local_rank)
                fn()
            except Exception as e:                          # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=run, args=(fn,)) for fn in fns]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if errs:
            raise errs[0]

    def device_step(b, h=0):
        return ses[h].align_device(b.queries.data_ptr(), b.lens.data_ptr(), b.n, b.wpq)
    for s in range(args.warmup):
        for h in range(T):
            device_step(batches[s], h)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    e0.record(stream)
    for kk in range(args.steps):
        stats.append(device_step(batches[args.warmup + kk]))
    e1.record(stream)
    stream.synchronize()
    barrier()
    tt = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    t_one = float(tt[0])
    # the same K steps with T batches in flight: step k on handle k mod T
    timed_batches = [batches[args.warmup + kk] for kk in range(args.steps)]
    in_threads([(lambda h=h: [device_step(batches[s_], h) for s_ in range(args.warmup)]) for h in range(T)])
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = api.launch_count()
    d0.record(stream)
    in_threads([(lambda h=h: [device_step(b_, h) for b_ in timed_batches[h::T]]) for h in range(T)])
    for h in range(1, T):
        fin = torch.cuda.Event()
        fin.record(streams[h])
        stream.wait_event(fin)
    d1.record(stream)
    stream.synchronize()
    barrier()
    launches = api.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join()
    tt = torch.tensor([d0.elapsed_time(d1) / 1e3], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    t_total = float(tt[0])
    value = world * N * args.steps / t_total
    api.set_timing(gi.handle, True)
    device_step(batches[0])
    barrier()
    api.read_timing(gi.handle)
    for kk in range(args.steps):
        device_step(batches[args.warmup + kk])
    barrier()
    ms_search, n_search = api.read_timing(gi.handle)
    api.set_timing(gi.handle, False)

    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h
    host_sets = [(pinned(batches[args.warmup + kk].queries), pinned(batches[args.warmup + kk].lens)) for kk in range(args.steps)]
    lasts = [None] * T

    def e2e_on(h, sets):
        for q, l in sets:
            lasts[h] = ses[h].align(q.data_ptr(), l.data_ptr(), N, wpq, copy=False)
    in_threads([(lambda h=h: e2e_on(h, host_sets[:2])) for h in range(T)])
    barrier()
    t0 = time.perf_counter()
    in_threads([(lambda h=h: e2e_on(h, host_sets[h::T])) for h in range(T)])
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    last = lasts[0]
    if world > 1:
        torch.distributed.all_reduce(te, op=torch.distributed.ReduceOp.MAX)
    t_e2e = float(te[0])
    if rank != 0:
        return
    K = args.steps
    names_s = ["s3_search_easy_kernel", "s3_search_kernel<items>", "s3_search_kernel<spine>", "s3_search_kernel<tasks>",
               "s3_heavy_merge_kernel", "s3_isbad_fixup_kernel"]
    kernels = {nm: {"ms_per_step": ms / K, "launches_per_step": cnt / K} for nm, ms, cnt in zip(names_s, ms_search, n_search) if cnt}
    t_search = sum(ms_search) / 1e3
    try:
        sector_rate = random_sector_rate(gi, device)
    except Exception:                                        # noqa: BLE001
        sector_rate = None
    traffic = {}
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path))
    tkey = f"search_launch_k{k}"
    per_load = float(traffic.get("random_sector_probe_dram_bytes_per_load") or 64.0)
    roof = {"kernel": f"search launch, k <= {k} ({ncases} cases)", "bound": "hbm", "unit": "GB/s", "ms_per_launch": 1e3 * t_search / K,
            "share_of_step": (t_search / K) / (t_total / K), "peak": sector_rate * per_load / 1e9 if sector_rate else None,
            "peak_source": "s3_random_sector_probe in this run (independent random 32-byte loads over the index's bucket array) x "
                           f"{per_load:.0f} DRAM bytes per load (ncu of the probe kernel, profiles/ncu_traffic.json)",
            "random_loads_per_s": sector_rate, "traffic": traffic.get(tkey), "achieved": None, "frac": None}
    if traffic.get(tkey) and sector_rate:
        roof["achieved"] = traffic[tkey] / (t_search / K) / 1e9
        roof["frac"] = roof["achieved"] / roof["peak"]
    out = {"metric": f"reads/s searched and located (SE {L} bp, <= {k} mismatches, 3.1 Gbp synth ref)", "value": value, "unit": "reads/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u32", "data": "synthetic",
           "value_one_batch_in_flight": world * N * args.steps / t_one, "ms_per_step_one_batch_in_flight": 1e3 * t_one / args.steps,
           "config": {"workload": f"se_{L}bp_k{k}_genome{args.genome_bp}bp: per step and GPU {N} reads through <= {k}-mismatch search ({ncases} cases, both "
                                  "strands, round-1 slots), answer collection and locate (s3_se_align)",
                      "genome_bp": args.genome_bp, "repeat_fraction": args.repeat_fraction, "reads_per_step_per_gpu": N,
                      "batches_in_flight": T,
                      "timing": f"value: K steps, queries resident in HBM (one 8-byte count read per step is part of the chain), {T} batches in flight (step k on "
                                "handle k mod T: the index handle and its s3_index_clones, a host thread each), CUDA events; value_one_batch_in_flight: the same "
                                "steps back to back on one handle; e2e: s3_se_align with pinned host queries in, occurrences out, wall clock, the same threads",
                      "l2": "inputs larger than L2: the 56 GB index is touched at random, a different read batch every step",
                      "parallelism": f"reads sharded over {world} GPU(s), index replicated, no collective"},
           "clocks": sampler.result(),
           "e2e": {"value": world * N * args.steps / t_e2e, "unit": "reads/s", "h2d_bytes_per_step": last["h2d_bytes"], "d2h_bytes_per_step": last["d2h_bytes"],
                   "ms_per_step": 1e3 * t_e2e / args.steps},
           "gpu_launches": int(launches), "roofline": roof, "kernels": kernels,
           "pipeline": {"ranges_per_step": float(np.mean([int(r.numRanges) for r in stats])),
                        "occurrences_per_step": float(np.mean([int(r.numOccurrences) for r in stats])),
                        "reads_with_a_slot_overflow_per_step": None}}
    if world == 1 and not args.no_cpu_baseline:
        import helpers
        import pe_chain_oracle
        ref_s = helpers.load_ref_search()
        n = min(args.cpu_sample // 16, N)
        b = make_se_batch(genome, n, L, seed=4343)
        q = b.queries.cpu().numpy().view(np.uint32)
        lens = b.lens.cpu().numpy().view(np.uint32)

        class HI:
            pass
        hi = HI()
        hi.bwt, hi.occ, hi.rbwt, hi.rocc = host["bwt"], host["occ"], host["rbwt"], host["rocc"]
        hi.isa0, hi.risa0, hi.n = int(host["meta"][0]), int(host["meta"][1]), int(host["meta"][2])
        allowed = formats.SA_RANGES_ROUND1[k]
        wpa = 2 * allowed
        bad = np.zeros(formats.ceil32(n), np.uint8)
        views = []
        qq = q.copy()
        t0 = time.perf_counter()
        nrank = 0
        for case in range(ncases):
            a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
            if ref_s is not None:
                nrank += helpers.ref_launch(ref_s, hi, case, qq, lens, n, wpq, a, bad, 0, k, allowed, wpa, nthreads=threads)
            else:
                nrank += helpers.oracle_launch(helpers.load_oracle(), hi, case, q, lens, n, wpq, a, bad, 0, k, allowed, wpa)
            views.append(formats.answers_view(a, n, wpa))
        t_cpu = time.perf_counter() - t0
        kernel_code = {"what": "the reference's GPU search kernels (DV-Kernel.cu) compiled for the host, OpenMP over reads", "reads": int(n),
                       "seconds": t_cpu, "search_reads_per_s": n / t_cpu, "rank_queries_per_read": nrank / n}
        ref_c = ref_cpu_handle(host, threads)
        if ref_c is not None:
            # the reference's own CPU search on a sample of its own size (it is much faster per read than the kernel code on CPU cores)
            nb = min(args.cpu_sample, 4 * N)
            bb = make_se_batch(genome, nb, L, seed=4344)
            rd = np.ascontiguousarray(bb.reads.cpu().numpy().astype(np.uint8))
            t0 = time.perf_counter()
            res = ref_c.search(rd, k, ncases, 1000, threads)
            t_ref = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": nb / t_ref, "unit": "reads/s", "cores": threads, "kind": "reference",
                                   "sample": f"{nb} reads of the bench workload through the reference's CPU search (ProcessReadDoubleStrand2 per case on "
                                             f"SRAModelConstruct's 16G models, <= {k} mismatches, {ncases} cases, both strands, lookup tables + "
                                             "check-and-extend, MaxOutputPerRead 1000, OpenMP over reads)",
                                   "cpu_model": cpu_model(), "search_path": "ProcessReadDoubleStrand2",
                                   "cpu_search": {"reads": int(nb), "seconds": t_ref, "occurrences": int(res["counts"][:, 3].sum()),
                                                  "reads_with_a_hit": int((res["counts"][:, 3] > 0).sum())},
                                   "kernel_code": kernel_code, "rank_queries_per_read": nrank / n}
            del bb, rd, res
        else:
            out["cpu_baseline"] = {"value": n / t_cpu, "unit": "reads/s", "cores": threads if ref_s is not None else 1, "kind": "reference" if ref_s is not None else "port",
                                   "sample": f"{n} reads, <= {k} mismatches, {ncases} cases, both strands: reference kernel sources (DV-Kernel.cu) compiled for the host, OpenMP over reads",
                                   "cpu_model": cpu_model(), "rank_queries_per_read": nrank / n, "search_path": "kernel_code", "kernel_code": kernel_code}
        # parity of the sample through the chain
        al = api.SingleAligner(gi, n, num_mismatch=k, max_output_per_read=1000, report_best=False)
        got = al.align(q, lens, n, wpq)
        al.free()
        col = pe_chain_oracle.collect(views, allowed, hi.n, 1000)
        sa = host["sa"]
        ok, off, located = True, 0, 0
        for r, (ranges, tot, more) in enumerate(col):
            a0, a1 = int(got["occ_offsets"][r]), int(got["occ_offsets"][r + 1])
            want = np.concatenate([sa[l:rr + 1] for l, rr, _, _ in ranges]).astype(np.uint32) if ranges else np.zeros(0, np.uint32)
            wf = np.array([[st, mm] for l, rr, st, mm in ranges for _ in range(rr - l + 1)], np.uint8).reshape(-1, 2)
            ok &= a0 == off and a1 - a0 == len(want) and np.array_equal(got["positions"][a0:a1], want) and np.array_equal(got["occ_flags"][a0:a1], wf)
            ok &= int(got["read_flags"][r]) == int(more)
            off = a1
            located += len(want)
        if not ok:
            log("SE CHAIN PARITY FAILURE at full size")
        true_pos = b.pos.cpu().numpy()
        first = got["occ_offsets"][:-1]
        has = np.diff(got["occ_offsets"].astype(np.int64)) > 0
        at_truth = int((np.abs(got["positions"][first[has]].astype(np.int64) - true_pos[has]) <= 0).sum()) if has.any() else 0
        out["parity_at_full_size"] = {"reads": int(n), "cases": ncases, "occurrences": int(located), "occurrences_bit_exact": bool(ok),
                                      "reads_with_hits": int(has.sum()), "reads_whose_first_hit_is_the_simulated_position": at_truth,
                                      "reads_with_a_slot_overflow": int(got["read_flags"].sum()),
                                      "checker": "oracle/pe_chain_oracle.collect over " + ("the reference's kernels compiled for the host" if ref_s is not None else "the oracle port")}
        out["pipeline"]["reads_with_a_slot_overflow_per_step"] = float(got["read_flags"].mean() * N)
    print(json.dumps(out), flush=True)
    for h in reversed(range(T)):                                  # clones are freed before the handle they were made from
        ses[h].free()
        api.GPUINDEXFree(gis[h])
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-bp", type=int, default=int(os.environ.get("S3_GENOME_BP", 3_100_000_000)))
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("S3_PAIRS", 524_288)), help="pairs per step per GPU")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("S3_CPU_SAMPLE", 2_097_152)))
    ap.add_argument("--parity-pairs", type=int, default=int(os.environ.get("S3_PARITY_PAIRS", 32_768)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--repeat-fraction", type=float, default=float(os.environ.get("S3_REPEAT_FRACTION", 0.2)))
    ap.add_argument("--config", default="pe100", choices=["pe100", "pe100_deep", "pe100_chain", "se100_k4", "se150_dp"],
                    help="pe100 (default; pe100_deep is the same): the BASELINE metric's workload, config 4 -- search + mate-rescue DP + deep DP of the "
                         "both-unaligned pairs; pe100_chain: the same without the deep-DP stage; se100_k4: config 2, single-end 100 bp, <= 4 mismatches, "
                         "search + collect + locate (s3_se_align); se150_dp: config 3, 150 bp reads with indels, long-read search chain + single-read DP from "
                         "seeds.  A 45 %%-repeat genome: --repeat-fraction 0.45")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: the timing rules ask for >= 3 warm-up steps")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if api.load_library().s3_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; soap3dp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if args.impl == "reference" and rank != 0:
        return
    if world > 1 and args.impl == "ours":
        # NCCL announces its version on stdout when the first communicator is made: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.distributed.init_process_group("nccl", device_id=device)
            torch.distributed.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L = args.read_len
    genome, host = get_index(args.genome_bp, 3, device, rank, world if args.impl == "ours" else 1, args.repeat_fraction)
    threads = os.cpu_count() or 1
    workload = (f"pe_2x{L}bp_insert{INSERT_LO}-{INSERT_HI}_genome{args.genome_bp}bp: per step and GPU {args.pairs} read pairs through k<=2 search "
                f"(4 cases, both strands, round-1 slots), answer collection, routing, locate, pairing and mate-rescue DP (400 bp windows) with CIGARs"
                + ("" if args.config == "pe100_chain" else "; then deep DP (DPForUnalignPairs2) of the pairs with no occurrence of either read: seeds of both mates, "
                   "seeding driver, candidate position pairs, left window + DP, right window + DP, CIGARs"))

    if args.impl == "reference":
        vals = []
        info = None
        per_step = min(2 * args.pairs, max(args.cpu_sample // 2, 4096))        # the GPU arm's reads per step (1,048,576), about 4 s of CPU work
        for s in range(args.warmup + args.steps):
            info = cpu_arm(host, genome, L, per_step, 1000 + s, RESCUE_FRACTION_HINT, threads, kernel_code_reads=32768, gpu_legs=False)
            if s >= args.warmup:
                vals.append(info)
        v = float(np.mean([x["value"] for x in vals]))
        out = {"impl": "reference", "metric": "reads/s aligned (2x100bp PE, 3.1 Gbp synth ref)", "value": v,
               "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * per_step / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u32", "data": "synthetic",
               "config": {"workload": workload, "genome_bp": args.genome_bp, "repeat_fraction": args.repeat_fraction, "sample_reads_per_step": per_step,
                          "sample": "the reference arm runs the same stages on a bounded sample per step: k<=2 search of "
                                    f"{per_step} reads + mate-rescue DP of the same fraction of pairs as the GPU arm rescues"},
               "cpu_baseline": {"value": v, "unit": "reads/s", "cores": info["cores"], "kind": info["kind"],
                                "sample": info["sample"], "cpu_model": cpu_model(), "search_path": info["search_path"],
                                "cpu_search": info.get("cpu_search"), "kernel_code": info["kernel_code"], "dp_gcups": info["dp_gcups"]},
               "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out), flush=True)
        return

    t0 = time.time()
    gi = upload_index(host, local_rank)
    if not (rank == 0 and world == 1):
        host.pop("sa", None)                                  # 12 GB of host memory per rank, only the parity check reads it
    if os.environ.get("S3_SPLIT_BUDGET"):                     # tuning experiments (profiles/variants.sh)
        api.set_split_budget(gi, int(os.environ["S3_SPLIT_BUDGET"]))
    log(f"index on device in {time.time() - t0:.1f}s: {gi.device_bytes / 1e9:.2f} GB (32-byte single-sector buckets, seed tables, "
        f"suffix array + inverse + packed text)")
    stream = torch.cuda.ExternalStream(gi.stream, device=device)
    total = args.warmup + args.steps
    if os.environ.get("S3_PROBE_ONLY"):                       # ncu of the probe itself: DRAM bytes per random 32-byte load (profiles/)
        print(json.dumps({"random_sector_loads_per_s": random_sector_rate(gi, device)}), flush=True)
        api.GPUINDEXFree(gi)
        return
    if os.environ.get("S3_L2_REGION"):                        # ncu captures with a window in place (profiles/exp_l2_persist.sh)
        api.set_l2_persist(gi, int(os.environ["S3_L2_REGION"]))
    if args.config == "se150_dp":
        return run_stage_config(args, gi, host, genome, device, local_rank, rank, world, threads)
    if args.config == "se100_k4":
        return run_se(args, gi, host, genome, device, local_rank, rank, world, numa, stream, L, threads, k=4)
    t0 = time.time()
    batches = [make_batch(genome, args.pairs, L, seed=100 + 1000 * rank + s) for s in range(total)]
    N, wpq = batches[0].n, batches[0].wpq
    par = api.pe_params(num_mismatch=K_MISMATCH, insert_low=INSERT_LO, insert_high=INSERT_HI, scores=DP_SCORES, read_length=L,
                        max_windows=N // 2)
    pe = api.PairAligner(gi, N, L, par)
    # more handles on the same index arrays, each with its own chain and stage workspace: T batches in flight, one host thread each
    with_deep = args.config != "pe100_chain"
    T = max(int(os.environ.get("S3_IN_FLIGHT", "4")), 1)
    sp = api.stage_params(insert_low=INSERT_LO, insert_high=INSERT_HI, scores=DP_SCORES)
    gis, pes, streams = [gi], [pe], [stream]
    for _ in range(T - 1):
        g2 = api.index_clone(gi)
        gis.append(g2)
        pes.append(api.PairAligner(g2, N, L, par))
        streams.append(torch.cuda.ExternalStream(g2.stream, device=device))
    torch.cuda.synchronize()
    log(f"{total} batches of {N} reads prepared in {time.time() - t0:.1f}s")

    def in_threads(fns):
        errs = []

        def run(fn):
            try:
                torch.cuda.set_device(local_rank)
                fn()
            except Exception as e:                          # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=run, args=(fn,)) for fn in fns]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if errs:
            raise errs[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    deep_stats = []

    def device_step(b, h=0, keep=None):
        # the chain on queries resident in HBM, then (config 4) deep DP of the pairs it left with no occurrence of either read,
        # picked on the device from the chain's routes; results stay where the entries put them
        r = pes[h].align_device(b.queries.data_ptr(), b.lens.data_ptr(), b.n, b.wpq)
        if with_deep:
            d = pes[h].deep_dp(sp, counts_only=True)
            if keep is not None:
                keep.append(d)
        return r

    # ---- device-resident timing: queries in HBM, results left there ---------------------------------------------
    for s in range(args.warmup):
        for h in range(T):
            device_step(batches[s], h)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats = []
    e0.record(stream)
    for k in range(args.steps):
        stats.append(device_step(batches[args.warmup + k], 0, deep_stats))
    e1.record(stream)
    stream.synchronize()
    barrier()
    tt = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    t_one = float(tt[0])
    reads_per_rank = N * args.steps
    value_one = world * reads_per_rank / t_one
    # ---- the same K steps with T batches in flight: step k on handle k mod T ---------------------------------------
    timed_batches = [batches[args.warmup + k] for k in range(args.steps)]
    # (warm-up of the threaded shape: with T handles allocating at once the stream-ordered pool reaches its working size here, not in the timed pass)
    in_threads([(lambda h=h: [device_step(batches[s_], h) for s_ in range(args.warmup)]) for h in range(T)])
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = api.launch_count()
    d0.record(stream)
    in_threads([(lambda h=h: [device_step(b, h) for b in timed_batches[h::T]]) for h in range(T)])
    for h in range(1, T):
        fin = torch.cuda.Event()
        fin.record(streams[h])
        stream.wait_event(fin)
    d1.record(stream)
    stream.synchronize()
    barrier()
    launches = api.launch_count() - launches0
    tt = torch.tensor([d0.elapsed_time(d1) / 1e3], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    t_total = float(tt[0])
    value = world * reads_per_rank / t_total
    windows = [int(r.numWindows) for r in stats]
    wlen = INSERT_HI - INSERT_LO + L
    dp_cells = float(sum(windows)) * L * wlen
    routes = np.sum([list(r.routeCounts)[:9] for r in stats], axis=0)

    sampler.stop_flag = True
    sampler.join()

    # ---- the same steps with the per-kernel timing hooks on (events between the library's launches) ---------------
    api.set_timing(gi.handle, True)
    api.set_timing(pe.dp_handle, True, dp=True)
    pe.set_timing(True)
    chain_step = lambda b: pe.align_device(b.queries.data_ptr(), b.lens.data_ptr(), b.n, b.wpq)
    chain_step(batches[0])
    barrier()
    api.read_timing(gi.handle)
    api.read_timing(pe.dp_handle, dp=True)
    pe.read_timing()
    st0 = pe.read_timing()
    for k in range(args.steps):
        chain_step(batches[args.warmup + k])
    barrier()
    ms_search, n_search = api.read_timing(gi.handle)
    ms_dp, n_dp = api.read_timing(pe.dp_handle, dp=True)
    st1 = pe.read_timing()
    ms_stage = [b_ - a_ for a_, b_ in zip(st0, st1)]
    api.set_timing(gi.handle, False)
    api.set_timing(pe.dp_handle, False, dp=True)
    pe.set_timing(False)

    # ---- L2 access-policy windows over the index arrays (north_star choice 2), measured: the search launch with each ----
    l2_exp = None
    if os.environ.get("S3_L2_EXPERIMENT") and rank == 0:
        l2_exp = {}
        names = {0: "none", 1: "forward buckets", 2: "reverse buckets", 3: "seed table fwd1", 4: "seed table rev0", 5: "packed text", 6: "suffix array"}
        allowed, wpa, ncases = formats.SA_RANGES_ROUND1[K_MISMATCH], 2 * formats.SA_RANGES_ROUND1[K_MISMATCH], formats.NUM_CASES[K_MISMATCH]
        ans = [torch.empty(formats.ceil32(N) * wpa, dtype=torch.int32, device=device) for _ in range(ncases)]
        ptrs = [a_.data_ptr() for a_ in ans]
        for region in (0, 1, 2, 3, 4, 5, 6, 0):
            api.set_l2_persist(gi, region)
            api.set_timing(gi.handle, True)
            b0 = batches[0]
            api.search_round1_device(gi, b0.queries.data_ptr(), b0.lens.data_ptr(), b0.n, b0.wpq, K_MISMATCH, ncases, allowed, wpa, ptrs, 0)
            torch.cuda.synchronize()
            api.read_timing(gi.handle)
            for kk in range(args.steps):
                bb = batches[args.warmup + kk]
                api.search_round1_device(gi, bb.queries.data_ptr(), bb.lens.data_ptr(), bb.n, bb.wpq, K_MISMATCH, ncases, allowed, wpa, ptrs, 0)
            torch.cuda.synchronize()
            ms, cnt = api.read_timing(gi.handle)
            api.set_timing(gi.handle, False)
            key = names[region] + (" (again)" if region == 0 and "none" in l2_exp else "")
            l2_exp[key] = {"easy_kernel_ms": ms[0] / args.steps, "enumerator_ms": ms[1] / args.steps, "search_launch_ms": sum(ms) / args.steps}
        api.set_l2_persist(gi, 0)
        del ans

    # ---- end to end through the host-pointer entry: queries from pinned host memory, results into host memory -----
    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h
    host_sets = [(pinned(batches[args.warmup + k].queries), pinned(batches[args.warmup + k].lens)) for k in range(args.steps)]
    pageable_sets = [(q.numpy().copy(), l.numpy().copy()) for q, l in host_sets[:max(4, T)]]

    def timed(fn):
        barrier()
        t0 = time.perf_counter()
        r = fn()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t[0]), r

    def e2e_steps_on(h, sets, keep=None):
        # the next batch's queries go up (s3_pe_prefetch, a copy stream) while this batch is aligned; the deep stage works on the
        # batch's device copy, its records land in host memory
        out = None
        ptr = lambda x: x.data_ptr() if hasattr(x, "data_ptr") else x
        for k, (q, l) in enumerate(sets):
            if k + 1 < len(sets):
                pes[h].prefetch(ptr(sets[k + 1][0]), ptr(sets[k + 1][1]), N, wpq)
            out = pes[h].align(ptr(q), ptr(l), N, wpq, copy=False)
            if with_deep:
                d = pes[h].deep_dp(sp, counts_only=True)
                if keep is not None:
                    keep.append(d)
        return out
    for s in range(max(min(args.warmup, len(host_sets)) // 2, 1)):          # warm-up with the prefetch path: both input buffers of every handle get allocated
        for h in range(T):
            e2e_steps_on(h, host_sets[:2])
    deep_e2e = []
    t_e2e_one, last = timed(lambda: e2e_steps_on(0, host_sets, deep_e2e))
    # T caller threads, a handle each: the reference's own shape (its main thread searches batch k + 1 while a DP engine thread aligns batch k)
    t_e2e, _ = timed(lambda: in_threads([(lambda h=h: e2e_steps_on(h, host_sets[h::T])) for h in range(T)]))
    h2d, d2h = last["h2d_bytes"], last["d2h_bytes"]
    if deep_e2e:
        # the deep stage's records (56 bytes per paired alignment, runs, the pairs without a candidate); its internal count / id reads are not counted
        d2h += int(np.mean([56 * d["num_hits"] + 4 * d["num_runs"] + 4 * d["num_unseeded"] for d in deep_e2e]))
    e2e_value = world * reads_per_rank / t_e2e
    e2e_steps_on(0, pageable_sets[:1])
    t_page_one, _ = timed(lambda: e2e_steps_on(0, pageable_sets))
    t_page, _ = timed(lambda: in_threads([(lambda h=h: e2e_steps_on(h, pageable_sets[h::T])) for h in range(T)]))
    pageable_value = world * N * len(pageable_sets) / t_page
    # what the link gives: one 256 MiB pinned copy each way, alone
    probe_h = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
    probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    link = {}
    for name, (dst, src) in (("h2d_gbs", (probe_d, probe_h)), ("d2h_gbs", (probe_h, probe_d))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link[name] = (256 << 20) / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del probe_h, probe_d
    if rank != 0:
        return

    # ---- rooflines ---------------------------------------------------------------------------------------------
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = {}
    tr_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes per launch from the last ncu --set full capture
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path))
    K = args.steps
    names_s = ["s3_search_easy_kernel", "s3_search_kernel<items>", "s3_search_kernel<spine>", "s3_search_kernel<tasks>",
               "s3_heavy_merge_kernel", "s3_isbad_fixup_kernel"]
    names_d = ["s3_dp_sweep16_kernel", "s3_dp_resweep16_kernel", "s3_dp_traceback16_kernel", "s3_dp pass 2 (second sweep + traceback of what left its window)"]
    kernels = {}
    for nm, ms, cnt in list(zip(names_s, ms_search, n_search)) + list(zip(names_d, ms_dp, n_dp)):
        if cnt:
            kernels[nm] = {"ms_per_step": ms / K, "launches_per_step": cnt / K}
    stage_names = ["search launch", "collect + route (+ first count read)", "locate + sort", "pairing + window count (+ second count read)",
                   "window descriptors", "DP (pack + sweep + second sweep + traceback)", "CIGAR runs + records"]
    stages = {nm: ms / K for nm, ms in zip(stage_names, ms_stage)}
    t_step_hooks = sum(ms_stage) / K
    t_search = sum(ms_search) / 1e3
    t_dp = sum(ms_dp) / 1e3
    t_sweep = ms_dp[0] / 1e3
    dpx = {}
    dpx_path = os.path.join(ROOT, "profiles", "r01_dpx_rate.json")              # tools/dpx_rate.cu on this pool's B200
    if os.path.exists(dpx_path):
        dpx = json.load(open(dpx_path))
    dpx_peak = float(dpx.get("gcups_peak_5op", 7324.6))
    dp_gcups = dp_cells / t_dp / 1e9 if t_dp > 0 else 0.0
    sweep_gcups = dp_cells / t_sweep / 1e9 if t_sweep > 0 else 0.0
    try:
        sector_rate = random_sector_rate(gi, device)
    except Exception as e:                                   # noqa: BLE001
        log("random sector probe failed:", e)
        sector_rate = None
    # the search launch against the random-access rate of this run, both in DRAM bytes: achieved = DRAM bytes the launch's kernels
    # moved (ncu dram__bytes_read + write, profiles/ncu_traffic.json) / launch time; peak = the probe's independent random 32-byte
    # loads per second x the DRAM bytes ncu counts per such load (HBM3e moves a 64-byte burst per random 32-byte sector:
    # profiles/ncu_traffic.json "random_sector_probe_dram_bytes_per_load")
    per_load = float(traffic.get("random_sector_probe_dram_bytes_per_load") or 64.0)
    search_roof = {"kernel": "search launch (s3_search_easy_kernel + s3_search_kernel<items|spine|tasks> + merge)", "bound": "hbm",
                   "unit": "GB/s", "ms_per_launch": 1e3 * t_search / K, "share_of_step": (t_search / K) / (t_step_hooks / 1e3) if t_step_hooks else None,
                   "peak": sector_rate * per_load / 1e9 if sector_rate else None,
                   "peak_source": "s3_random_sector_probe in this run (independent random 32-byte loads over the index's bucket array) x "
                                  f"{per_load:.0f} DRAM bytes per load (ncu of the probe kernel, profiles/ncu_traffic.json)",
                   "random_loads_per_s": sector_rate, "traffic": traffic.get("search_launch")}
    if traffic.get("search_launch") and sector_rate:
        search_roof["achieved"] = traffic["search_launch"] / (t_search / K) / 1e9
        search_roof["frac"] = search_roof["achieved"] / search_roof["peak"]
        search_roof["frac_of_hbm_copy_peak"] = search_roof["achieved"] / hbm_peak
    dp_traffic = None
    if all(traffic.get(k2) is not None for k2 in ("s3_dp_sweep16_kernel", "s3_dp_resweep16_kernel", "s3_dp_traceback16_kernel")):
        dp_traffic = sum(traffic[k2] for k2 in ("s3_dp_sweep16_kernel", "s3_dp_resweep16_kernel", "s3_dp_traceback16_kernel"))
    sweep_roof = {"kernel": "s3_dp_sweep16_kernel", "bound": "dpx", "achieved": sweep_gcups, "peak": dpx_peak, "unit": "GCUPS",
                  "frac": sweep_gcups / dpx_peak, "traffic": traffic.get("s3_dp_sweep16_kernel"),
                  "peak_source": "tools/dpx_rate.cu on B200 (profiles/r01_dpx_rate.json): VIMNMX3/VIADDMNMX .16x2 issue rate x SMs x clock x "
                                 "64 cells / 5 instructions per cell pair (SURVEY.md 8d)",
                  "cells_per_launch": dp_cells / K, "ms_per_launch": 1e3 * t_sweep / K,
                  "share_of_step": (t_sweep / K) / (t_step_hooks / 1e3) if t_step_hooks else None,
                  "hbm": {"achieved_gbs": (traffic["s3_dp_sweep16_kernel"] / (t_sweep / K) / 1e9) if traffic.get("s3_dp_sweep16_kernel") else None,
                          "peak_gbs": hbm_peak, "peak_source": peak_src}}
    dominant = max(kernels.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if kernels else "s3_dp_sweep16_kernel"
    out = {
        "metric": "reads/s aligned (2x100bp PE, 3.1 Gbp synth ref)",
        "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "value_one_batch_in_flight": value_one, "ms_per_step_one_batch_in_flight": 1e3 * t_one / args.steps,
        "config": {"workload": workload, "genome_bp": args.genome_bp, "repeat_fraction": args.repeat_fraction, "pairs_per_step_per_gpu": args.pairs,
                   "step": "s3_pe_align_device: search -> collect -> route -> locate -> pairing -> rescue windows -> DP -> CIGAR runs, nothing "
                           "taken from the simulator's truth; reads whose round-1 slot overflowed are reported (route 8), not searched again"
                           + ("; then s3_pe_deep_dp on the same handle: the S3_PE_NONE pairs picked on the device, the stage's records into host memory "
                              "(its logic -- rounds, which pairs go on -- runs on the host between device steps, like the reference's wrapper)" if with_deep else ""),
                   "batches_in_flight": T,
                   "timing": f"value: K steps, queries resident in HBM, the chain's results left there, {T} batches in flight (step k on handle k mod {T}: the "
                             "index handle and its s3_index_clones, one host thread each), CUDA events around the whole; value_one_batch_in_flight: the "
                             "same K steps back to back on one handle (two 4-byte count reads per step are part of the chain); kernels / "
                             "stages: the chain of those K steps once more with the library's timing hooks on; e2e: the K steps through s3_pe_align"
                             + (" + s3_pe_deep_dp" if with_deep else "") + ", queries from pinned host memory, results into host memory, wall clock, the same caller threads as value",
                   "l2": "inputs larger than L2: 56 GB of index (buckets, seed tables, suffix array, text) touched at random, a different "
                         "32 MiB read batch every step",
                   "parallelism": f"reads sharded over {world} GPU(s), index replicated, no collective"
                                  + (f"; rank 0 bound to the CPUs of NUMA node {numa}" if numa is not None else "")},
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "one_thread_value": world * reads_per_rank / t_e2e_one, "one_thread_ms_per_step": 1e3 * t_e2e_one / args.steps,
                "mode": "s3_pe_align (host-pointer C ABI): queries from pinned host memory in, routes + pairings + rescue records + CIGAR runs "
                        "into host memory out; the next step's queries are uploaded by s3_pe_prefetch under this step's kernels"
                        + ("; then s3_pe_deep_dp: paired deep-DP alignments + CIGAR runs into host memory" if with_deep else ""),
                "pageable_value": pageable_value, "pageable_ms_per_step": 1e3 * t_page / len(pageable_sets),
                "pageable_one_thread_value": world * N * len(pageable_sets) / t_page_one,
                "link": link},
        "gpu_launches": int(launches),
        "roofline": sweep_roof if dominant.startswith("s3_dp") else search_roof,
        "search": search_roof,
        "dp": {"kernel": "s3_dp_sweep16_kernel + s3_dp_resweep16_kernel + s3_dp_traceback16_kernel", "gcups": dp_gcups,
               "dpx_peak_gcups": dpx_peak, "frac_of_dpx_peak": dp_gcups / dpx_peak, "sweep": sweep_roof,
               "alignments_per_step": float(np.mean(windows)), "cells_per_step": dp_cells / args.steps, "ms_per_step": 1e3 * t_dp / args.steps,
               "dram_bytes_per_step": dp_traffic, "algorithmic_bytes_per_step": 0.5 * dp_cells / args.steps,
               "note": "algorithmic bytes = 0.5 B per cell (SURVEY.md 8d: inputs + traceback bits)"},
        "routes_per_step": {nm: float(routes[c]) / args.steps for c, nm in enumerate(
            ["both unaligned (deep DP list)", "both hit (paired or rescued below)", "first read rescues its mate", "second read rescues its mate",
             "-", "first read: too many hits", "second read: too many hits", "-", "round-1 slot overflow"]) if nm != "-"},
        "pipeline": {"ranges_per_step": float(np.mean([int(r.numRanges) for r in stats])),
                     "occurrences_per_step": float(np.mean([int(r.numOccurrences) for r in stats])),
                     "rescue_windows_per_step": float(np.mean(windows)), "cigar_runs_per_step": float(np.mean([int(r.numRuns) for r in stats]))},
        "stages_ms_per_step": stages,
        "kernels": kernels,
    }
    if with_deep and deep_stats:
        out["deep_dp"] = {"entry": "s3_pe_deep_dp (DPForUnalignPairs2, DV-DPForBothUnalign.cu:245): per step",
                          "pairs": float(np.mean([d["num_input"] for d in deep_stats])), "seeds": float(np.mean([d["num_seeds"] for d in deep_stats])),
                          "candidate_pairs": float(np.mean([d["num_candidates"] for d in deep_stats])),
                          "paired_alignments": float(np.mean([d["num_hits"] for d in deep_stats])),
                          "pairs_without_a_candidate": float(np.mean([d["num_unseeded"] for d in deep_stats])),
                          "ms_per_step_one_batch_in_flight": 1e3 * t_one / args.steps - t_step_hooks,
                          "note": "ms = the one-in-flight step minus the chain's stages timed by the hooks (wall time of the stage incl. its host side)"}
    if l2_exp is not None:
        out["l2_persistence_experiment"] = l2_exp
    if world == 1 and not args.no_cpu_baseline:
        frac = float(np.mean(windows)) / args.pairs
        t0 = time.time()
        cb = cpu_arm(host, genome, L, args.cpu_sample, 999, frac, threads, kernel_code_reads=args.cpu_sample // 4, keep_hits=65536)
        log(f"cpu baseline ({cb['kind']}, {cb['cores']} threads) took {time.time() - t0:.1f}s: {cb['value']:.0f} reads/s")
        out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        out["cpu_baseline"]["cpu_model"] = cpu_model()
        out["cpu_baseline"]["rank_queries_per_read"] = cb["rank_queries_per_read"]
        out["cpu_baseline"]["dp_gcups"] = cb["dp_gcups"]
        out["cpu_baseline"]["search_path"] = cb["search_path"]
        out["cpu_baseline"]["cpu_search"] = cb.get("cpu_search")
        out["cpu_baseline"]["kernel_code"] = cb["kernel_code"]
        out["cpu_baseline"]["same_config"] = ("same workload generator, read length, k, cases, MaxOutputPerRead and rescue fraction as the GPU arm; bounded sample"
                                              + ("; the CPU arm does not run the deep-DP stage of the both-unaligned pairs (the reference has no CPU DP engine): it does less work than the GPU arm" if with_deep else ""))
        if cb.get("gpu_reference_search"):
            out["search"]["reference_cuda_kernels_on_this_gpu"] = cb["gpu_reference_search"]
            if cb["gpu_reference_search"].get("reads_per_s"):
                out["search"]["speedup_over_reference_cuda_kernels"] = (reads_per_rank / t_search) / cb["gpu_reference_search"]["reads_per_s"]
        if cb.get("gpu_reference_dp"):
            out["dp"]["reference_cuda_kernels_on_this_gpu"] = cb["gpu_reference_dp"]
            if cb["gpu_reference_dp"].get("gcups"):
                out["dp"]["speedup_over_reference_cuda_kernels"] = dp_gcups / cb["gpu_reference_dp"]["gcups"]
        cpu_parity = None
        if "_hits" in cb:
            try:
                cpu_parity = cpu_search_parity(gi, cb, L)
                log("search vs the reference's CPU search at full size:", cpu_parity)
            except Exception as e:                           # noqa: BLE001
                cpu_parity = {"error": str(e)[:300]}
        del cb
        t0 = time.time()
        try:
            out["parity_at_full_size"] = chain_parity(gi, host, genome, L, args.parity_pairs, 4242, par, local_rank)
            log(f"chain parity on {args.parity_pairs} pairs took {time.time() - t0:.1f}s: {out['parity_at_full_size']}")
        except Exception as e:                               # noqa: BLE001
            out["parity_at_full_size"] = {"error": str(e)[:300]}
        out["parity_at_full_size"]["search_vs_reference_cpu_search"] = cpu_parity
        if with_deep:
            try:
                out["parity_at_full_size"]["deep_dp_stage"], _ = stage_parity(args, gi, host, genome, L, wpq, batches[-1].reads.cpu().numpy(), sp, False)
                log("deep DP stage parity at full size:", out["parity_at_full_size"]["deep_dp_stage"])
            except Exception as e:                           # noqa: BLE001
                out["parity_at_full_size"]["deep_dp_stage"] = {"error": str(e)[:300]}
    print(json.dumps(out), flush=True)
    for h in reversed(range(T)):                                  # clones are freed before the handle they were made from
        pes[h].free()
        api.GPUINDEXFree(gis[h])
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
