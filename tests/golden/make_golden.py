"""Generates tests/golden/*.json from the REFERENCE itself (run in the build
container, where /root/reference exists and oracle/build_ref.sh has produced
oracle/_ref/).  The fixtures pin the oracle (and through it the CUDA path):

  index_digests.json   sha256 of the files written by the reference's own
                       soap3-dp-builder + BGS-Build for a seeded genome
  search_golden.json   sha256 of the answer arrays produced by the reference's
                       search kernels (DV-Kernel.cu compiled for the host) for
                       every mismatch level / case / round on seeded reads,
                       plus a few fully spelled-out rows
  dp_golden.json       sha256 of scores/hitLocs/counts/patterns produced by the
                       reference's DP kernels (DV-DPfunctions.cu:35-512 compiled
                       for the host) on seeded batches

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import (ROOT, HostIndex, fmindex, formats, load_ref_dp, load_ref_search, make_dp_batch, pattern_end,  # noqa: E402
                     ref_dp, ref_launch)
from soap3dp_b200 import synth  # noqa: E402

GENOME_N, GENOME_SEED = 200_000, 1234


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_reads(G, L, n, seed):
    rs = synth.simulate_single_end(G, n, L, seed=seed, sub_rate=0.015)
    lens = rs.lengths.numpy().astype(np.uint32)
    lens[::5] = L - 1
    lens[3::11] = L - 7
    return rs.reads.numpy(), lens


def dp_digest(b, out):
    sc, hit, cnt, pat = out[:4]
    h = hashlib.sha256()
    for t in range(b.n):
        if sc[t] >= b.cutoff[t]:
            w = pat[t * b.pat_len:(t + 1) * b.pat_len]
            h.update(w[:pattern_end(w) + 1].tobytes())
        else:
            h.update(b"-")
    return {"scores_sha256": sha(sc[:b.n]), "hitLocs": sha(hit[:b.n]), "maxScoreCounts": sha(cnt[:b.n]), "patterns": h.hexdigest()}


def main():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    G = synth.random_genome(GENOME_N, seed=GENOME_SEED)
    # ---- index: run the reference's own builders ---------------------------
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "g.fa")
        seq = "".join(np.array(list("ACGT"))[G.numpy()])
        with open(fa, "w") as f:
            f.write(">chr1\n")
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")
        subprocess.check_call([os.path.join(ref_dir, "soap3-dp-builder"), fa], stdout=subprocess.DEVNULL)
        subprocess.check_call([os.path.join(ref_dir, "BGS-Build"), fa + ".index"], stdout=subprocess.DEVNULL)
        dig = {"genome": {"n": GENOME_N, "seed": GENOME_SEED, "sha256": sha(G.numpy())}}
        for ext in ("bwt", "rev.bwt", "fmv.gpu", "rev.fmv.gpu", "sa"):
            raw = np.fromfile(fa + ".index." + ext, dtype=np.uint32)
            dig[ext] = {"header": [int(x) for x in raw[:5]], "payload_sha256": sha(raw[5:])}
        ref_idx = fmindex.load_reference_index(fa + ".index")
    json.dump(dig, open(os.path.join(HERE, "index_digests.json"), "w"), indent=1)
    # ---- search: the reference kernels on the host --------------------------
    hi = HostIndex(ref_idx)
    rlib = load_ref_search()
    out = {"genome_seed": GENOME_SEED, "genome_n": GENOME_N, "sets": []}
    for L, n, seed in ((100, 1500, 7), (36, 1500, 8), (150, 1000, 9)):
        reads, lens = golden_reads(G, L, n, seed)
        wpq = formats.word_per_query(L)
        q = formats.pack_queries(reads, lens, wpq)
        lens_up = np.zeros(formats.ceil32(n), np.uint32)
        lens_up[:n] = lens
        entry = {"L": L, "n": n, "seed": seed, "launches": []}
        for k in range(5):
            for rnd, allowed in ((0, formats.SA_RANGES_ROUND1[k]), (1, formats.SA_RANGES_ROUND2[k])):
                wpa = 2 * allowed
                bad = np.zeros(formats.ceil32(n), np.uint8)
                qq = q.copy()
                for case in range(formats.NUM_CASES[k]):
                    if rnd == 1:
                        qq = q.copy()
                    a = np.zeros(formats.ceil32(n) * wpa, np.uint32)
                    nr = ref_launch(rlib, hi, case, qq, lens_up, n, wpq, a, bad, rnd, k, allowed, wpa, nthreads=1)
                    v = formats.answers_view(a, n, wpa)
                    entry["launches"].append({"k": k, "round": rnd, "case": case, "sha256": sha(v),
                                              "rank_queries_ref": int(nr),
                                              "rows": {str(i): [int(x) for x in v[i]] for i in (0, 1, 17, n - 1)}})
        out["sets"].append(entry)
    json.dump(out, open(os.path.join(HERE, "search_golden.json"), "w"), indent=0)
    # ---- DP: the reference kernels on the host ------------------------------
    dlib = load_ref_dp()
    dp = {"genome_seed": GENOME_SEED, "genome_n": GENOME_N, "batches": []}
    for mode in ("single", "rescue"):
        for L in (100, 150, 60):
            for scores in ((1, -2, -3, -1), (2, -3, -5, -2)):
                b = make_dp_batch(G, 600, L, mode, seed=L + len(mode), indel_rate=0.006)
                d = dp_digest(b, ref_dp(dlib, b, scores))
                d.update({"mode": mode, "L": L, "scores": list(scores), "n": 600})
                dp["batches"].append(d)
    json.dump(dp, open(os.path.join(HERE, "dp_golden.json"), "w"), indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
